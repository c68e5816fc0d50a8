"""Builds csrc/*.cu into lib/libcgl_b200.so for sm_100a with nvcc (in-tree, no JIT cache).
Every translation unit is compiled to an object on its own (in parallel, re-used while neither it nor any header
changed) and the objects are linked into the shared library."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libcgl_b200.so")
SOURCES = ["arch.cu", "dstep.cu", "gstep.cu", "mix.cu", "comm.cu", "data.cu", "fl.cu", "conv.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "cgl_b200.h"))
    return hs


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: a variant build (other -D switches) next to the product library, for profiles/ comparisons."""
    if out is None and not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else [])
    tag = hashlib.sha1(" ".join(flags).encode()).hexdigest()[:10]
    objdir = os.path.join(OBJDIR, tag)
    os.makedirs(objdir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    log = []

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, src[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(path)):
            return obj
        cmd = [nvcc] + flags + ["-c", "-o", obj, path]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        log.append(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out or LIB] + objs + ["-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write("".join(log))
    return out or LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out, extra=extra))
