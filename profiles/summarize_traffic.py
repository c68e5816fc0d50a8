"""DRAM traffic per launch of every engine kernel class, from one ncu capture of a few bench rounds:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/traffic_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline
    python profiles/summarize_traffic.py gpurun_out/traffic_r2.csv mnist [traffic_r2.json] > /tmp/t.json && mv /tmp/t.json profiles/traffic_r2.json   (merges datasets)
The classes are bench.py's (cgl_profile_* tags); a class launch = one Linear product (the weight-gradient class
includes its bias-gradient kernel, as the event pairs in bench.py do)."""
import csv
import json
import os
import re
import sys
from collections import defaultdict


def classify(name):
    # the TMA-fed kernels and the CTA-pair kernel: template arguments (A_KMAJOR, EPI, ...)
    t = re.search(r"tc_(?:tma_gemm|tma_persistent|pair_gemm)_kernel<(?:\(bool\))?(\d), (?:\(int\))?(\d)", name)
    if t:
        return ("linear_fwd[tcgen05]" if int(t.group(2)) == 0 else "linear_bwd_data[tcgen05]"), True
    m = re.search(r"tc_(?:grouped|persistent|sweep)_gemm_kernel<(?:\(bool\))?(\d), (?:\(bool\))?(\d), (?:\(int\))?(\d)", name)
    tag = "[tcgen05]"
    if not m:
        m = re.search(r"cgl::grouped_gemm_kernel<(?:\(bool\))?(\d), (?:\(bool\))?(\d), (?:\(int\))?(\d)", name)
        tag = "[ffma]"
    if m:
        a, b, epi = int(m.group(1)), int(m.group(2)), int(m.group(3))
        if epi == 0:
            return "linear_fwd" + tag, True
        if epi == 2:
            return "linear_wgrad+adam" + tag, True
        if epi == 1:
            return "linear_bwd_data" + tag, True
        # EPI_STORE: data gradient without a derivative (A MN-major, B K-major) or a stored weight gradient
        tc_bwd = (tag == "[tcgen05]" and a == 0 and b == 1) or (tag == "[ffma]" and a == 1 and b == 0)
        return ("linear_bwd_data" if tc_bwd else "linear_wgrad") + tag, True
    if "bias_grad_kernel<(bool)1>" in name or "bias_grad_kernel<1>" in name:
        return "linear_wgrad+adam[tcgen05]", False      # rides with the weight-gradient launch
    for key, cls in (("head_kernel", "head_loss"), ("head_stream_kernel", "head_loss"), ("bn_fwd_kernel", "batchnorm_fwd"), ("bn_bwd_kernel", "batchnorm_bwd"),
                     ("bn_fwd_smem_kernel", "batchnorm_fwd"), ("bn_bwd_smem_kernel", "batchnorm_bwd"),
                     ("wsum_kernel", "mix/aggregate"), ("bcast_mix_kernel", "mix/aggregate"),
                     ("mix_csr_kernel", "mix/aggregate"), ("dxg_reduce_kernel", "mix/aggregate"),
                     ("act_bwd_kernel", "elementwise")):
        if "cgl::" + key in name:
            return cls, True
    return None, False


def main(path, dataset):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per_launch = defaultdict(dict)
    names = {}
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
        per_launch[r["ID"]][r["Metric Name"]] = v * scale
        names[r["ID"]] = r["Kernel Name"]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])   # launches, read, write, ms
    for i, met in per_launch.items():
        cls, counts = classify(names[i])
        if cls is None:
            continue
        a = agg[cls]
        a[0] += 1 if counts else 0
        a[1] += met.get("dram__bytes_read.sum", 0.0)
        a[2] += met.get("dram__bytes_write.sum", 0.0)
        a[3] += met.get("gpu__time_duration.sum", 0.0)
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), sys.argv[3] if len(sys.argv) > 3 else "traffic_r2.json")
    doc = json.load(open(out_path)) if os.path.exists(out_path) else {}
    doc[dataset] = {c: {"dram_bytes_per_launch": (a[1] + a[2]) / max(a[0], 1), "dram_read_bytes_per_launch": a[1] / max(a[0], 1),
                        "dram_write_bytes_per_launch": a[2] / max(a[0], 1), "launches_captured": a[0],
                        "ms_per_launch_under_ncu": a[3] / max(a[0], 1)} for c, a in sorted(agg.items())}
    doc["_how"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                   "bench.py --steps 1 --warmup 1 (4 rounds); averages per class launch; profiles/summarize_traffic.py")
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "mnist")
