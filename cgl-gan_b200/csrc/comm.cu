// comm.cu -- the one collective on the path: cross-GPU sum of the packed aggregation vector.
// NCCL is resolved at run time (dlopen of the libnccl.so.2 that torch already mapped) so that the
// library loads on a CPU-only host for the ABI tests; every entry point fails with CGL_ENCCL if
// NCCL cannot be found.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

extern "C" int cgl_wsum(int C, int64_t n, const float* w, const int32_t* rows, const float* src, int64_t ld_src,
                        float* out, cgl_stream_t stream);

namespace cgl {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm*, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_get_error_string error_string = nullptr;
  bool tried = false;
};
static NcclApi g_nccl;

static bool load_nccl() {
  if (g_nccl.tried) return g_nccl.handle != nullptr;
  g_nccl.tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    set_error("NCCL not found: %s", dlerror());
    return false;
  }
  g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.all_reduce = (fn_all_reduce)dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.error_string = (fn_get_error_string)dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce) {
    set_error("NCCL symbols missing");
    g_nccl.handle = nullptr;
    return false;
  }
  return true;
}

struct Comm {
  nccl_comm comm;
  int nranks, rank;
};

static int nccl_fail(const char* what, int rc) {
  set_error("%s failed: %s", what, g_nccl.error_string ? g_nccl.error_string(rc) : "?");
  return CGL_ENCCL;
}

}  // namespace cgl

using namespace cgl;

extern "C" int cgl_comm_unique_id(uint8_t out_id[128]) {
  CGL_REQUIRE(out_id != nullptr, "out_id is NULL");
  if (!load_nccl()) return CGL_ENCCL;
  nccl_unique_id id;
  int rc = g_nccl.get_unique_id(&id);
  if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out_id, id.internal, 128);
  return CGL_OK;
}

extern "C" int cgl_comm_init(int nranks, int rank, const uint8_t id[128], cgl_comm_t* out_comm) {
  CGL_REQUIRE(id && out_comm && nranks > 0 && rank >= 0 && rank < nranks, "bad communicator arguments");
  if (!load_nccl()) return CGL_ENCCL;
  nccl_unique_id uid;
  memcpy(uid.internal, id, 128);
  Comm* c = new Comm();
  c->nranks = nranks;
  c->rank = rank;
  int rc = g_nccl.comm_init_rank(&c->comm, nranks, uid, rank);
  if (rc != 0) {
    delete c;
    return nccl_fail("ncclCommInitRank", rc);
  }
  *out_comm = c;
  return CGL_OK;
}

extern "C" int cgl_comm_destroy(cgl_comm_t comm) {
  if (!comm) return CGL_OK;
  Comm* c = (Comm*)comm;
  if (g_nccl.comm_destroy) g_nccl.comm_destroy(c->comm);
  delete c;
  return CGL_OK;
}

extern "C" int cgl_allreduce_sum(cgl_comm_t comm, float* buf, int64_t n, cgl_stream_t stream) {
  CGL_REQUIRE(comm && buf && n >= 0, "bad arguments");
  if (!load_nccl()) return CGL_ENCCL;
  Comm* c = (Comm*)comm;
  if (n == 0) return CGL_OK;
  // ncclFloat32 = 7, ncclSum = 0
  int rc = g_nccl.all_reduce(buf, buf, (size_t)n, 7, 0, c->comm, (cudaStream_t)stream);
  if (rc != 0) return nccl_fail("ncclAllReduce", rc);
  return CGL_OK;
}

extern "C" int cgl_mix_allreduce(cgl_comm_t comm, int C_local, int64_t n, const float* w_local, const int32_t* rows,
                                 const float* src, int64_t ld_src, float* out, cgl_stream_t stream) {
  CGL_REQUIRE(comm && out, "bad arguments");
  int rc = cgl_wsum(C_local, n, w_local, rows, src, ld_src, out, stream);
  if (rc) return rc;
  return cgl_allreduce_sum(comm, out, n, stream);
}
