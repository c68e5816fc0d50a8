// tc_gemm.cuh -- grouped GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM)
// with fp32-grade results through the 3xTF32 split (K2 of SURVEY.md section 2).
//
//   D[g][m][n] = sum_k A[g](m,k) * B[g](n,k)        one group g = one simulated client / edge server
//
// Orientation: the TMEM lane (= MMA M) dimension is always the dimension that is CONTIGUOUS in the
// output, so that the epilogue's 32 lanes of a warp touch 128 consecutive bytes:
//   forward   y [row][out]  : m = out feature, n = batch row,   k = in feature   (A = W  K-major, B = x  K-major)
//   data grad dx[row][in]   : m = in feature,  n = batch row,   k = out feature  (A = W  MN-major, B = dy K-major)
//   weight    W [out][in]   : m = in feature,  n = out feature, k = batch row    (A = x  MN-major, B = dy MN-major)
// The batch (200 = real|fake rows, or 100) is never the M dimension: M is tiled by 128 and the layer
// widths 784/512/256/1024 fill it, while N takes any multiple of 16 up to 256 (200 -> 208, 100 -> 112).
//
// Precision: the reference is strict fp32 (torch default, allow_tf32 = False). Each fp32 operand x is
// split in registers into hi = tf32(x) (round to nearest), lo = x - hi (exact); three MMAs per k-step accumulate
// lo*hi + hi*lo (correction region) and hi*hi (rotating main regions) in fp32 TMEM accumulators that the
// epilogue adds up (the dropped lo*lo term is ~2^-22 relative; see "TMEM plan" below for the regions).
// Because the split needs a register pass anyway, operands are staged global -> registers -> shared
// (whole 64/128-byte row segments, written straight into the canonical UMMA shared-memory layouts,
// conflict-free) instead of by TMA; ragged / concatenated / index-selected row sources
// (RowMap) come for free.
//
// Pipeline per CTA (128 x BN output tile, k-blocks of BKT = 32 (16) rows, 2..4 shared-memory stages, no CTA-wide
// barrier in the loop):
//   LW loader warps : LDG (2..3 k-blocks in flight per thread) -> split -> STS stage s -> fence.proxy.async
//                     -> mbarrier arrive full[s]; wait empty[s] before a stage is overwritten
//   next warp       : one thread waits full[s], issues 3 tcgen05.mma per k-step of 8 (the split products),
//                     tcgen05.commit -> empty[s]; a last commit -> done
//   epilogue        : the loader warps wait done, tcgen05.ld 32x32b.x16 of every region, fused bias+activation /
//                     activation derivative / Adam, coalesced stores.
// Variants in use (launch_tc_gemm, linear.cuh):
//   forward        : LW = 16, one CTA per SM, 3 stages            (the loader loop is the limiter: 4 warps per scheduler)
//   data gradient  : tc_persist.cuh (persistent, dedicated epilogue warps); batch tiles wider than 128: LW = 8 here
//   weight gradient: OCC = 2 (two CTAs per SM), BKT = 16, two 32 KB stages, fused Adam epilogue
#pragma once
#include <stdlib.h>

#include "gemm.cuh"

namespace cgl {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_LOADER_THREADS = 256;  // warps 0..7: operand staging, then the epilogue
constexpr int TC_MMA_WARP = 8;          // warp 8: TMEM allocation and the single MMA-issuing thread
constexpr int TC_THREADS = TC_LOADER_THREADS + 32;
#ifndef TC_WG_UNR
#define TC_WG_UNR 3      // float4 triples (W, m, v) in flight per thread in the Adam epilogue of the weight-gradient kernel
#endif
#ifndef TC_WG_DEPTH
#define TC_WG_DEPTH 3    // 16-row k-blocks of loads in flight per thread of the weight-gradient kernel
#endif
#ifndef TC_WG_STAGES
#define TC_WG_STAGES 2   // its 32 KB stages (the Adam epilogue's transpose buffer needs 64 KB)
#endif
#ifndef TC_PF_BLOCKS
#define TC_PF_BLOCKS 8    // distance (k-blocks) of the 512-byte row prefetches ahead of the operand loads (tune bit 2048)
#endif
#ifndef TC_LW16_DEPTH
#define TC_LW16_DEPTH 3   // k-blocks of loads in flight per thread of the 16-loader-warp forward kernel
#endif
constexpr int TC_MAX_BN = 256;
constexpr int TC_TUNE_DEFAULT = 1 | 8 | 32 | 64 | 256 | 512 | 1024 | 131072;
constexpr uint32_t TC_WAIT_HINT_NS = 20000u;

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef TC_WAIT_MODE
#define TC_WAIT_MODE 0   // 0: try_wait with a suspend-time hint; 1: try_wait with the default time limit; 2: test_wait (pure polling)
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
#if TC_WAIT_MODE == 1
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
#elif TC_WAIT_MODE == 2
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
#endif
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(TC_WAIT_HINT_NS)   // the thread may stay suspended this long: fewer polls,
      : "memory");                                    // and the wake-up is still signalled by the barrier
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if TC_WAIT_MODE != 0
  const long long c0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - c0 > 4000000000ll) __trap();   // ~2 s: a lost arrival must fail loudly, not hang the GPU
  }
  return;
#endif
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) __trap();   // 2 s: a lost arrival must fail loudly, not hang the GPU
  }
}
// a wait that lasts a whole main loop (the epilogue warps of the persistent kernel): back off between polls so the
// waiting warp leaves the issue slots of its scheduler to the loaders
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(200);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32 (K = 8 per instruction)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// The issue loop runs on a CONVERGED warp: every lane executes it with identical operands and the instructions themselves run
// on the lane elect.sync picks. Inside an `if (lane == 0)` region nvcc wraps each tcgen05.mma in a waterfall loop (ELECT /
// R2UR / UTCHMMA / BRA.U.ANY) and the issuing thread needs ~135 clk per MMA -- more than twice what the tensor core takes
// for a 128 x 112 x 8 tf32 MMA (56 clk; profiles/umma_rate_probe_r2.log, profiles/tma_ablate_r2.log).
// One k-step of the 3xTF32 product: corr += lo*hi, corr += hi*lo, main (+)= hi*hi.
__device__ __forceinline__ void umma_kstep_ss_warp(uint32_t d_corr, uint32_t d_main, uint64_t dah, uint64_t dal, uint64_t dbh,
                                                   uint64_t dbl, uint32_t idesc, uint32_t acc_corr, uint32_t acc_main) {
  asm volatile(
      "{\n\t.reg .pred pc, pm, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pc, %7, 0;\n\t"
      "setp.ne.b32 pm, %8, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %3, %4, %6, pc;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %5, %6, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], %2, %4, %6, pm;\n\t}"
      ::"r"(d_corr), "r"(d_main), "l"(dah), "l"(dal), "l"(dbh), "l"(dbl), "r"(idesc), "r"(acc_corr), "r"(acc_main)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(bar)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// HBM -> L2 prefetch of `bytes` (multiple of 16, 16-byte aligned address): no registers, no completion to wait for
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// hi = x rounded to tf32 (10 mantissa bits), round-half-up in magnitude: two integer instructions.
// (cvt.rna.tf32.f32 compiles to a ~10-instruction emulation on sm_100a and dominated the loader warps.)
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: start >> 4 in [0,14), leading byte
// offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type in
// [61,64): 0 = SWIZZLE_NONE, 1 = SWIZZLE_128B_BASE32B -- the only layout tf32 accepts MN-major --,
// 2 = SWIZZLE_128B).
constexpr uint32_t UMMA_LAYOUT_NONE = 0, UMMA_LAYOUT_SW128_BASE32B = 1, UMMA_LAYOUT_SW128 = 2;
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4.
__host__ __device__ inline uint32_t umma_idesc_tf32(bool a_mn_major, bool b_mn_major, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// ---- operand staging --------------------------------------------------------------------------
// A staged tile is T (= 128, or BN rounded up to 32) lines of the M/N dimension by TC_BK = 32 of K, as
// hi and lo copies. One warp-wide float4 access covers a "patch" whose 32 float4s land in 512
// consecutive bytes of shared memory (conflict-free) and read whole 64/128-byte row segments.
//   K-major  (k contiguous in global; lines = m/n), SWIZZLE_128B: line t is one 128-byte row (32 k), atoms of
//       8 rows, the 16-byte chunk index XORed with the row index (Swizzle<3,4,3> on the byte address):
//       byte(t, k) = t*128 + (((k/4) ^ (t%8)) * 16) + (k%4)*4            SBO = 1024 (next 8 lines); LBO unused
//       patch p = 4 lines x 32 k;  lane -> (t = 4*p + lane/8, k = 4*(lane%8)): every warp load reads four whole
//       128-byte rows (full cache lines: the SM's outstanding-request budget is what bounds the loaders)
//       one MMA (8 k) = 32 B further inside the swizzle span
//   MN-major (m/n contiguous in global; lines = k), SWIZZLE_128B_BASE32B, atom = 4 k x 32 t (4 rows of 128 B,
//       the 32-byte chunk index XORed with the row index: Swizzle<2,5,2> on the byte address):
//       byte(t, k) = (t/32)*4096 + (k/4)*512 + (k%4)*128 + ((((t%32)/8) ^ (k%4)) * 32) + (t%8)*4
//       LBO = 4096 (next 32 t), SBO = 512 (next 4 k); one MMA (8 k) = 2 atoms = 1024 B further
//       patch p = 4 k x 32 t: k group p%8, t group p/8;  lane -> (k = 4*(p%8) + lane/8, t = 32*(p/8) + 4*(lane%8))
// BKT = K extent of a staged k-block: 32, or 16 for MN-major operands only (KG = BKT / 4 groups of 4 k-rows per 32 lines;
// the K-major layout is tied to 32: one 128-byte swizzle row)
// operand loads: read-only path (default) or, with -DTC_LD_STREAM, streaming loads that do not allocate in L1 (an operand
// element is read once per CTA and the CTA leaves ~30 KB of L1)
__device__ __forceinline__ float4 tc_ldg4(const float* p) {
#ifdef TC_LD_STREAM
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
#else
  return __ldg(reinterpret_cast<const float4*>(p));
#endif
}
template <bool KMAJOR, int BKT = 32>
__device__ __forceinline__ float4 tc_patch_load(const Rows& R, int p, int lane, int t0, int dimT, int k0, int dimK) {
  static_assert(BKT == 32 || (!KMAJOR && BKT == 16), "unsupported k-block");
  constexpr int KG = BKT / 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (KMAJOR) {
    const int t = t0 + 4 * p + (lane >> 3);
    const int k = k0 + 4 * (lane & 7);
    if (t < dimT && k < dimK) v = tc_ldg4(row_ptr(R, t) + k);
  } else {
    const int k = k0 + 4 * (p % KG) + (lane >> 3);
    const int t = t0 + 32 * (p / KG) + 4 * (lane & 7);
    if (k < dimK && t < dimT) v = tc_ldg4(row_ptr(R, k) + t);
  }
  return v;
}
template <bool KMAJOR, int BKT = 32>
__device__ __forceinline__ uint32_t tc_patch_offset(int p, int lane) {
  constexpr int KG = BKT / 4;
  if (KMAJOR) {
    const int r = 4 * p + (lane >> 3);
    return (uint32_t)(r * 128 + (((lane & 7) ^ (r & 7)) << 4));
  }
  const int kr = lane >> 3, c16 = lane & 7;
  return (uint32_t)((p / KG) * (KG * 512) + (p % KG) * 512 + kr * 128 + (((c16 >> 1) ^ kr) << 5) + (c16 & 1) * 16);
}
__device__ __forceinline__ void tc_split_store(char* hi, char* lo, uint32_t off, const float4& v) {
  // lo = x - hi is exact in fp32 (|lo| <= 2^-11 |x|, either sign); the tensor core drops its low 13 bits,
  // an error below 2^-21 |x| without a preferred direction relative to x
  float4 h, l;
  h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

// Bring-up instrumentation: when cgl_debug_set_timeline() installs a buffer, CTA (0,0,g) writes clock64() stamps
// of its milestones to timeline[g*16 + i] (i: 0 entry, 1 setup done, 2 first loads issued, 3 first stage
// stored, 4 last stage stored, 5 accumulator complete, 6 epilogue done, 7 exit; 8 MMA first full, 9 MMA last commit).
__device__ long long* g_tc_timeline = nullptr;
#define TC_STAMP(i)                                                                      \
  do {                                                                                   \
    if (g_tc_timeline && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) g_tc_timeline[blockIdx.z * 16 + (i)] = clock64(); \
  } while (0)

// EPI_* as in gemm.cuh. Output element (m, n) lives at C + n*ldc + m.
struct TcParams {
  int M, N, K;
  int bn;         // N tile (multiple of 16, <= 256)
  int n_per;      // tc_tma.cuh: row pitch of the batch tiles (<= bn)
  int n_stages;   // shared-memory stages (1..TC_MAX_STAGES)
  int n_main;     // hi*hi accumulator regions (see "TMEM plan")
  int tmem_cols;  // TMEM columns to allocate: power of two >= (n_main + 1) * round_up(bn, 32)
  RowMap A, B;
  float* cbase; long long c_gstride; const int* cidx; long long c_off; int ldc;
  int c_vec;  // host-verified: every output row start (and adam_m / adam_v) is 16-byte aligned, ldc % 4 == 0
  const float* bias_base; long long bias_gstride; const int* bias_idx; long long bias_off;  // bias[m] (EPI_FWD)
  int act; float slope;
  const float* saved; long long saved_gstride;  // EPI_BWD_DATA: saved[g] + n*ldc + m
  float* adam_m; float* adam_v; const int* step; float lr, b1, b2, eps;  // EPI_ADAM
  const AdamScalars* scal;  // EPI_ADAM: [G] precomputed scalars of this step (NULL: derive from step)
  int tune;  // bits (tc_tune()): 1 = L2 prefetch of the Adam tile under the main loop, 8 / 16 = persistent kernel
             // (tc_persist.cuh) for the data-gradient / forward product, 32 = 16 loader warps in the one-tile forward kernel,
             // 64 = 16-row stages (two of them) in the weight-gradient kernel, 128 = 16 loader warps in the one-tile
             // data-gradient kernel (with 8 cleared; measured equal to the persistent kernel: 5.26 against 5.30 ms per round),
             // 256 / 512 = CTA pairs (tc_pair.cuh, cta_group::2) for the forward / data-gradient product,
             // 1024 = two CTAs per SM for short-K (<= 43 k-steps) forward / data-gradient products,
             // 2048 = 512-byte row prefetches into L2, 4096 / 8192 = bring-up ablations (no global loads / no MMAs),
             // 16384 / 32768 = persistent kernel with the lean 16-warp loader loop (tc_sweep.cuh) for forward / data gradient,
             // 65536 = CTA pairs also for an odd number of M tiles,
             // 131072 = TMA-fed kernel with the weights in TMEM (tc_tma.cuh) for forward / data gradient (524288: its B operand
             // through registers instead of TMA, 1048576: its one-thread MMA issue loop; 4096 / 8192: its ablations; 2097152: only the
             // hi*hi regions the accumulation cap needs; 8388608: CTA pairs; 16777216: persistent variant, tc_tma_persist.cuh;
             // 33554432: the TMA-written batch tile stays in place as hi (the tensor core truncates), only lo is stored;
             // 4: TWO issuing warps that take the k-blocks in turn (token barrier); 2: look-ahead barrier tests in the (single)
             // issuing warp; 262144: A converters load the next k-block under their TMEM stores; bring-up ablations of the
             // feed: 67108864 / 134217728 no B / A transfers, 268435456 / 536870912 A converters / B warps run their barrier
             // protocol only, 1073741824 two hi*hi regions at any K (timing only) -- profiles/tma_feed_abl_r2.log)
};

// TMEM plan (512 columns, 1 CTA per SM). The tensor core TRUNCATES when it adds into an fp32
// accumulator, a one-sided error that grows with the number of accumulations and that the reductions
// downstream (BatchNorm, weight gradients, the client sum) do not average away. So the accumulator is
// split into regions of `stride` = round_up(bn, 32) columns:
//   region 0            : the two correction products (lo*hi + hi*lo), 2^-11 of the result
//   regions 1..n_main   : hi*hi, k-step ks goes to region 1 + ks % n_main
// and the epilogue adds the regions in fp32 with round-to-nearest. The host picks bn so that no region
// takes more than TC_MAX_ACCUM accumulations (tc_pick_bn).
constexpr int TC_MAX_ACCUM = 43;  // K = 1024 -> 128 k-steps over 3 regions (one 112-wide tile for a batch of 100)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_MAX_MAIN = 7;
__host__ __device__ inline int tc_region_stride(int bn) { return (bn + 31) & ~31; }
__host__ __device__ inline int tc_n_main(int bn) {
  int r = TC_TMEM_COLS / tc_region_stride(bn) - 1;
  return r > TC_MAX_MAIN ? TC_MAX_MAIN : r;
}

// NB = B patches per loader warp (4: bn <= 128, 8: bn <= 256); 8 loader warps + 1 MMA warp.
// OCC = CTAs per SM the variant is built for. OCC 2 (short-K weight gradients: one stage, half of TMEM, fewer
// registers) lets one CTA's HBM-bound Adam epilogue run under another CTA's main loop.
// LW = loader warps (8, or 16 for the long-K one-CTA-per-SM variants: the loader loop is ~330 instructions per
// k-block at ~0.35 IPC per scheduler with two warps each, so four warps per scheduler hide more of its latency)
// BKT = K extent of a stage (32; 16 for the two-CTAs-per-SM weight-gradient variant: two 32 KB stages instead of one
// of 64 KB, so a store overlaps the MMAs of the previous k-block, and 16 registers of loads per k-block instead of 32)
template <bool A_KMAJOR, bool B_KMAJOR, int EPI, int NB, int OCC, int LW, int BKT>
__global__ void __launch_bounds__(LW * 32 + 32, OCC) tc_grouped_gemm_kernel(const TcParams p) {
  constexpr int LT = LW * 32;          // loader (= epilogue) threads
  constexpr int MMAW = LW;             // the warp after them allocates TMEM and issues the MMAs
  constexpr int KG = BKT / 4;          // groups of 4 k-rows per k-block
  constexpr int NA = 4 * KG / LW;      // A patches per loader warp (a 128-line tile has 4 * KG)
  constexpr int NBW = NB * KG / LW;    // B patches per loader warp
  extern __shared__ __align__(1024) char tc_smem[];
  __shared__ __align__(8) unsigned long long bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_done;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.z;
  const int m0 = blockIdx.y * TC_BM;
  const int n0 = blockIdx.x * p.bn;
  const int bn = p.bn;
  const int nst = p.n_stages;
  if (warp == 0) TC_STAMP(0);

  // stage layout: [A hi | A lo | B hi | B lo]
  const uint32_t a_bytes = TC_BM * BKT * 4;
  const int bn_pad = (bn + 31) & ~31;  // MN-major staging works in groups of 32 lines
  const uint32_t b_bytes = (uint32_t)bn_pad * BKT * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  // the swizzled layout XORs absolute address bits [7,9) into [5,7): every buffer starts 1024-aligned
  char* smem = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);

  const int stride = tc_region_stride(bn);
  const int n_main = p.n_main;
  const uint32_t tmem_cols = (uint32_t)p.tmem_cols;
  const int nkb = (p.K + BKT - 1) / BKT;
  const int nks = (p.K + 7) >> 3;  // k-steps of 8 that carry data

  if (tid == 0) {
    for (int i = 0; i < nst; ++i) {
      mbar_init(smem_u32(&bar_full[i]), LT);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    mbar_init(smem_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMAW) tmem_alloc(smem_u32(&tmem_slot), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (warp == 0) TC_STAMP(1);

  if (warp == MMAW) {
    // ===== MMA issuer: the warp runs the loop converged, the elected lane issues (umma_kstep_ss_warp) =====
    {
      const uint32_t idesc = umma_idesc_tf32(!A_KMAJOR, !B_KMAJOR, bn);
      // K-major : SWIZZLE_128B, SBO = 1024 (next 8 lines), LBO unused, a k-step of 8 = 32 B inside the span
      // MN-major: SWIZZLE_128B_BASE32B, LBO = 4096 (next 32 lines), SBO = 512 (next 4 k), a k-step of 8 = 1024 B
      const uint32_t a_lbo = A_KMAJOR ? 16u : (uint32_t)(KG * 512), a_sbo = A_KMAJOR ? 1024u : 512u;
      const uint32_t b_lbo = B_KMAJOR ? 16u : (uint32_t)(KG * 512), b_sbo = B_KMAJOR ? 1024u : 512u;
      constexpr uint64_t a_dstep = (A_KMAJOR ? 32u : 1024u) >> 4, b_dstep = (B_KMAJOR ? 32u : 1024u) >> 4;   // descriptor units
      const uint32_t a_lay = A_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      const uint32_t b_lay = B_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
      const uint32_t main_lo = tmem_u + (uint32_t)stride, main_hi = tmem_u + (uint32_t)(n_main * stride);
      const uint32_t bar_f = smem_u32(&bar_full[0]), bar_e = smem_u32(&bar_empty[0]);
      const uint32_t smem0 = smem_u32(smem);
      const bool no_mma = (p.tune & 8192) != 0;   // (bring-up: what the loop costs without the tensor core)
      uint32_t d_main = main_lo;
      int ks = 0, s = 0;  // k-step counter, stage
      uint32_t par = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(bar_f + 8u * s, par);
        if (kb == 0) TC_STAMP(8);
        tc_fence_after();
        const uint32_t sa_hi = smem0 + (uint32_t)s * stage_bytes;
        const uint64_t dah0 = umma_desc(sa_hi, a_lbo, a_sbo, a_lay);
        const uint64_t dal0 = umma_desc(sa_hi + a_bytes, a_lbo, a_sbo, a_lay);
        const uint64_t dbh0 = umma_desc(sa_hi + 2 * a_bytes, b_lbo, b_sbo, b_lay);
        const uint64_t dbl0 = umma_desc(sa_hi + 2 * a_bytes + b_bytes, b_lbo, b_sbo, b_lay);
#pragma unroll
        for (int j = 0; j < BKT / 8; ++j) {
          if (ks < nks) {
            if (!no_mma)
              umma_kstep_ss_warp(tmem_u, d_main, dah0 + a_dstep * j, dal0 + a_dstep * j, dbh0 + b_dstep * j, dbl0 + b_dstep * j,
                                 idesc, ks > 0 ? 1u : 0u, ks >= n_main ? 1u : 0u);
            d_main = (d_main == main_hi) ? main_lo : d_main + (uint32_t)stride;
            ++ks;
          }
        }
        umma_commit_warp(bar_e + 8u * s);
        if (++s == nst) { s = 0; par ^= 1u; }
      }
      umma_commit_warp(smem_u32(&bar_done));
      TC_STAMP(9);
    }
    __syncwarp();
  } else {
    // epilogue operands that do not depend on the accumulator are requested now, under the main loop
    const int q = warp & 3;        // TMEM lane quarter this warp may read
    const int half = warp >> 2;    // the LW/4 warps of a quarter take the 16-column chunks in turn
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < p.M;
    const int rowid = p.cidx ? p.cidx[g] : g;
    float bias = 0.f;
    if (EPI == EPI_FWD && p.bias_base && m_ok) {
      const int brow = p.bias_idx ? p.bias_idx[g] : g;
      bias = __ldg(p.bias_base + (long long)brow * p.bias_gstride + p.bias_off + m);
    }
    AdamScalars as = {};
    if (EPI == EPI_ADAM) as = p.scal ? p.scal[g] : make_adam_scalars(p.step[rowid], p.lr, p.b1, p.b2, p.eps);

    // ===== loader warps: global -> registers (several k-blocks in flight) -> split -> shared =====
    const Rows RA = resolve(p.A, g);
    const Rows RB = resolve(p.B, g);
    // patches of the B tile that hold data. K-major B (lines = batch rows): the lines beyond N are neither loaded nor
    // stored -- whatever the stage holds there only reaches accumulator columns n >= N, which the epilogue never
    // reads (for a batch of 100 in a 112-wide tile that is 7 of 32 patches of STS traffic on the L1TEX pipe).
    const int b_lines = (p.N - n0 < bn) ? (p.N - n0) : bn;
    const int npb = B_KMAJOR ? ((b_lines + 3) >> 2) : (bn_pad >> 5) * KG;
    constexpr int DEPTH = (LW == 16) ? TC_LW16_DEPTH : (BKT == 16 ? TC_WG_DEPTH : ((NB == 4 && OCC == 1) ? 3 : 2));  // k-blocks of global loads in flight per thread
    float4 ra[DEPTH][NA], rb[DEPTH][NBW];
    auto load_block = [&](int kb, float4 (&qa)[NA], float4 (&qb)[NBW]) {
      const int k0 = kb * BKT;
      if (p.tune & 4096) {          // bring-up: no global loads at all (what the loop costs without the memory system)
#pragma unroll
        for (int i = 0; i < NA; ++i) qa[i] = make_float4(1.f, 2.f, 3.f, (float)kb);
#pragma unroll
        for (int i = 0; i < NBW; ++i) qb[i] = make_float4(1.f, 2.f, 3.f, (float)kb);
        return;
      }
#pragma unroll
      for (int i = 0; i < NA; ++i) qa[i] = tc_patch_load<A_KMAJOR, BKT>(RA, warp + LW * i, lane, m0, p.M, k0, p.K);
#pragma unroll
      for (int i = 0; i < NBW; ++i) {
        const int pp = warp + LW * i;
        qb[i] = (pp < npb) ? tc_patch_load<B_KMAJOR, BKT>(RB, pp, lane, n0, p.N, k0, p.K) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    int st_s = 0;            // stage of the next store
    uint32_t st_par = 1;     // parity to wait for on empty[st_s]; nothing to wait for during the first pass
    bool st_first = true;
    auto store_block = [&](const float4 (&qa)[NA], const float4 (&qb)[NBW]) {
      if (!st_first) mbar_wait(smem_u32(&bar_empty[st_s]), st_par);  // MMAs that read this stage are done
      char* a_hi = smem + (size_t)st_s * stage_bytes;
      char* a_lo = a_hi + a_bytes;
      char* b_hi = a_hi + 2 * a_bytes;
      char* b_lo = b_hi + b_bytes;
#pragma unroll
      for (int i = 0; i < NA; ++i) tc_split_store(a_hi, a_lo, tc_patch_offset<A_KMAJOR, BKT>(warp + LW * i, lane), qa[i]);
#pragma unroll
      for (int i = 0; i < NBW; ++i) {
        const int pp = warp + LW * i;
        if (pp < npb) tc_split_store(b_hi, b_lo, tc_patch_offset<B_KMAJOR, BKT>(pp, lane), qb[i]);
      }
      // generic-proxy stores -> visible to the tensor core (async proxy), then one arrival per thread. (One arrival per
      // WARP after a __syncwarp was measured: 3 % slower -- the arrivals are not what the loop costs.)
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bar_full[st_s]));
      if (++st_s == nst) {
        st_s = 0;
        if (st_first) { st_first = false; st_par = 0; } else { st_par ^= 1u; }
      }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
      if (d < nkb) load_block(d, ra[d], rb[d]);
    if (warp == 0) TC_STAMP(2);
    auto adam_tile_prefetch = [&]() {
      if (!(EPI == EPI_ADAM && p.c_vec && (p.tune & 1))) return;
      // the W / m / v tile this CTA streams in its epilogue (24 B per parameter, the HBM-bound part of a round)
      // starts its way HBM -> L2 now, under the main loop: one row of the tile (<= 512 bytes) per request
      const int n_valid = (p.N - n0 < bn) ? (p.N - n0) : bn;
      const uint32_t m_bytes = (uint32_t)(((p.M - m0 < TC_BM) ? (p.M - m0) : TC_BM) * 4);
      const long long tile0 = (long long)rowid * p.c_gstride + p.c_off + (long long)n0 * p.ldc + m0;
      for (int n = tid; n < n_valid; n += LT) {
        const long long off = tile0 + (long long)n * p.ldc;
        l2_prefetch_bulk(p.cbase + off, m_bytes);
        l2_prefetch_bulk(p.adam_m + off, m_bytes);
        l2_prefetch_bulk(p.adam_v + off, m_bytes);
      }
    };
    const int kb_prefetch = nkb > DEPTH ? nkb - DEPTH - 1 : 0;   // right after the last operand loads are issued
    // K-major operands are read as one 128-byte segment per row per k-block: 128 (+ batch) rows that lie K*4 bytes apart.
    // DRAM then serves 128-byte requests scattered over as many pages, at ~2 TB/s (ncu: 29 % of peak) with every loader
    // warp waiting on its loads. TC_PF_BLOCKS k-blocks ahead of the loads, one thread per row asks L2 for the next
    // 512 contiguous bytes of its row (4 k-blocks) in ONE request, so that DRAM sees long bursts.
    constexpr int PF_SPAN = 4;                     // k-blocks per prefetch request (4 x 128 B)
    auto burst_prefetch = [&](int kb) {
      if (!(p.tune & 2048) || (kb % PF_SPAN) != 0) return;
      const int k = (kb + TC_PF_BLOCKS) * BKT;
      if (k >= p.K) return;
      const uint32_t bytes = (uint32_t)(((p.K - k < PF_SPAN * BKT) ? (p.K - k) : PF_SPAN * BKT) * 4);
      if (A_KMAJOR && tid < TC_BM) {
        if (m0 + tid < p.M) l2_prefetch_bulk(row_ptr(RA, m0 + tid) + k, bytes);
      } else if (B_KMAJOR && tid >= TC_BM && tid < TC_BM + b_lines) {
        l2_prefetch_bulk(row_ptr(RB, n0 + tid - TC_BM) + k, bytes);
      }
    };
    if (p.tune & 2048) {                           // the rows' first TC_PF_BLOCKS k-blocks, before the first loads
      for (int kb = -TC_PF_BLOCKS; kb < 0; kb += PF_SPAN) burst_prefetch(kb - (kb % PF_SPAN));
    }

    for (int kb0 = 0; kb0 < nkb; kb0 += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int kb = kb0 + d;
        if (kb < nkb) {
          store_block(ra[d], rb[d]);
          if (kb == 0 && warp == 0) TC_STAMP(3);
          if (kb + DEPTH < nkb) load_block(kb + DEPTH, ra[d], rb[d]);
          burst_prefetch(kb);
          if (kb == kb_prefetch) adam_tile_prefetch();
        }
      }
    }
    if (warp == 0) TC_STAMP(4);

    // ===== epilogue: TMEM -> registers -> fused op -> global =====
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    if (warp == 0) TC_STAMP(5);

    float* C = p.cbase + (long long)rowid * p.c_gstride + p.c_off;
    const int n_used = nks < n_main ? nks : n_main;

    float* Mo = nullptr;
    float* Vo = nullptr;
    if (EPI == EPI_ADAM) {
      Mo = p.adam_m + (long long)rowid * p.c_gstride + p.c_off;
      Vo = p.adam_v + (long long)rowid * p.c_gstride + p.c_off;
    }
    const float* S = (EPI == EPI_BWD_DATA && p.saved) ? p.saved + (long long)g * p.saved_gstride : nullptr;

    const int nch = bn >> 4;
    // accumulator chunk: 16 columns of this thread's lane, summed over the TMEM regions
    auto tmem_chunk = [&](int c, float (&v)[16]) {
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
      if (OCC == 2) {  // register-lean variant: always exactly one main region plus the corrections
        float t[16];
        tmem_ld16(taddr + (uint32_t)stride, v);
        tmem_ld16(taddr, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += t[j];
        return;
      }
      // up to four regions are requested before one wait (the usual plans have 2 or 4 regions)
      uint32_t r0[16], r1[16], r2[16], r3[16];
      tmem_ld16_async(taddr + (uint32_t)stride, r0);                    // main region 1
      tmem_ld16_async(taddr, r1);                                       // corrections
      if (n_used >= 2) tmem_ld16_async(taddr + (uint32_t)(2 * stride), r2);
      if (n_used >= 3) tmem_ld16_async(taddr + (uint32_t)(3 * stride), r3);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = __uint_as_float(r0[j]);
        if (n_used >= 2) a += __uint_as_float(r2[j]);
        if (n_used >= 3) a += __uint_as_float(r3[j]);
        v[j] = a;
      }
      float t[16];
      for (int r = 4; r <= n_used; ++r) {
        tmem_ld16(taddr + (uint32_t)(r * stride), t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += t[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r1[j]);      // the small correction last
    };
    auto chunk_ok = [&](int c) { return c < nch && n0 + c * 16 < p.N; };  // warp-uniform

    if (EPI == EPI_ADAM && p.c_vec) {
      // Adam is the HBM-bound part of a round (24 B per parameter). The accumulator tile goes TMEM -> shared
      // ([n][m], m contiguous like W) through the operand stages, which are idle now; then all 256 threads
      // stream W / m / v as float4 along m: 512 contiguous bytes per warp request and UNR*3 independent 16-byte
      // loads in flight per thread, instead of 4-byte accesses that leave the memory pipeline mostly empty.
      float* T = reinterpret_cast<float*>(smem);
      for (int c = half; chunk_ok(c); c += LW / 4) {
        float gv[16];
        tmem_chunk(c, gv);
#pragma unroll
        for (int j = 0; j < 16; ++j) T[(c * 16 + j) * TC_BM + q * 32 + lane] = gv[j];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(LT) : "memory");  // the 8 loader / epilogue warps only
      const int n_valid = (p.N - n0 < bn) ? (p.N - n0) : bn;
      const int m4_valid = ((p.M - m0 < TC_BM) ? (p.M - m0) : TC_BM) >> 2;   // M % 4 == 0 (MN-major A operand)
      const int items = n_valid * (TC_BM / 4);
      constexpr int UNR = (OCC == 2) ? TC_WG_UNR : 4;
      for (int i0 = tid; i0 < items; i0 += LT * UNR) {
        float4 w4[UNR], a4[UNR], v4[UNR];
        long long off[UNR];
        bool ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int i = i0 + u * LT;
          const int n = i >> 5, mq = i & 31;
          ok[u] = i < items && mq < m4_valid;
          off[u] = (long long)(n0 + n) * p.ldc + m0 + mq * 4;
          if (ok[u]) {
            w4[u] = *reinterpret_cast<const float4*>(C + off[u]);
            a4[u] = *reinterpret_cast<const float4*>(Mo + off[u]);
            v4[u] = *reinterpret_cast<const float4*>(Vo + off[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (ok[u]) {
            const int i = i0 + u * LT;
            const float4 g4 = *reinterpret_cast<const float4*>(T + (i >> 5) * TC_BM + (i & 31) * 4);
            adam_update_fast(w4[u].x, a4[u].x, v4[u].x, g4.x, as);
            adam_update_fast(w4[u].y, a4[u].y, v4[u].y, g4.y, as);
            adam_update_fast(w4[u].z, a4[u].z, v4[u].z, g4.z, as);
            adam_update_fast(w4[u].w, a4[u].w, v4[u].w, g4.w, as);
            *reinterpret_cast<float4*>(C + off[u]) = w4[u];
            *reinterpret_cast<float4*>(Mo + off[u]) = a4[u];
            *reinterpret_cast<float4*>(Vo + off[u]) = v4[u];
          }
        }
      }
    } else if (EPI == EPI_ADAM) {
      // unaligned packed rows: scalar accesses, one TMEM chunk at a time
      for (int c = half; chunk_ok(c); c += LW / 4) {
        const int nb = n0 + c * 16;
        float g[16];
        tmem_chunk(c, g);
        if (!m_ok) continue;
        float w[16], mm[16], vv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const long long off = (long long)(nb + j) * p.ldc + m;
          const bool ok = nb + j < p.N;
          w[j] = ok ? C[off] : 0.f;
          mm[j] = ok ? Mo[off] : 0.f;
          vv[j] = ok ? Vo[off] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (nb + j < p.N) {
            const long long off = (long long)(nb + j) * p.ldc + m;
            adam_update_fast(w[j], mm[j], vv[j], g[j], as);
            C[off] = w[j]; Mo[off] = mm[j]; Vo[off] = vv[j];
          }
        }
      }
    } else if (EPI != EPI_ADAM && p.c_vec) {
      // same route for the other epilogues: accumulators to shared ([n][m]), then float4 rows of the output
      // (512 contiguous bytes per warp store; bias / saved activations as float4 too)
      float* T = reinterpret_cast<float*>(smem);
      for (int c = half; chunk_ok(c); c += LW / 4) {
        float gv[16];
        tmem_chunk(c, gv);
#pragma unroll
        for (int j = 0; j < 16; ++j) T[(c * 16 + j) * TC_BM + q * 32 + lane] = gv[j];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(LT) : "memory");
      const int n_valid = (p.N - n0 < bn) ? (p.N - n0) : bn;
      const int m4_valid = ((p.M - m0 < TC_BM) ? (p.M - m0) : TC_BM) >> 2;   // M % 4 == 0 with c_vec
      const int items = n_valid * (TC_BM / 4);
      const int mq = tid & 31;                       // this thread's float4 column of the tile: fixed over the loop
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI == EPI_FWD && p.bias_base && mq < m4_valid) {
        const int brow = p.bias_idx ? p.bias_idx[g] : g;
        b4 = __ldg(reinterpret_cast<const float4*>(p.bias_base + (long long)brow * p.bias_gstride + p.bias_off + m0 + mq * 4));
      }
      constexpr int UNR = 4;
      for (int i0 = tid; i0 < items; i0 += LT * UNR) {
        float4 s4[UNR];
        if (EPI == EPI_BWD_DATA && S) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * LT;
            if (i < items && mq < m4_valid)
              s4[u] = __ldg(reinterpret_cast<const float4*>(S + (long long)(n0 + (i >> 5)) * p.ldc + m0 + mq * 4));
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int i = i0 + u * LT;
          if (i < items && mq < m4_valid) {
            float4 o = *reinterpret_cast<const float4*>(T + (i >> 5) * TC_BM + mq * 4);
            if (EPI == EPI_FWD) {
              o.x = act_fwd(o.x + b4.x, p.act, p.slope); o.y = act_fwd(o.y + b4.y, p.act, p.slope);
              o.z = act_fwd(o.z + b4.z, p.act, p.slope); o.w = act_fwd(o.w + b4.w, p.act, p.slope);
            }
            if (EPI == EPI_BWD_DATA && S) {
              o.x *= act_bwd_from_out(s4[u].x, p.act, p.slope); o.y *= act_bwd_from_out(s4[u].y, p.act, p.slope);
              o.z *= act_bwd_from_out(s4[u].z, p.act, p.slope); o.w *= act_bwd_from_out(s4[u].w, p.act, p.slope);
            }
            *reinterpret_cast<float4*>(C + (long long)(n0 + (i >> 5)) * p.ldc + m0 + mq * 4) = o;
          }
        }
      }
    } else {
      for (int c = half; chunk_ok(c); c += LW / 4) {
        const int nb = n0 + c * 16;
        float v[16];
        tmem_chunk(c, v);
        if (!m_ok) continue;
        if (EPI == EPI_BWD_DATA && S) {
          // all 16 saved activations first: the stores below may alias them as far as the compiler knows,
          // and one dependent global load per store serialises the whole epilogue on memory latency
          float sv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            sv[j] = (nb + j < p.N) ? __ldg(S + (long long)(nb + j) * p.ldc + m) : 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= act_bwd_from_out(sv[j], p.act, p.slope);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (nb + j < p.N) {
            const long long off = (long long)(nb + j) * p.ldc + m;
            float o = v[j];
            if (EPI == EPI_FWD) o = act_fwd(o + bias, p.act, p.slope);
            C[off] = o;
          }
        }
      }
    }
  }

  if (warp == 0) TC_STAMP(6);
  tc_fence_before();
  __syncthreads();
  if (warp == MMAW) tmem_dealloc(tmem_d, tmem_cols);
  if (warp == 0) TC_STAMP(7);
}

// N tile: the largest balanced tile (multiple of 16, <= 256) whose TMEM plan keeps every hi*hi region
// at or below TC_MAX_ACCUM accumulations for this K.
static inline int tc_pick_bn(int N, int K) {
  const int nks = (K + 7) / 8;
  int need = (nks + TC_MAX_ACCUM - 1) / TC_MAX_ACCUM;
  if (need > TC_MAX_MAIN) need = TC_MAX_MAIN;
  for (int tiles = (N + TC_MAX_BN - 1) / TC_MAX_BN;; ++tiles) {
    const int per = (N + tiles - 1) / tiles;
    const int bn = (per + 15) / 16 * 16;
    if (tc_n_main(bn) >= need || bn <= 32) return bn;
  }
}
static inline size_t tc_stage_bytes(int bn) {
  const size_t bn_pad = ((size_t)bn + 31) & ~(size_t)31;
  return 2 * TC_BM * TC_BK * 4 + 2 * bn_pad * TC_BK * 4;
}
constexpr size_t TC_SMEM_BUDGET = 225 * 1024;  // of the 227 KB a CTA may use; static barriers take the rest
static inline int tc_pick_stages(int bn) {
  int n = (int)((TC_SMEM_BUDGET - 1024) / tc_stage_bytes(bn));
  return n > TC_MAX_STAGES ? TC_MAX_STAGES : (n < 2 ? 2 : n);
}

static inline int tc_pow2_cols(int cols) {
  int c = 32;
  while (c < cols) c <<= 1;
  return c;
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI, int NB, int OCC, int LW = 8, int BKT = 32>
static inline cudaError_t launch_tc_gemm_nb(const TcParams& p, int G, cudaStream_t stream) {
  static unsigned long long attr_set = 0;  // per template instantiation, one bit per device
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(tc_grouped_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI, NB, OCC, LW, BKT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BUDGET);
    if (e != cudaSuccess) return e;
  }
  const size_t smem = (size_t)p.n_stages * (tc_stage_bytes(p.bn) * BKT / TC_BK) + 1024;
  dim3 grid((p.N + p.bn - 1) / p.bn, (p.M + TC_BM - 1) / TC_BM, G);
  tc_grouped_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI, NB, OCC, LW, BKT><<<grid, LW * 32 + 32, smem, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

// CGL_TUNE=<bits> in the environment: experiment switches of the tcgen05 kernels (TcParams::tune)
static inline int tc_tune() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CGL_TUNE");
    v = e ? atoi(e) : TC_TUNE_DEFAULT;
  }
  return v;
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
static inline cudaError_t launch_tc_gemm(TcParams p, int G, cudaStream_t stream) {
  if (G <= 0 || p.M <= 0 || p.N <= 0) return cudaSuccess;
  p.tune = tc_tune();
  const int nks = (p.K + 7) / 8;
  if constexpr ((EPI == EPI_ADAM || EPI == EPI_STORE) && !A_KMAJOR && !B_KMAJOR) {
   if (nks <= TC_MAX_ACCUM) {
    // short-K weight gradient: its epilogue (24 B per parameter for Adam) is the HBM-bound part of a round.
    // 128-wide tiles, one stage (64 KB) and 256 TMEM columns -> two CTAs per SM overlap epilogue and main loop.
    const int tiles = (p.N + 127) / 128;
    p.bn = ((p.N + tiles - 1) / tiles + 15) / 16 * 16;
    p.n_main = 1;
    p.tmem_cols = tc_pow2_cols(2 * tc_region_stride(p.bn));
    if (p.tune & 64) {   // two 32 KB stages of 16 k-rows
      p.n_stages = TC_WG_STAGES;
      return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI, 4, 2, 8, 16>(p, G, stream);
    }
    p.n_stages = 1;
    return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI, 4, 2>(p, G, stream);
  }
  }
  if constexpr (EPI != EPI_ADAM && B_KMAJOR) {
    if (nks <= TC_MAX_ACCUM && (p.tune & 1024)) {
      // short-K forward / data gradient (the 128 x 256 layer of the 2DMG discriminator, the generator trunks): a tile is
      // 4-8 k-blocks of MMAs between ~3 us of start-up and ~4 us of epilogue, so one CTA per SM leaves the tensor core idle
      // most of the time. One accumulation region suffices (<= 43 accumulations): batch tiles of <= 128 lines, one 64 KB
      // stage and 256 TMEM columns -> two CTAs per SM, one's start-up / epilogue under the other's main loop.
      const int tiles = (p.N + 127) / 128;
      p.bn = ((p.N + tiles - 1) / tiles + 15) / 16 * 16;
      p.n_main = 1;
      p.tmem_cols = tc_pow2_cols(2 * tc_region_stride(p.bn));
      p.n_stages = 1;
      return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI, 4, 2>(p, G, stream);
    }
  }
  p.bn = tc_pick_bn(p.N, p.K);
  p.n_stages = tc_pick_stages(p.bn);
  p.n_main = tc_n_main(p.bn);
  p.tmem_cols = TC_TMEM_COLS;
  if (p.bn <= 128 && EPI == EPI_FWD && (p.tune & 32)) return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI_FWD, 4, 1, 16>(p, G, stream);
  if (p.bn <= 128 && !A_KMAJOR && B_KMAJOR && (p.tune & 128)) return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI, 4, 1, 16>(p, G, stream);
  if (p.bn <= 128) return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI, 4, 1>(p, G, stream);
  return launch_tc_gemm_nb<A_KMAJOR, B_KMAJOR, EPI, 8, 1>(p, G, stream);
}

// The tensor-core path needs float4-addressable operands: 16-byte aligned bases / strides, the
// contiguous extent a multiple of 4, and enough work per tile to be worth an MMA tile.
static inline bool tc_rowmap_ok(const RowMap& m, int contiguous_extent) { return m.vec && (contiguous_extent % 4 == 0); }

}  // namespace cgl
