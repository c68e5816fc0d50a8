// tc_pair.cuh -- the tcgen05 grouped GEMM of tc_gemm.cuh on CTA PAIRS (tcgen05.mma.cta_group::2): two CTAs of a cluster,
// on the two SMs of a TPC, compute a 256 x bn output tile of one group together. Each CTA stages and owns its own 128 rows
// of the A operand (the weights) and its own 128 TMEM lanes of the accumulator, exactly as before; the B operand (the
// batch: bn <= 128 lines of activations, K-major) is staged ONCE PER PAIR -- each CTA splits and stores HALF of its lines,
// and the tensor cores of both SMs read both halves.
//
// Why (profiles/ncu_fwd_r1.md): the one-CTA kernel is bound by the L1TEX data pipe. Per 32-wide k-block a CTA moved 30 KB
// of global loads, 60 KB of hi/lo STS and 90 KB of tensor-core operand reads over the same 128 B/clk path against 672 clk
// of MMA time. In a pair each CTA loads 23 KB, stores 46 KB, and its tensor core reads 12 x (4 KB of A + 1.75 KB of its
// half of B) = 69 KB: 138 KB instead of 180 KB per k-block, and a quarter fewer loader instructions (the loaders are also
// bound by their own instruction stream). The stage shrinks to 46 KB, so four stages fit instead of three.
//
// Protocol (one leader: cluster rank 0):
//   loaders (16 warps in EACH CTA): LDG -> split -> STS into the CTA's own stage s -> fence.proxy.async -> arrive on the
//       CTA's OWN full[s] (CTA scope, as in the one-CTA kernel: a cluster-scope release per loader warp showed up as a
//       `membar` stall of 5 warps per issue slot and made the pair kernel 50 % slower than the one-CTA kernel)
//   forwarder (one thread of the peer CTA's otherwise idle MMA warp): waits the peer's full[s], then ONE remote arrive
//       (release.cluster) on the leader's peer_full[s]
//   MMA thread (leader only): waits its own full[s] and peer_full[s] (acquire.cluster), issues the 3 split products per k-step with
//       cta_group::2 (A / B descriptors address the same shared-memory offsets in both CTAs, D the same TMEM columns),
//       tcgen05.commit.cta_group::2 multicast -> empty[s] of BOTH CTAs; after the last k-block -> done of both CTAs
//   epilogue (the loader warps of each CTA): its own 128 accumulator rows, as in tc_grouped_gemm_kernel.
//   TMEM is allocated / freed with cta_group::2 by the same warp of both CTAs; cluster barriers fence set-up and exit.
#pragma once
#include <cooperative_groups.h>

#include "tc_gemm.cuh"

namespace cgl {

constexpr int TCP2_LW = 16;   // loader warps per CTA
#ifndef TCP2_DEPTH
#define TCP2_DEPTH 3          // k-blocks of global loads in flight per loader thread (12 registers each)
#endif

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory variable in the CTA of rank `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(TC_WAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait_cluster(bar, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) __trap();   // 2 s: a lost arrival must fail loudly, not hang the GPU
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {  // one full warp of EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once every MMA issued so far is done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA), N = the full tile
__host__ __device__ inline uint32_t umma_idesc_tf32_pair(bool a_mn_major, bool b_mn_major, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// B is K-major (lines = batch rows). EPI: EPI_FWD / EPI_BWD_DATA / EPI_STORE (the Adam epilogue stays on the OCC = 2 kernel).
template <bool A_KMAJOR, int EPI>
__global__ void __launch_bounds__(TCP2_LW * 32 + 32, 1) tc_pair_gemm_kernel(const TcParams p) {
  constexpr int LW = TCP2_LW;
  constexpr int LT = LW * 32;
  constexpr int MMAW = LW;
  constexpr int BKT = 32, KG = 8;
  constexpr int NA = 4 * KG / LW;      // 2 A patches per loader warp
  extern __shared__ __align__(1024) char tc_smem[];
  __shared__ __align__(8) unsigned long long bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_peer[TC_MAX_STAGES];   // leader: the peer's stage s is full
  __shared__ __align__(8) unsigned long long bar_done;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int g = blockIdx.z;
  // the pair lies along x (a kernel that uses cta_group::2 is only accepted with an even cluster width in x)
  const int m0 = blockIdx.x * TC_BM;           // this CTA's 128 rows (may lie beyond M for the odd last pair: zeros)
  const int n0 = blockIdx.y * p.bn;
  const int bn = p.bn;
  const int bh = bn >> 1;                      // lines of B this CTA stages (multiple of 8)
  const int nb0 = n0 + (int)rank * bh;         // first of them
  const int nst = p.n_stages;

  const uint32_t a_bytes = TC_BM * BKT * 4;
  const uint32_t b_bytes = (uint32_t)bh * BKT * 4;          // multiple of 1024
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  char* smem = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);

  const int stride = tc_region_stride(bn);
  const int n_main = p.n_main;
  const uint32_t tmem_cols = (uint32_t)p.tmem_cols;
  const int nkb = (p.K + BKT - 1) / BKT;
  const int nks = (p.K + 7) >> 3;

  if (tid == 0) {
    for (int i = 0; i < nst; ++i) {
      mbar_init(smem_u32(&bar_full[i]), LT);       // every loader thread of THIS CTA
      mbar_init(smem_u32(&bar_empty[i]), 1);
      mbar_init(smem_u32(&bar_peer[i]), 1);        // the peer's forwarder (used on the leader only)
    }
    mbar_init(smem_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMAW) tmem_alloc_pair(smem_u32(&tmem_slot), tmem_cols);
  tc_fence_before();
  cluster_sync_all();          // barriers initialised and TMEM allocated in both CTAs before anything is signalled
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp == MMAW) {
    // ===== MMA issuer: one thread of the leader CTA =====
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc_tf32_pair(!A_KMAJOR, false, bn);
      const uint32_t a_lbo = A_KMAJOR ? 16u : (uint32_t)(KG * 512), a_sbo = A_KMAJOR ? 1024u : 512u;
      const uint32_t a_step = A_KMAJOR ? 32u : 1024u;
      const uint32_t a_lay = A_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      int ks = 0, s = 0, reg = 0;
      uint32_t par = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&bar_full[s]), par);
        mbar_wait_cluster(smem_u32(&bar_peer[s]), par);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t sa_lo = sa_hi + a_bytes, sb_hi = sa_hi + 2 * a_bytes, sb_lo = sb_hi + b_bytes;
#pragma unroll
        for (int j = 0; j < BKT / 8; ++j) {
          if (ks < nks) {
            const uint64_t dah = umma_desc(sa_hi + j * a_step, a_lbo, a_sbo, a_lay);
            const uint64_t dal = umma_desc(sa_lo + j * a_step, a_lbo, a_sbo, a_lay);
            const uint64_t dbh = umma_desc(sb_hi + j * 32u, 16u, 1024u, UMMA_LAYOUT_SW128);
            const uint64_t dbl = umma_desc(sb_lo + j * 32u, 16u, 1024u, UMMA_LAYOUT_SW128);
            const uint32_t main_col = (uint32_t)((1 + reg) * stride);
            if (++reg == n_main) reg = 0;
            umma_tf32_pair(tmem_d, dal, dbh, idesc, ks > 0 ? 1u : 0u);
            umma_tf32_pair(tmem_d, dah, dbl, idesc, 1u);
            umma_tf32_pair(tmem_d + main_col, dah, dbh, idesc, ks >= n_main ? 1u : 0u);
            ++ks;
          }
        }
        umma_commit_pair(smem_u32(&bar_empty[s]));
        if (++s == nst) { s = 0; par ^= 1u; }
      }
      umma_commit_pair(smem_u32(&bar_done));
    } else if (rank == 1 && lane == 0) {
      // ===== forwarder: this CTA's stage s is full -> one cluster-scope arrive on the leader =====
      const uint32_t peer0 = mapa_shared(smem_u32(&bar_peer[0]), 0);
      int s = 0;
      uint32_t par = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&bar_full[s]), par);
        mbar_arrive_cluster(peer0 + 8u * (uint32_t)s);
        if (++s == nst) { s = 0; par ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int half = warp >> 2;
    const int rowid = p.cidx ? p.cidx[g] : g;

    // ===== loader warps =====
    const Rows RA = resolve(p.A, g);
    const Rows RB = resolve(p.B, g);
    // this CTA's half of the B tile: only the lines that exist are loaded and stored (see tc_grouped_gemm_kernel)
    int b_lines = p.N - nb0;
    b_lines = b_lines < 0 ? 0 : (b_lines > bh ? bh : b_lines);
    const int npb = (b_lines + 3) >> 2;          // <= 16 = one patch per loader warp
    const bool b_mine = warp < npb;
    constexpr int DEPTH = TCP2_DEPTH;
    float4 ra[DEPTH][NA], rb[DEPTH];
    auto load_block = [&](int kb, float4 (&qa)[NA], float4& qb) {
      const int k0 = kb * BKT;
#pragma unroll
      for (int i = 0; i < NA; ++i) qa[i] = tc_patch_load<A_KMAJOR, BKT>(RA, warp + LW * i, lane, m0, p.M, k0, p.K);
      qb = b_mine ? tc_patch_load<true, BKT>(RB, warp, lane, nb0, p.N, k0, p.K) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    int st_s = 0;
    uint32_t st_par = 1;
    bool st_first = true;
    auto store_block = [&](const float4 (&qa)[NA], const float4& qb) {
      if (!st_first) mbar_wait(smem_u32(&bar_empty[st_s]), st_par);   // the pair's MMAs that read this stage are done
      char* a_hi = smem + (size_t)st_s * stage_bytes;
      char* a_lo = a_hi + a_bytes;
      char* b_hi = a_hi + 2 * a_bytes;
      char* b_lo = b_hi + b_bytes;
#pragma unroll
      for (int i = 0; i < NA; ++i) tc_split_store(a_hi, a_lo, tc_patch_offset<A_KMAJOR, BKT>(warp + LW * i, lane), qa[i]);
      if (b_mine) tc_split_store(b_hi, b_lo, tc_patch_offset<true, BKT>(warp, lane), qb);
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
    for (int kb0 = 0; kb0 < nkb; kb0 += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int kb = kb0 + d;
        if (kb < nkb) {
          store_block(ra[d], rb[d]);
          if (kb + DEPTH < nkb) load_block(kb + DEPTH, ra[d], rb[d]);
        }
      }
    }

    // ===== epilogue: this CTA's 128 rows =====
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    float* C = p.cbase + (long long)rowid * p.c_gstride + p.c_off;
    const int n_used = nks < n_main ? nks : n_main;
    const float* S = (EPI == EPI_BWD_DATA && p.saved) ? p.saved + (long long)g * p.saved_gstride : nullptr;
    const int nch = bn >> 4;
    auto tmem_chunk = [&](int c, float (&v)[16]) {
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
      uint32_t r0[16], r1[16], r2[16], r3[16];
      tmem_ld16_async(taddr + (uint32_t)stride, r0);
      tmem_ld16_async(taddr, r1);
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
      for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r1[j]);
    };
    auto chunk_ok = [&](int c) { return c < nch && n0 + c * 16 < p.N; };
    // accumulators to shared ([n][m], through the operand stages: every MMA of the pair has completed), then float4 rows
    float* T = reinterpret_cast<float*>(smem);
    for (int c = half; chunk_ok(c); c += LW / 4) {
      float gv[16];
      tmem_chunk(c, gv);
#pragma unroll
      for (int j = 0; j < 16; ++j) T[(c * 16 + j) * TC_BM + q * 32 + lane] = gv[j];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(LT) : "memory");
    const int n_valid = (p.N - n0 < bn) ? (p.N - n0) : bn;
    int m_rows = p.M - m0;
    m_rows = m_rows < 0 ? 0 : (m_rows > TC_BM ? TC_BM : m_rows);
    const int m4_valid = m_rows >> 2;                  // M % 4 == 0 (c_vec)
    const int items = n_valid * (TC_BM / 4);
    const int mq = tid & 31;
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
  }

  tc_fence_before();
  cluster_sync_all();          // neither CTA leaves (or frees TMEM) while the other may still read its operands / be signalled
  if (warp == MMAW) tmem_dealloc_pair(tmem_d, tmem_cols);
}

// true: launched (or failed with *err set); false: not applicable (the caller uses the one-CTA kernels)
template <bool A_KMAJOR, int EPI>
static inline bool launch_tc_pair(TcParams p, int G, cudaStream_t stream, cudaError_t* err) {
  if (G <= 0 || p.M <= 0 || p.N <= 0) return false;
  if (!p.c_vec || p.M <= TC_BM) return false;              // float4 epilogue only; a single M tile has nothing to pair
  // an odd number of M tiles would add a CTA that stages its half of B for nothing (out = 784: 8 CTAs for 6.1 tiles of
  // rows; measured 4 - 25 % slower than the one-CTA kernels there), an even number is never slower (profiles/fwd_variants_r2.md)
  if ((((p.M + TC_BM - 1) / TC_BM) & 1) && !(tc_tune() & 65536)) return false;
  p.bn = tc_pick_bn(p.N, p.K);
  if (p.bn > 128 || p.K < 2 * TC_BK) return false;
  p.n_main = tc_n_main(p.bn);
  p.tmem_cols = TC_TMEM_COLS;
  p.tune = tc_tune();
  const size_t stage = 2 * (size_t)TC_BM * TC_BK * 4 + 2 * (size_t)(p.bn / 2) * TC_BK * 4;
  int nst = (int)((TC_SMEM_BUDGET - 1024) / stage);
  p.n_stages = nst > TC_MAX_STAGES ? TC_MAX_STAGES : nst;
  size_t smem = (size_t)p.n_stages * stage;
  const size_t t_bytes = (size_t)p.bn * TC_BM * 4;          // the epilogue's transpose buffer lives in the stages
  if (smem < t_bytes) smem = t_bytes;
  smem += 1024;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    *err = cudaFuncSetAttribute(tc_pair_gemm_kernel<A_KMAJOR, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)TC_SMEM_BUDGET);
    if (*err != cudaSuccess) return true;
  }
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((m_tiles + 1) / 2 * 2, (p.N + p.bn - 1) / p.bn, G);
  cfg.blockDim = dim3(TCP2_LW * 32 + 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  *err = cudaLaunchKernelEx(&cfg, tc_pair_gemm_kernel<A_KMAJOR, EPI>, p);
  if (*err != cudaSuccess && getenv("CGL_DEBUG_LAUNCH")) {
    int nclusters = -1;
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&nclusters, tc_pair_gemm_kernel<A_KMAJOR, EPI>, &cfg);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, tc_pair_gemm_kernel<A_KMAJOR, EPI>);
    fprintf(stderr, "tc_pair launch failed: %s; grid (%u,%u,%u) block %u smem %zu; max active clusters %d (%s); regs %d static smem %zu maxdyn %d\n",
            cudaGetErrorString(*err), cfg.gridDim.x, cfg.gridDim.y, cfg.gridDim.z, cfg.blockDim.x, smem, nclusters,
            cudaGetErrorString(e2), fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes);
  }
  count_launch();
  if (*err == cudaSuccess) *err = cudaGetLastError();
  return true;
}

}  // namespace cgl
