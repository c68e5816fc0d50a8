// tc_sweep.cuh -- persistent forward kernel: the LEAN 16-warp loader loop of the one-tile kernel (tc_gemm.cuh) inside a
// tile loop, with a dedicated MMA warp and a dedicated epilogue warpgroup (tc_persist.cuh's roles).
//
// Why: in-kernel timeline + launch arithmetic of the one-tile forward kernel (profiles/fwd_sweep_r2.md), K = 1024 tile:
//   start-up 2.8 us (TMEM allocation, barrier set-up, first global loads) + main loop 29.5 us + epilogue 2.8 us + 4.4 us
//   between the exit of a CTA and the first instruction of the next one on that SM (198 KB of shared memory and 512 TMEM
//   columns per CTA: one CTA per SM, nothing overlaps) = 39.7 us per tile, 25 % of it not the main loop (K = 512: 23.3 us, 32 %).
// tc_persist.cuh removes that for the data-gradient product with 8 loader warps; for the forward product 8 loader warps
// are too few, and its flat k-block stream with 16 warps (tc_persistent_gemm_kernel<.., LW = 16>) costs 420 instructions per
// warp and k-block against 190 here -- issue-bound, 20 - 50 % slower than the one-tile kernel (measured). Here the loader
// loop is the one-tile kernel's, verbatim: everything that depends on the tile is set up once per tile, outside the k loop.
//   warps  0..15 : loaders (80 registers: 3 k-blocks of loads in flight). The load pipeline drains at a tile boundary; the
//                  MMA warp still has up to n_stages staged k-blocks to work on meanwhile.
//   warp   16    : one thread issues the MMAs (17..19 only fill the warpgroup; it keeps 32 registers)
//   warps 20..23 : epilogue, one warp per TMEM lane quarter (128 registers): accumulator -> registers, release TMEM
//                  (acc_empty), then bias / activation / derivative and the stores under the next tile's main loop.
#pragma once
#include "tc_persist.cuh"

namespace cgl {

constexpr int TCS_LW = 16;
constexpr int TCS_MMA_WARP = 16;
constexpr int TCS_EPI_WARP0 = 20;
constexpr int TCS_THREADS = 24 * 32;
#ifndef TCS_DEPTH
#define TCS_DEPTH 3
#endif

template <bool A_KMAJOR, bool B_KMAJOR, int EPI, int MAXCH>
__global__ void __launch_bounds__(TCS_THREADS, 1) tc_sweep_gemm_kernel(const TcParams p, const int G) {
  constexpr int LW = TCS_LW, LT = LW * 32, BKT = 32, KG = 8;
  constexpr int NA = 4 * KG / LW, NBW = 4 * KG / LW;     // 2 + 2 patches per loader warp (bn <= 128)
  extern __shared__ __align__(1024) char tc_smem[];
  __shared__ __align__(8) unsigned long long bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_acc_full;
  __shared__ __align__(8) unsigned long long bar_acc_empty;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bn = p.bn;
  const int nst = p.n_stages;
  const uint32_t a_bytes = TC_BM * BKT * 4;
  const int bn_pad = (bn + 31) & ~31;
  const uint32_t b_bytes = (uint32_t)bn_pad * BKT * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  char* smem = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  const int stride = tc_region_stride(bn);
  const int n_main = p.n_main;
  const int nkb = (p.K + BKT - 1) / BKT;
  const int nks = (p.K + 7) >> 3;

  // tiles: t -> (group, m tile, n tile), n fastest: CTAs that run side by side share a group's B operand in L2
  const int tiles_n = (p.N + bn - 1) / bn;
  const int tiles_m = (p.M + TC_BM - 1) / TC_BM;
  const int tiles_pg = tiles_m * tiles_n;
  const int total_tiles = G * tiles_pg;
  const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_coords = [&](int i, int& g, int& m0, int& n0) {
    const int t = (int)blockIdx.x + i * (int)gridDim.x;
    g = t / tiles_pg;
    const int r = t - g * tiles_pg;
    const int mt = r / tiles_n;
    m0 = mt * TC_BM;
    n0 = (r - mt * tiles_n) * bn;
  };

  if (tid == 0) {
    for (int i = 0; i < nst; ++i) {
      mbar_init(smem_u32(&bar_full[i]), LT);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    mbar_init(smem_u32(&bar_acc_full), 1);
    mbar_init(smem_u32(&bar_acc_empty), 32 * TCP_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TCS_MMA_WARP) tmem_alloc(smem_u32(&tmem_slot), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp < TCS_MMA_WARP) {
    // ===== loaders: per tile, the loop of tc_grouped_gemm_kernel =====
    constexpr int DEPTH = TCS_DEPTH;
    int st_s = 0;
    uint32_t st_par = 1;
    bool st_first = true;
    for (int i = 0; i < my_tiles; ++i) {
      int g, m0, n0;
      tile_coords(i, g, m0, n0);
      const Rows RA = resolve(p.A, g);
      const Rows RB = resolve(p.B, g);
      const int b_lines = (p.N - n0 < bn) ? (p.N - n0) : bn;
      const int npb = B_KMAJOR ? ((b_lines + 3) >> 2) : (bn_pad >> 5) * KG;
      float4 ra[DEPTH][NA], rb[DEPTH][NBW];
      auto load_block = [&](int kb, float4 (&qa)[NA], float4 (&qb)[NBW]) {
        const int k0 = kb * BKT;
#pragma unroll
        for (int u = 0; u < NA; ++u) qa[u] = tc_patch_load<A_KMAJOR, BKT>(RA, warp + LW * u, lane, m0, p.M, k0, p.K);
#pragma unroll
        for (int u = 0; u < NBW; ++u) {
          const int pp = warp + LW * u;
          qb[u] = (pp < npb) ? tc_patch_load<B_KMAJOR, BKT>(RB, pp, lane, n0, p.N, k0, p.K) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto store_block = [&](const float4 (&qa)[NA], const float4 (&qb)[NBW]) {
        if (!st_first) mbar_wait(smem_u32(&bar_empty[st_s]), st_par);
        char* a_hi = smem + (size_t)st_s * stage_bytes;
        char* a_lo = a_hi + a_bytes;
        char* b_hi = a_hi + 2 * a_bytes;
        char* b_lo = b_hi + b_bytes;
#pragma unroll
        for (int u = 0; u < NA; ++u) tc_split_store(a_hi, a_lo, tc_patch_offset<A_KMAJOR, BKT>(warp + LW * u, lane), qa[u]);
#pragma unroll
        for (int u = 0; u < NBW; ++u) {
          const int pp = warp + LW * u;
          if (pp < npb) tc_split_store(b_hi, b_lo, tc_patch_offset<B_KMAJOR, BKT>(pp, lane), qb[u]);
        }
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
    }
  } else if (warp < TCS_EPI_WARP0) {
    // ===== MMA issuer =====
    setmaxnreg_dec<32>();
    if (warp == TCS_MMA_WARP && lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(!A_KMAJOR, !B_KMAJOR, bn);
      const uint32_t a_lbo = A_KMAJOR ? 16u : 4096u, a_sbo = A_KMAJOR ? 1024u : 512u;
      const uint32_t b_lbo = B_KMAJOR ? 16u : 4096u, b_sbo = B_KMAJOR ? 1024u : 512u;
      const uint32_t a_step = A_KMAJOR ? 32u : 1024u, b_step = B_KMAJOR ? 32u : 1024u;
      const uint32_t a_lay = A_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      const uint32_t b_lay = B_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      int s = 0;
      uint32_t par = 0;
      for (int i = 0; i < my_tiles; ++i) {
        if (i > 0) {   // all four quarters of tile i-1 are out of TMEM
          mbar_wait(smem_u32(&bar_acc_empty), (uint32_t)((i - 1) & 1));
          tc_fence_after();
        }
        int ks = 0, reg = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&bar_full[s]), par);
          tc_fence_after();
          const uint32_t sa_hi = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t sa_lo = sa_hi + a_bytes, sb_hi = sa_hi + 2 * a_bytes, sb_lo = sb_hi + b_bytes;
#pragma unroll
          for (int j = 0; j < BKT / 8; ++j) {
            if (ks < nks) {
              const uint64_t dah = umma_desc(sa_hi + j * a_step, a_lbo, a_sbo, a_lay);
              const uint64_t dal = umma_desc(sa_lo + j * a_step, a_lbo, a_sbo, a_lay);
              const uint64_t dbh = umma_desc(sb_hi + j * b_step, b_lbo, b_sbo, b_lay);
              const uint64_t dbl = umma_desc(sb_lo + j * b_step, b_lbo, b_sbo, b_lay);
              const uint32_t main_col = (uint32_t)((1 + reg) * stride);
              if (++reg == n_main) reg = 0;
              umma_tf32(tmem_d, dal, dbh, idesc, ks > 0 ? 1u : 0u);
              umma_tf32(tmem_d, dah, dbl, idesc, 1u);
              umma_tf32(tmem_d + main_col, dah, dbh, idesc, ks >= n_main ? 1u : 0u);
              ++ks;
            }
          }
          umma_commit(smem_u32(&bar_empty[s]));
          if (++s == nst) { s = 0; par ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full));
      }
    }
  } else {
    // ===== epilogue warps, one per TMEM lane quarter =====
    setmaxnreg_inc<128>();
    const int q = warp & 3;
    const int nch = bn >> 4;
    const int n_used = nks < n_main ? nks : n_main;
    for (int i = 0; i < my_tiles; ++i) {
      int g, m0, n0;
      tile_coords(i, g, m0, n0);
      const int m = m0 + q * 32 + lane;
      const bool m_ok = m < p.M;
      const int rowid = p.cidx ? p.cidx[g] : g;
      float bias = 0.f;
      if (EPI == EPI_FWD && p.bias_base && m_ok) {
        const int brow = p.bias_idx ? p.bias_idx[g] : g;
        bias = __ldg(p.bias_base + (long long)brow * p.bias_gstride + p.bias_off + m);
      }
      mbar_wait_relaxed(smem_u32(&bar_acc_full), (uint32_t)(i & 1));
      tc_fence_after();
      float v[MAXCH][16];
#pragma unroll
      for (int c = 0; c < MAXCH; ++c) {
        if (c < nch) {   // warp-uniform
          const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
          // regions in the order of tc_grouped_gemm_kernel (main 1, main 2, ..., corrections last): the same bits
          uint32_t r0[16], t[16];
          tmem_ld16_async(taddr + (uint32_t)stride, r0);
          if (n_used >= 2) tmem_ld16_async(taddr + (uint32_t)(2 * stride), t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[c][j] = (n_used >= 2) ? __uint_as_float(r0[j]) + __uint_as_float(t[j]) : __uint_as_float(r0[j]);
          for (int r = 3; r <= n_used; ++r) {
            tmem_ld16_async(taddr + (uint32_t)(r * stride), t);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[c][j] += __uint_as_float(t[j]);
          }
          tmem_ld16_async(taddr, t);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[c][j] += __uint_as_float(t[j]);
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_acc_empty));   // the accumulator is in registers: the next tile may overwrite TMEM

      float* C = p.cbase + (long long)rowid * p.c_gstride + p.c_off;
      const float* S = (EPI == EPI_BWD_DATA && p.saved) ? p.saved + (long long)g * p.saved_gstride : nullptr;
#pragma unroll
      for (int c = 0; c < MAXCH; ++c) {
        const int nb = n0 + c * 16;
        if (c < nch && nb < p.N && m_ok) {
          if (EPI == EPI_BWD_DATA && S) {
            float sv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = (nb + j < p.N) ? __ldg(S + (long long)(nb + j) * p.ldc + m) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[c][j] *= act_bwd_from_out(sv[j], p.act, p.slope);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (nb + j < p.N) {
              float o = v[c][j];
              if (EPI == EPI_FWD) o = act_fwd(o + bias, p.act, p.slope);
              C[(long long)(nb + j) * p.ldc + m] = o;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TCS_MMA_WARP) tmem_dealloc(tmem_d, TC_TMEM_COLS);
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI, int MAXCH>
static inline void launch_tc_sweep_inst(const TcParams& p, int G, int grid, size_t smem, cudaStream_t stream, cudaError_t* err) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    *err = cudaFuncSetAttribute(tc_sweep_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI, MAXCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)TC_SMEM_BUDGET);
    if (*err != cudaSuccess) return;
  }
  tc_sweep_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI, MAXCH><<<grid, TCS_THREADS, smem, stream>>>(p, G);
  count_launch();
  *err = cudaGetLastError();
}

// true: launched (or failed with *err set); false: not applicable
template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
static inline bool launch_tc_sweep(TcParams p, int G, cudaStream_t stream, cudaError_t* err) {
  if (!(tc_tune() & (A_KMAJOR ? 16384 : 32768))) return false;
  if ((tc_tune() & 1024) && (p.K + 7) / 8 <= TC_MAX_ACCUM) return false;   // short K: two CTAs per SM (tc_gemm.cuh)
  p.bn = tc_pick_bn(p.N, p.K);
  if (p.bn > 128 || p.K < 4 * TC_BK) return false;
  p.n_stages = tc_pick_stages(p.bn);
  p.n_main = tc_n_main(p.bn);
  p.tmem_cols = TC_TMEM_COLS;
  p.tune = tc_tune();
  const size_t smem = (size_t)p.n_stages * tc_stage_bytes(p.bn) + 1024;
  const long long tiles = (long long)G * ((p.M + TC_BM - 1) / TC_BM) * ((p.N + p.bn - 1) / p.bn);
  const int sms = tc_num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);
  *err = cudaSuccess;
  if (p.bn <= 112) launch_tc_sweep_inst<A_KMAJOR, B_KMAJOR, EPI, 7>(p, G, grid, smem, stream, err);
  else launch_tc_sweep_inst<A_KMAJOR, B_KMAJOR, EPI, 8>(p, G, grid, smem, stream, err);
  return true;
}

}  // namespace cgl
