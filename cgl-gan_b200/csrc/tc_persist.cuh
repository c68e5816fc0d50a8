// tc_persist.cuh -- persistent, warp-specialised variant of the tcgen05 grouped GEMM (tc_gemm.cuh) for the
// forward and data-gradient products (long K: 256..1024, the batch as N <= 128).
//
// Why: tc_grouped_gemm_kernel runs one output tile per CTA and one CTA per SM (its TMEM plan takes all 512
// columns, its stages ~180 KB), so every tile pays its own start-up (first global loads: ~3 us) and its own
// epilogue (~4 us) with the tensor core and the loaders idle -- 16 % of a K = 1024 tile, 30 % of a K = 512 tile
// (in-kernel timeline, profiles/tc_timeline.py). Here one CTA per SM walks over its tiles:
//   warps 0..7  : loaders. ONE flat stream of k-blocks over all the CTA's tiles (global -> registers, DEPTH
//                 k-blocks in flight -> hi/lo split -> shared stage), so the loads of tile i+1 are in flight and its
//                 first stages are full while tile i is still in its MMA tail and epilogue
//   warp 8      : one thread issues the MMAs of a tile (before the first one it waits for acc_empty), then the whole
//                 warp is the epilogue of TMEM lane quarter 0
//   warps 9..11 : epilogue of lane quarters 1..3. Wait acc_full, pull the quarter of the accumulator tile out of
//                 TMEM into registers (summing the accumulation regions), arrive on acc_empty -- the tensor core is
//                 released after ~1 us -- and only then apply bias / activation / derivative and store (lane =
//                 contiguous output index: each warp store is one full 128-byte line), under the next main loop.
// 12 warps: registers are handed out per 4 warps, so 384 threads keep 168 registers per thread (the loaders hold 3
// k-blocks = 96 registers of loads in flight, the epilogue a 32 x 112 accumulator slice = 112).
// Operand staging, shared-memory layouts, the 3xTF32 split and the TMEM region plan are those of tc_gemm.cuh.
#pragma once
#include "tc_gemm.cuh"

namespace cgl {

// Variants (bring-up; profiles/ compares them):
//   TCP_VARIANT 0: 13 warps (dedicated MMA warp), 128 registers each (registers are handed out per 4 warps), DEPTH 2
//   (12 warps with warp 8 issuing the MMAs AND acting as the epilogue of lane quarter 0 measured 2x slower: the
//    MMA-issuing lane then runs inside a long-lived divergent warp.)
//   TCP_VARIANT 2: 16 warps in 4 warpgroups with setmaxnreg: loaders 160 registers (DEPTH 3), the MMA warpgroup 40,
//                  the epilogue warpgroup 152
#ifndef TCP_VARIANT
#define TCP_VARIANT 0
#endif
constexpr int TCP_EPI_WARPS = 4;
constexpr int TCP_EPI_WARP0 = (TCP_VARIANT == 2) ? 12 : 9;   // first epilogue warp
constexpr int TCP_THREADS = (TCP_VARIANT == 2) ? 512 : 416;
constexpr int TCP_DEPTH = (TCP_VARIANT == 0) ? 2 : 3;
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// MAXCH: 16-column chunks of the N tile the epilogue keeps in registers (7: bn <= 112, 8: bn <= 128)
// LW: loader warps. 8: the layout above (TCP_VARIANT). 16: 24 warps in six warpgroups -- loaders 0..15, the MMA warp 16
// (17..19 only fill its warpgroup), epilogue 20..23 -- 768 threads start with 80 registers each; the MMA warpgroup gives
// up all but 32, the epilogue warpgroup takes 128 (it holds a 32 x 112 accumulator slice), the loaders keep 80 (two
// k-blocks = 32 registers of loads in flight). Used for the FORWARD product: 16 loader warps were what made the one-tile
// forward kernel fast (tc_gemm.cuh), and a tile there pays ~2.8 us of start-up, ~2.8 us of epilogue and 2.4 - 4.4 us of
// CTA turn-around per 20 - 35 us (in-kernel timeline + launch arithmetic, profiles/fwd_sweep_r2.md).
template <int LW>
struct TcpLayout {
  static constexpr int MMA_WARP = LW;
  static constexpr int EPI_WARP0 = (LW == 16) ? 20 : TCP_EPI_WARP0;
  static constexpr int THREADS = (LW == 16) ? 768 : TCP_THREADS;
  static constexpr int DEPTH = (LW == 16) ? 2 : TCP_DEPTH;
};
template <bool A_KMAJOR, bool B_KMAJOR, int EPI, int MAXCH, int LW = 8>
__global__ void __launch_bounds__(TcpLayout<LW>::THREADS, 1) tc_persistent_gemm_kernel(const TcParams p, const int G) {
  constexpr int MMAW = TcpLayout<LW>::MMA_WARP;
  constexpr int EPIW0 = TcpLayout<LW>::EPI_WARP0;
  constexpr int NPA = 32 / LW;        // A patches per loader warp (a 128-line tile has 32 patches of 4 lines / 4 k-rows)
  extern __shared__ __align__(1024) char tc_smem[];
  __shared__ __align__(8) unsigned long long bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_acc_full;
  __shared__ __align__(8) unsigned long long bar_acc_empty;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bn = p.bn;
  const int nst = p.n_stages;
  constexpr int NB = 32 / LW;         // B patches per loader warp (bn <= 128)

  const uint32_t a_bytes = TC_BM * TC_BK * 4;
  const int bn_pad = (bn + 31) & ~31;
  const uint32_t b_bytes = (uint32_t)bn_pad * TC_BK * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  char* smem = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);

  const int stride = tc_region_stride(bn);
  const int n_main = p.n_main;
  const int nkb = (p.K + TC_BK - 1) / TC_BK;
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
      mbar_init(smem_u32(&bar_full[i]), LW * 32);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    mbar_init(smem_u32(&bar_acc_full), 1);
    mbar_init(smem_u32(&bar_acc_empty), 32 * TCP_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMAW) tmem_alloc(smem_u32(&tmem_slot), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  if (warp < MMAW) {
    if (LW == 8 && TCP_VARIANT == 2) setmaxnreg_inc<160>();
    // ===== loaders: one flat stream of k-blocks over all tiles of this CTA =====
    const int npb = bn_pad >> 2;
    constexpr int DEPTH = TcpLayout<LW>::DEPTH;
    float4 ra[DEPTH][NPA], rb[DEPTH][NB];
    const int total = my_tiles * nkb;
    // Per-tile addressing state of the load stream, rebuilt only when the stream enters a new tile (the compiler
    // hoists this by itself in the one-tile-per-CTA kernel; here the tile changes inside the loop):
    //   K-major operand : pointer to (line t_u, k = 4*(lane%8)) per patch, NULL outside the matrix; a k-block adds k0
    //   MN-major operand: column offset t_u per patch (-1 outside); a k-block resolves ONE row pointer (all patches
    //                     of a thread share k = k0 + 4*(warp%8) + lane/8) and adds the offsets
    Rows RA = {}, RB = {};
    const float* pa[NPA];
    const float* pb[NB];
    int ta[NPA], tb[NB];
    const int kq = 4 * (lane & 7);                 // K-major: k offset inside a k-block
    const int kr = 4 * (warp & 7) + (lane >> 3);   // MN-major: k row inside a k-block
    auto setup_tile = [&](int i) {
      int g, m0l, n0l;
      tile_coords(i, g, m0l, n0l);
      RA = resolve(p.A, g);
      RB = resolve(p.B, g);
#pragma unroll
      for (int u = 0; u < NPA; ++u) {
        const int pp = warp + LW * u;
        if (A_KMAJOR) {
          const int t = m0l + 4 * pp + (lane >> 3);
          pa[u] = (t < p.M) ? row_ptr(RA, t) + kq : nullptr;
        } else {
          const int t = m0l + 32 * (pp >> 3) + 4 * (lane & 7);
          ta[u] = (t < p.M) ? t : -1;
        }
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int pp = warp + LW * u;
        if (B_KMAJOR) {
          const int t = n0l + 4 * pp + (lane >> 3);
          pb[u] = (pp < npb && t < p.N) ? row_ptr(RB, t) + kq : nullptr;
        } else {
          const int t = n0l + 32 * (pp >> 3) + 4 * (lane & 7);
          tb[u] = (pp < npb && t < p.N) ? t : -1;
        }
      }
    };
    int ld_kb = 0, ld_tile = 0;
    auto load_block = [&](int j, float4 (&qa)[NPA], float4 (&qb)[NB]) {
      (void)j;                      // loads are issued in increasing j: counters instead of a division per k-block
      const int kb = ld_kb;
      if (kb == 0) setup_tile(ld_tile);
      if (++ld_kb == nkb) { ld_kb = 0; ++ld_tile; }
      const int k0 = kb * TC_BK;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_KMAJOR) {
        const bool kok = k0 + kq < p.K;
#pragma unroll
        for (int u = 0; u < NPA; ++u) qa[u] = (kok && pa[u]) ? __ldg(reinterpret_cast<const float4*>(pa[u] + k0)) : zero;
      } else {
        const int k = k0 + kr;
        const float* rp = (k < p.K) ? row_ptr(RA, k) : nullptr;
#pragma unroll
        for (int u = 0; u < NPA; ++u) qa[u] = (rp && ta[u] >= 0) ? __ldg(reinterpret_cast<const float4*>(rp + ta[u])) : zero;
      }
      if (B_KMAJOR) {
        const bool kok = k0 + kq < p.K;
#pragma unroll
        for (int u = 0; u < NB; ++u) qb[u] = (kok && pb[u]) ? __ldg(reinterpret_cast<const float4*>(pb[u] + k0)) : zero;
      } else {
        const int k = k0 + kr;
        const float* rp = (k < p.K) ? row_ptr(RB, k) : nullptr;
#pragma unroll
        for (int u = 0; u < NB; ++u) qb[u] = (rp && tb[u] >= 0) ? __ldg(reinterpret_cast<const float4*>(rp + tb[u])) : zero;
      }
    };
    int st_s = 0;
    uint32_t st_par = 1;
    bool st_first = true;
    int st_kb = 0, st_tile = 0;
    int npb_st = npb;   // B patches of the tile being STORED that hold data (K-major B: lines beyond N are skipped, see tc_gemm.cuh)
    auto store_tile_setup = [&](int i) {
      if (!B_KMAJOR) return;
      int g, m0s, n0s;
      tile_coords(i, g, m0s, n0s);
      const int lines = (p.N - n0s < bn) ? (p.N - n0s) : bn;
      npb_st = (lines + 3) >> 2;
    };
    auto store_block = [&](const float4 (&qa)[NPA], const float4 (&qb)[NB]) {
      if (!st_first) mbar_wait(smem_u32(&bar_empty[st_s]), st_par);
      char* a_hi = smem + (size_t)st_s * stage_bytes;
      char* a_lo = a_hi + a_bytes;
      char* b_hi = a_hi + 2 * a_bytes;
      char* b_lo = b_hi + b_bytes;
#pragma unroll
      for (int u = 0; u < NPA; ++u) tc_split_store(a_hi, a_lo, tc_patch_offset<A_KMAJOR>(warp + LW * u, lane), qa[u]);
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int pp = warp + LW * u;
        if (pp < npb_st) tc_split_store(b_hi, b_lo, tc_patch_offset<B_KMAJOR>(pp, lane), qb[u]);
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
      if (d < total) load_block(d, ra[d], rb[d]);
    for (int j0 = 0; j0 < total; j0 += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const int j = j0 + d;
        if (j < total) {
          if (st_kb == 0) store_tile_setup(st_tile);      // (no division in this loop: it is issue-bound)
          if (++st_kb == nkb) { st_kb = 0; ++st_tile; }
          store_block(ra[d], rb[d]);
          if (j + DEPTH < total) load_block(j + DEPTH, ra[d], rb[d]);
        }
      }
    }
  } else if (warp < EPIW0) {
    // ===== MMA issuer: one thread of warp MMAW (the rest of its warpgroup, if any, only gives its registers away) =====
    if (LW == 16) setmaxnreg_dec<32>();
    else if (TCP_VARIANT == 2) setmaxnreg_dec<40>();
    if (warp == MMAW) {
      const uint32_t idesc = umma_idesc_tf32(!A_KMAJOR, !B_KMAJOR, bn);
      const uint32_t a_lbo = A_KMAJOR ? 16u : 4096u, a_sbo = A_KMAJOR ? 1024u : 512u;
      const uint32_t b_lbo = B_KMAJOR ? 16u : 4096u, b_sbo = B_KMAJOR ? 1024u : 512u;
      const uint32_t a_step = A_KMAJOR ? 32u : 1024u, b_step = B_KMAJOR ? 32u : 1024u;
      const uint32_t a_lay = A_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      const uint32_t b_lay = B_KMAJOR ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW128_BASE32B;
      int s = 0;
      uint32_t par = 0;
      for (int i = 0; i < my_tiles; ++i) {
        if (lane == 0) {
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
            for (int j = 0; j < TC_BK / 8; ++j) {
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
    }
  } else {
    // ===== epilogue warps, one per TMEM lane quarter =====
    if (LW == 16) setmaxnreg_inc<128>();
    else if (TCP_VARIANT == 2) setmaxnreg_inc<152>();
    const int q = warp & 3;   // the TMEM lane quarter a warp may read is fixed by its index
    const int nch = bn >> 4;
    const int n_used = nks < n_main ? nks : n_main;
    auto epilogue_tile = [&](int i) {
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
          // regions in the order of tc_grouped_gemm_kernel (main 1, main 2, ..., corrections last): same bits
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
    };
    for (int i = 0; i < my_tiles; ++i) epilogue_tile(i);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMAW) tmem_dealloc(tmem_d, TC_TMEM_COLS);
}

static inline int tc_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// true: launched (or failed with *err set); false: not applicable (the caller uses tc_grouped_gemm_kernel)
template <bool A_KMAJOR, bool B_KMAJOR, int EPI, int MAXCH, int LW>
static inline void launch_tc_persistent_inst(const TcParams& p, int G, int grid, size_t smem, cudaStream_t stream, cudaError_t* err) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    *err = cudaFuncSetAttribute(tc_persistent_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI, MAXCH, LW>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BUDGET);
    if (*err != cudaSuccess) return;
  }
  tc_persistent_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI, MAXCH, LW><<<grid, TcpLayout<LW>::THREADS, smem, stream>>>(p, G);
  count_launch();
  *err = cudaGetLastError();
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
static inline bool launch_tc_persistent(TcParams p, int G, cudaStream_t stream, cudaError_t* err) {
  // measured (profiles/linear_bench.py): the data-gradient product gains 10-20 % from the persistent kernel with 8 loader
  // warps; the forward product loses with 8 (its loaders are bound by their own instruction stream) and runs the
  // 16-loader-warp layout (tune bit 16)
  if (!(tc_tune() & (A_KMAJOR ? 16 : 8))) return false;
  if ((tc_tune() & 1024) && (p.K + 7) / 8 <= TC_MAX_ACCUM) return false;   // short K: the two-CTAs-per-SM variant (tc_gemm.cuh)
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
  constexpr int LWX = A_KMAJOR ? 16 : 8;
  if (p.bn <= 112) launch_tc_persistent_inst<A_KMAJOR, B_KMAJOR, EPI, 7, LWX>(p, G, grid, smem, stream, err);
  else launch_tc_persistent_inst<A_KMAJOR, B_KMAJOR, EPI, 8, LWX>(p, G, grid, smem, stream, err);
  return true;
}

}  // namespace cgl
