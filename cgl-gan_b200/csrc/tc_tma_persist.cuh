// tc_tma_persist.cuh -- the TMA-fed grouped GEMM of tc_tma.cuh (weights by TMA -> tcgen05.st -> A operand in TMEM, batch
// operand by TMA, split in place) as a PERSISTENT kernel: one CTA per SM walks over its tiles; the TMA thread, the A converters
// and the B warps run ahead into the next tile while four dedicated epilogue warps drain the accumulator of the finished one.
//
// Why (profiles/tma_ablate_r2.log: per tile ~4.7 us of set-up, pipeline fill and epilogue on top of a main loop of 6 - 25 us):
// with one tile per CTA and one CTA per SM (TMEM and shared memory are both full) nothing overlaps those phases; the short-K
// layers (K = 256 / 512: 8 - 16 k-blocks per tile) spent a third of their time there. Here the barriers, the TMEM allocation and the
// tensor-map prefetch happen once per CTA, the operand rings are full again when the accumulator is released, and the MMA
// warp only idles while the epilogue warps read TMEM (they release it before they store).
//
// Warps (18): 0-3 epilogue (lane quarter = warp % 4) | 4-7 B warps | 8-15 A converters (quarter = warp % 4, k-block half) |
// 16 MMA (converged, the elected lane issues) | 17 TMA (one thread). Barriers as in tc_tma.cuh plus acc_full (commit after
// the last k-block of a tile) / acc_empty (4 epilogue warps). The epilogue stores straight from registers: lane = output
// feature, so a warp store writes 128 contiguous bytes of one batch row.
#pragma once
#include "tc_tma.cuh"

namespace cgl {

constexpr int TCTP_THREADS = 18 * 32;

#ifdef TCTP_DEBUG
__device__ __forceinline__ void mbar_wait_tag(uint32_t bar, uint32_t parity, int tag, int a, int b) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait(bar, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 1000000000ull) {
      if ((threadIdx.x & 31) == 0) printf("TIMEOUT cta %d warp %d tag %d parity %u a %d b %d\n", blockIdx.x, threadIdx.x >> 5, tag, parity, a, b);
      __trap();
    }
  }
}
#define MW(bar, par, tag, a, b) mbar_wait_tag(bar, par, tag, a, b)
#else
#define MW(bar, par, tag, a, b) mbar_wait(bar, par)
#endif

template <bool A_KMAJOR, int EPI>
__global__ void __launch_bounds__(TCTP_THREADS, 1)
tc_tma_persistent_kernel(const TcParams p, const int G, const int m_tiles, const int n_tiles,
                         const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapAt,
                         const __grid_constant__ CUtensorMap tmapB0, const __grid_constant__ CUtensorMap tmapB1) {
  constexpr int EW0 = 0, BW0 = 4, CW0 = 8, MMAW = 16, TMAW = 17;
  constexpr int BT_THREADS = 128;
  constexpr int BKT = 32;
  extern __shared__ __align__(1024) char tc_smem[];
  __shared__ __align__(8) unsigned long long bar_raw_full[TCT_NRAW];
  __shared__ __align__(8) unsigned long long bar_raw_empty[TCT_NRAW];
  __shared__ __align__(8) unsigned long long bar_a_full[TCT_MAX_AS];
  __shared__ __align__(8) unsigned long long bar_a_empty[TCT_MAX_AS];
  __shared__ __align__(8) unsigned long long bar_b_raw[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_b_full[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_b_empty[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_acc_full;
  __shared__ __align__(8) unsigned long long bar_acc_empty;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bn = p.bn;
  const int nsb = p.n_stages;
  const int n_main = p.n_main;
  const int nas = p.tmem_cols;         // (re-used field) A stages in TMEM
  const int stride = bn;
  const uint32_t a_col0 = (uint32_t)((1 + n_main) * bn);
  const uint32_t b_bytes = (uint32_t)bn * BKT * 4;
  const uint32_t bstage_bytes = 2 * b_bytes;
  char* smem = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  char* smem_raw = smem;
  char* smem_b = smem + TCT_NRAW * TCT_RAW_BYTES;
  const int nkb = (p.K + BKT - 1) / BKT;
  const int nks = (p.K + 7) >> 3;
  const int tiles = G * m_tiles * n_tiles;

  if (tid == 0) {
    for (int i = 0; i < TCT_NRAW; ++i) {
      mbar_init(smem_u32(&bar_raw_full[i]), 1);
      mbar_init(smem_u32(&bar_raw_empty[i]), 8);
    }
    for (int i = 0; i < TCT_MAX_AS; ++i) {
      mbar_init(smem_u32(&bar_a_full[i]), 4);
      mbar_init(smem_u32(&bar_a_empty[i]), 1);
    }
    for (int i = 0; i < TC_MAX_STAGES; ++i) {
      mbar_init(smem_u32(&bar_b_raw[i]), 1);
      mbar_init(smem_u32(&bar_b_full[i]), BT_THREADS);
      mbar_init(smem_u32(&bar_b_empty[i]), 1);
    }
    mbar_init(smem_u32(&bar_acc_full), 1);
    mbar_init(smem_u32(&bar_acc_empty), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMAW) tmem_alloc(smem_u32(&tmem_slot), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  // tile t -> (group, M tile, batch tile): batch tiles fastest, so the CTAs that share a weight tile run at the same time
  auto tile_coords = [&](int t, int& g, int& m0, int& n0) {
    const int nt = t % n_tiles;
    const int r = t / n_tiles;
    m0 = (r % m_tiles) * TC_BM;
    g = r / m_tiles;
    n0 = nt * p.n_per;
  };

  if (warp == TMAW) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapB0)) : "memory");
      RingPos r = {0, 0}, rb = {0, 0};
      const uint32_t b_tx = (uint32_t)p.n_per * 128u;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        int g, m0, n0;
        tile_coords(t, g, m0, n0);
        const int rowid = p.A.idx0 ? p.A.idx0[g] : g;
        const bool src1 = n0 >= p.B.rows0;
        const CUtensorMap* tmapB = src1 ? &tmapB1 : &tmapB0;
        int b_row = n0, b_grp;
        if (src1) { b_row -= p.B.rows0; b_grp = p.B.idx1 ? p.B.idx1[g] : g; }
        else b_grp = p.B.idx0 ? p.B.idx0[g] : g;
        for (int kb = 0; kb < nkb; ++kb) {
          if (rb.round > 0) MW(smem_u32(&bar_b_empty[rb.slot]), (rb.round - 1) & 1u, 1, rb.slot, t);
          const uint32_t bbar = smem_u32(&bar_b_raw[rb.slot]);
          mbar_arrive_expect_tx(bbar, b_tx);
          tma_load_3d(smem_u32(smem_b + (size_t)rb.slot * bstage_bytes), tmapB, kb * BKT, b_row, b_grp, bbar);
          rb.next(nsb);
          if (r.round > 0) MW(smem_u32(&bar_raw_empty[r.slot]), (r.round - 1) & 1u, 2, r.slot, t);
          const uint32_t bar = smem_u32(&bar_raw_full[r.slot]);
          const uint32_t dst = smem_u32(smem_raw + (size_t)r.slot * TCT_RAW_BYTES);
          if (A_KMAJOR) {                      // (tail maps: no box row beyond the matrix, see tc_tma.cuh)
            const bool tail = m0 + TC_BM > p.M;
            mbar_arrive_expect_tx(bar, tail ? (uint32_t)(p.M - m0) * 128u : (uint32_t)TCT_RAW_BYTES);
            tma_load_3d(dst, tail ? &tmapAt : &tmapA, kb * BKT, m0, rowid, bar);
          } else {
            const bool tail = kb * BKT + BKT > p.K;
            mbar_arrive_expect_tx(bar, tail ? (uint32_t)(p.K - kb * BKT) * 512u : (uint32_t)TCT_RAW_BYTES);
            tma_load_3d(dst, tail ? &tmapAt : &tmapA, m0, kb * BKT, rowid, bar);
          }
          r.next(TCT_NRAW);
        }
      }
    }
    __syncwarp();
  } else if (warp == MMAW) {
    // ===== MMA issuer (converged warp, elected lane) =====
    const uint32_t idesc = umma_idesc_tf32(false, false, bn);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
    const uint32_t tmem_a0 = tmem_u + a_col0;
    const uint32_t main_lo = tmem_u + (uint32_t)stride, main_hi = tmem_u + (uint32_t)(n_main * stride);
    const uint32_t sb0 = smem_u32(smem_b);
    const uint32_t bar_bf = smem_u32(&bar_b_full[0]), bar_be = smem_u32(&bar_b_empty[0]);
    const uint32_t bar_af = smem_u32(&bar_a_full[0]), bar_ae = smem_u32(&bar_a_empty[0]);
    uint32_t d_main = main_lo;
    int sb = 0;
    uint32_t pb = 0;
    // A stages: half h of a k-block always goes through the slots h, h + 2, ... (its own ring of nas / 2 stages, filled by the
    // four converter warps of that half), so a ragged last k-block (only half 0) never moves a half onto the other's slots
    const int a_depth = nas >> 1;
    int ai[2] = {0, 0};
    uint32_t ap[2] = {0, 0};
    auto block = [&](auto first_c, int nks_here, int ks0) {
      constexpr bool FIRST = decltype(first_c)::value;
      MW(bar_bf + 8u * sb, pb, 3, sb, ks0);
      tc_fence_after();
      const uint32_t sb_hi = sb0 + (uint32_t)sb * bstage_bytes;
      const uint64_t dbh0 = umma_desc(sb_hi, 16u, 1024u, UMMA_LAYOUT_SW128);
      const uint64_t dbl0 = umma_desc(sb_hi + b_bytes, 16u, 1024u, UMMA_LAYOUT_SW128);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h * 2 < nks_here) {
          const int sa = h + 2 * ai[h];
          MW(bar_af + 8u * sa, ap[h], 4, sa, ks0);
          tc_fence_after();
          const uint32_t ta = tmem_a0 + (uint32_t)(sa * 32);
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = h * 2 + jj;
            if (j < nks_here) {
              umma_kstep_ts_warp(tmem_u, d_main, ta + (uint32_t)(jj * 8), ta + (uint32_t)(jj * 8 + 16),
                                 dbh0 + (uint64_t)(2 * j), dbl0 + (uint64_t)(2 * j), idesc,
                                 FIRST ? (ks0 + j > 0 ? 1u : 0u) : 1u, FIRST ? (ks0 + j >= n_main ? 1u : 0u) : 1u);
              d_main = (d_main == main_hi) ? main_lo : d_main + (uint32_t)stride;
            }
          }
          umma_commit_warp(bar_ae + 8u * sa);
          if (++ai[h] == a_depth) { ai[h] = 0; ap[h] ^= 1u; }
        }
      }
      umma_commit_warp(bar_be + 8u * sb);
      if (++sb == nsb) { sb = 0; pb ^= 1u; }
    };
    const int nkb_full = nks >> 2;
    uint32_t tile_i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tile_i) {
      if (tile_i > 0) {
        MW(smem_u32(&bar_acc_empty), (tile_i - 1) & 1u, 5, (int)tile_i, t);   // the epilogue warps have read the previous accumulator
        tc_fence_after();
      }
      d_main = main_lo;
      int kb = 0;
      if (nkb_full > 0) { block(std::true_type{}, 4, 0); kb = 1; }
      for (; kb < nkb_full; ++kb) block(std::false_type{}, 4, kb * 4);
      if (kb < nkb) {
        if (kb == 0) block(std::true_type{}, nks - kb * 4, 0);
        else block(std::false_type{}, nks - kb * 4, kb * 4);
      }
      umma_commit_warp(smem_u32(&bar_acc_full));
    }
    __syncwarp();
  } else if (warp >= CW0) {
    // ===== A converter warps =====
    const int cw = warp - CW0;
    const int cq = cw & 3, ch = cw >> 2;
    const int row = cq * 32 + lane;
    const uint32_t t_lane = tmem_d + ((uint32_t)(cq * 32) << 16) + a_col0;
    const uint32_t bar_rf = smem_u32(&bar_raw_full[0]), bar_re = smem_u32(&bar_raw_empty[0]);
    const uint32_t bar_af = smem_u32(&bar_a_full[0]), bar_ae = smem_u32(&bar_a_empty[0]);
    RingPos rr = {0, 0};
    const int a_depth = nas >> 1;      // this half's own ring: slots ch, ch + 2, ...
    int a_i = 0;
    uint32_t a_round = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        MW(bar_rf + 8u * rr.slot, rr.round & 1u, 6, rr.slot, kb);
        const char* raw = smem_raw + (size_t)rr.slot * TCT_RAW_BYTES;
        float x[16];
        if (A_KMAJOR) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(raw + row * 128 + (((ch * 4 + c) ^ (row & 7)) << 4));
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) x[k] = *reinterpret_cast<const float*>(raw + (ch * 16 + k) * 512 + row * 4);
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float hv = tf32_hi(x[j]);
          hi[j] = __float_as_uint(hv);
          lo[j] = __float_as_uint(x[j] - hv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_re + 8u * rr.slot);
        rr.next(TCT_NRAW);
        if (kb * 4 + ch * 2 < nks) {
          const int a_slot = ch + 2 * a_i;
          if (a_round > 0) {
            MW(bar_ae + 8u * a_slot, (a_round - 1) & 1u, 7, a_slot, kb);
            tc_fence_after();
          }
          const uint32_t ta = t_lane + (uint32_t)(a_slot * 32);
          tmem_st16(ta, hi);
          tmem_st16(ta + 16u, lo);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_af + 8u * a_slot);
          if (++a_i == a_depth) { a_i = 0; ++a_round; }
        }
      }
    }
  } else if (warp >= BW0) {
    // ===== B warps: in-place split of the raw tile =====
    const int bt = tid - BW0 * 32;
    constexpr int IT = (TC_BM * 8 + BT_THREADS - 1) / BT_THREADS;   // 8 float4 per thread at most
    RingPos rb = {0, 0};
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int f4_used = ((p.n_per + 7) & ~7) * 8;
      for (int kb = 0; kb < nkb; ++kb) {
        MW(smem_u32(&bar_b_raw[rb.slot]), rb.round & 1u, 8, rb.slot, kb);
        char* b_hi = smem_b + (size_t)rb.slot * bstage_bytes;
        char* b_lo = b_hi + b_bytes;
        float4 v[IT];
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int f = bt + i * BT_THREADS;
          if (f < f4_used) v[i] = *reinterpret_cast<const float4*>(b_hi + f * 16);
        }
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int f = bt + i * BT_THREADS;
          if (f < f4_used) tc_split_store(b_hi, b_lo, (uint32_t)(f * 16), v[i]);
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bar_b_full[rb.slot]));
        rb.next(nsb);
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> fused op -> global (128 contiguous bytes per warp store) =====
    const int q = warp - EW0;                       // == warp % 4
    const int n_used = nks < n_main ? nks : n_main;
    const int nch = bn >> 4;
    uint32_t tile_i = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++tile_i) {
      int g, m0, n0;
      tile_coords(t, g, m0, n0);
      const int m = m0 + q * 32 + lane;
      const bool m_ok = m < p.M;
      const int n_valid = (p.N - n0 < p.n_per) ? (p.N - n0) : p.n_per;
      const int crow = p.cidx ? p.cidx[g] : g;
      float* C = p.cbase + (long long)crow * p.c_gstride + p.c_off;
      const float* S = (EPI == EPI_BWD_DATA && p.saved) ? p.saved + (long long)g * p.saved_gstride : nullptr;
      float bias = 0.f;
      if (EPI == EPI_FWD && p.bias_base && m_ok) {
        const int brow = p.bias_idx ? p.bias_idx[g] : g;
        bias = __ldg(p.bias_base + (long long)brow * p.bias_gstride + p.bias_off + m);
      }
      MW(smem_u32(&bar_acc_full), tile_i & 1u, 9, (int)tile_i, t);
      tc_fence_after();
      const int nch_used = (n_valid + 15) >> 4;
      for (int c = 0; c < nch && c < nch_used; ++c) {
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
        uint32_t r0[16], r1[16], r2[16], r3[16];
        tmem_ld16_async(taddr + (uint32_t)stride, r0);
        tmem_ld16_async(taddr, r1);
        if (n_used >= 2) tmem_ld16_async(taddr + (uint32_t)(2 * stride), r2);
        if (n_used >= 3) tmem_ld16_async(taddr + (uint32_t)(3 * stride), r3);
        tmem_ld_wait();
        if (c + 1 == nch_used || c + 1 == nch) {
          // the last chunk is in registers: the accumulator goes back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty));
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __uint_as_float(r0[j]);
          if (n_used >= 2) a += __uint_as_float(r2[j]);
          if (n_used >= 3) a += __uint_as_float(r3[j]);
          v[j] = a + __uint_as_float(r1[j]);
        }
        if (!m_ok) continue;
        const int nb = c * 16;
        if (EPI == EPI_BWD_DATA && S) {
          float sv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            sv[j] = (nb + j < n_valid) ? __ldg(S + (long long)(n0 + nb + j) * p.ldc + m) : 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= act_bwd_from_out(sv[j], p.act, p.slope);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (nb + j < n_valid) {
            float o = v[j];
            if (EPI == EPI_FWD) o = act_fwd(o + bias, p.act, p.slope);
            C[(long long)(n0 + nb + j) * p.ldc + m] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMAW) tmem_dealloc(tmem_d, TC_TMEM_COLS);
}

// true: launched (or failed with *err set); false: not applicable. Same conditions as launch_tc_tma with the batch operand by TMA,
// at most three hi*hi regions. Opt-in (tune bit 16777216): measured equal or 1 - 6 % slower than one tile per CTA
// (profiles/tma_persist_r2.log) -- the MMA warp idles while the four epilogue warps drain TMEM, which costs what the
// hidden set-up and pipeline fill save.
template <bool A_KMAJOR, int EPI>
static inline bool launch_tc_tma_persistent(TcParams p, int G, cudaStream_t stream, cudaError_t* err) {
  if (!(tc_tune() & 131072) || !(tc_tune() & 16777216)) return false;
  if (G <= 0 || p.M <= 0 || p.N <= 0) return false;
  if (!p.c_vec || !p.A.vec || !p.B.vec || p.A.rows0 != 0x7fffffff) return false;
  if (p.K < 2 * TC_BK || p.M < 64) return false;
  // short K (<= 43 k-steps: the 128 x 256 layer of the 2DMG discriminator, the generator trunks, the data gradient of 512 <- 256):
  // a tile is 2 - 8 k-blocks between its set-up and its epilogue, and the two-CTAs-per-SM variant of tc_gemm.cuh (one accumulation
  // region, 256 TMEM columns) overlaps those phases -- measured 0.64 against 0.70 ms on the 512 <- 256 data gradient
  if ((tc_tune() & 1024) && (p.K + 7) / 8 <= TC_MAX_ACCUM) return false;
  if (A_KMAJOR ? (p.K % 16 != 0) : (p.K % 8 != 0 || (p.M % TC_BM != 0 && p.K < TC_BM))) return false;
  if (p.K % 16 != 0) return false;                     // batch operand by TMA
  const int w_in = A_KMAJOR ? p.K : p.M, w_out = A_KMAJOR ? p.M : p.K;
  if (p.A.ld != w_in) return false;
  const bool dual = p.B.rows0 < p.N;
  int bn, per, n_main, nas;
  if (!tct_plan(p.N, p.K, dual ? p.B.rows0 : 0x7fffffff, &bn, &per, &n_main, &nas)) return false;
  if (n_main > 3) return false;
  nas &= ~1;                           // one ring of nas / 2 stages per k-block half
  const int rows_b0 = dual ? p.B.rows0 : p.N, rows_b1 = dual ? p.N - p.B.rows0 : 0;
  if (rows_b0 % per || rows_b1 % per) return false;
  CUtensorMap mapA, mapAt, mapB0, mapB1;
  const int box_rows = A_KMAJOR ? TC_BM : TC_BK, box_cols = A_KMAJOR ? TC_BK : TC_BM;
  const int tail_rows = w_out % box_rows;
  if (!tct_make_map(&mapA, p.A.base0, w_in, w_out, w_in, p.A.gstride0, box_cols, w_out < box_rows ? w_out : box_rows, A_KMAJOR)) return false;
  if (tail_rows && w_out > box_rows) {
    if (!tct_make_map(&mapAt, p.A.base0, w_in, w_out, w_in, p.A.gstride0, box_cols, tail_rows, A_KMAJOR)) return false;
  } else {
    mapAt = mapA;
  }
  if (!tct_make_map(&mapB0, p.B.base0, p.K, rows_b0, p.B.ld, p.B.gstride0, TC_BK, per, true)) return false;
  if (dual) { if (!tct_make_map(&mapB1, p.B.base1, p.K, rows_b1, p.B.ld, p.B.gstride1, TC_BK, per, true)) return false; }
  else mapB1 = mapB0;
  p.bn = bn;
  p.n_per = per;
  p.n_main = n_main;
  p.tmem_cols = nas;
  p.tune = tc_tune();
  const size_t bstage = 2 * (size_t)bn * TC_BK * 4;
  int nsb = (int)((TC_SMEM_BUDGET - 1024 - (size_t)TCT_NRAW * TCT_RAW_BYTES) / bstage);
  p.n_stages = nsb > TC_MAX_STAGES ? TC_MAX_STAGES : nsb;
  const size_t smem = (size_t)TCT_NRAW * TCT_RAW_BYTES + (size_t)p.n_stages * bstage + 1024;
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM, n_tiles = (p.N + per - 1) / per;
  const long long tiles = (long long)G * m_tiles * n_tiles;
  if (tiles > 0x7fffffff) return false;
  const int sms = tc_num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    *err = cudaFuncSetAttribute(tc_tma_persistent_kernel<A_KMAJOR, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)TC_SMEM_BUDGET);
    if (*err != cudaSuccess) return true;
  }
  tc_tma_persistent_kernel<A_KMAJOR, EPI><<<grid, TCTP_THREADS, smem, stream>>>(p, G, m_tiles, n_tiles, mapA, mapAt, mapB0, mapB1);
  count_launch();
  *err = cudaGetLastError();
  return true;
}

}  // namespace cgl
