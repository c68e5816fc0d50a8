// tc_tma.cuh -- the tcgen05 grouped GEMM of tc_gemm.cuh with the WEIGHT operand fed by TMA and held in TENSOR MEMORY:
//
//   A (the weights W[out][in] of one group: a plain dense matrix inside the packed parameter row) is fetched as raw fp32
//   tiles by cp.async.bulk.tensor (one elected thread, a 3-D tensor map (in, out, bank row) over the parameter bank,
//   zero fill beyond the matrix), four converter warps read their own accumulator lane's row of the tile from shared
//   memory, split it into hi = tf32(x) / lo = x - hi in registers and write both with tcgen05.st into TMEM, and the MMAs
//   take A from TMEM:    tcgen05.mma.cta_group::1.kind::tf32 [d_tmem], [a_tmem], b_desc, idesc
//   B (the batch rows: activations / output gradients, K-major) comes by TMA as well (BT = true): the raw tile lands in the
//   SWIZZLE_128B layout the MMA descriptors read, eight warps split it IN PLACE (LDS -> hi back to the same bytes, lo to the
//   twin buffer; the split is elementwise, so the swizzle never has to be computed), fence.proxy.async, arrive. Batches that
//   are concatenated from two sources (real | fake rows) use one tensor map per source and batch tiles that start at
//   multiples of the per-source row count; index-selected groups (a server's fake batch shared by its clients) are the
//   group coordinate. A B operand TMA cannot describe (unaligned, a tile straddling two sources) keeps the register path
//   of tc_gemm.cuh (BT = false): LDG -> split -> STS hi/lo by the same eight warps.
//
// Why (profiles/ncu_fwd_r1.md, ncu_fwd_pair2_r2.md): the shared-memory-operand kernels are bound by the L1TEX data pipe
// and by the loaders' instruction stream. Per 32-wide k-block a CTA moved A through that pipe as 16 KB of LDG, 32 KB of
// hi/lo STS and 12 x 4 KB of tensor-core operand reads (96 KB of the 180 KB); here A costs the 16 KB TMA write and 16 KB
// of LDS, the split copies never touch shared memory, and the tensor core reads them from TMEM (180 -> ~115 KB per
// k-block against 672 clk of MMA time). The loader warps only carry B: half of the LDG / STS instructions.
//
// TMEM plan (512 columns): accumulator regions of exactly bn columns (region 0: corrections, regions 1..n_main: hi*hi,
// at most TC_MAX_ACCUM accumulations each, as in tc_gemm.cuh) and, behind them, NAS >= 2 A stages of 32 columns
// (16 k: [hi 16 | lo 16]); bn = 112 and K = 784 / 1024 leave exactly two.
//
// Warps: LWB B warps | 8 A converter warps (lane quarter = warp % 4, k-block half = warp / 4) | 1 MMA warp (converged,
// the elected lane issues) | 1 TMA warp (one thread) | 1 more MMA warp (idle unless CGL_TUNE bit 4 asks for two issuing warps).
// Barriers: raw_full/raw_empty (TMA -> A converters), a_full (a ring of 2 x NAS barriers over the NAS stages) / a_empty
// (A converters -> MMA, 16 k), b_raw (TMA -> B warps), b_full/b_empty (B warps -> MMA -> TMA, 32 k), tok (dual issue), done.
// The epilogue (B + A converter warps) is the float4 one of tc_gemm.cuh.
//
// What paces it (profiles/tma_feed_abl_r2.log, tma_agents_r2.log): not the operand bytes and not the
// shared-memory pipe -- dropping every TMA transfer or the hi store of the batch operand changes nothing -- but the
// hand-shakes: ~1500 clk per k-block against ~700 clk of MMAs, the A-stage round trip (commit -> converter -> tcgen05.st ->
// arrive -> issue) where only two A stages fit, and each agent's serial chain per k-block right behind it. The bits
// 2 / 4 / 262144 / 33554432 and the ablations 67108864 ... 1073741824 are the experiments of that analysis (TcParams::tune).
#pragma once
#include <cuda.h>
#include <stdio.h>

#include <type_traits>

#include "tc_pair.cuh"

namespace cgl {

constexpr int TCT_NRAW = 4;            // raw fp32 A tiles (16 KB each) in flight by TMA
constexpr int TCT_RAW_BYTES = TC_BM * TC_BK * 4;
constexpr int TCT_MAX_AS = 4;          // A stages in TMEM (32 columns each)
#ifndef TCT_DEPTH
#define TCT_DEPTH 3                    // k-blocks of B loads in flight per loader thread
#endif

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32: A = 128 lanes x 8 columns (K-major: lane = m, column = k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same instruction issued from a CONVERGED warp: every lane executes the wrapper with identical operands and the
// instruction itself runs on the lane elect.sync picks. (Inside an `if (lane == 0)` region nvcc wraps every tcgen05.mma in
// a waterfall loop -- ELECT / R2UR / UTCHMMA / BRA.U.ANY -- and the issuing thread then needs ~135 clk per MMA, measured
// with the MMA-only ablation below: more than twice the 56 clk the tensor core takes for a 128 x 112 x 8 tf32 MMA.)
__device__ __forceinline__ void umma_tf32_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one k-step of the 3xTF32 product from a converged warp, one election: corr += lo*hi, corr += hi*lo, main (+)= hi*hi
__device__ __forceinline__ void umma_kstep_ts_warp(uint32_t d_corr, uint32_t d_main, uint32_t a_hi, uint32_t a_lo, uint64_t dbh,
                                                   uint64_t dbl, uint32_t idesc, uint32_t acc_corr, uint32_t acc_main) {
  asm volatile(
      "{\n\t.reg .pred pc, pm, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pc, %7, 0;\n\t"
      "setp.ne.b32 pm, %8, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%3], %4, %6, pc;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %5, %6, 1;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %4, %6, pm;\n\t}"
      ::"r"(d_corr), "r"(d_main), "r"(a_hi), "r"(a_lo), "l"(dbh), "l"(dbl), "r"(idesc), "r"(acc_corr), "r"(acc_main)
      : "memory");
}
// one arrival from a converged warp (the elected lane arrives; no divergent region around the tcgen05 issue loop)
__device__ __forceinline__ void mbar_arrive_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}"
      ::"r"(bar) : "memory");
}
// non-blocking test of a barrier phase (acquire, like the waits)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// the same k-step on a CTA pair (cta_group::2: M = 256, A from the TMEM of both CTAs, each CTA's shared memory holds half of
// the B lines), and the commit that arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_kstep_ts_warp_pair(uint32_t d_corr, uint32_t d_main, uint32_t a_hi, uint32_t a_lo, uint64_t dbh,
                                                        uint64_t dbl, uint32_t idesc, uint32_t acc_corr, uint32_t acc_main) {
  asm volatile(
      "{\n\t.reg .pred pc, pm, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pc, %7, 0;\n\t"
      "setp.ne.b32 pm, %8, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::tf32 [%0], [%3], %4, %6, pc;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::tf32 [%0], [%2], %5, %6, 1;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::tf32 [%1], [%2], %4, %6, pm;\n\t}"
      ::"r"(d_corr), "r"(d_main), "r"(a_hi), "r"(a_lo), "l"(dbh), "l"(dbl), "r"(idesc), "r"(acc_corr), "r"(acc_main)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_warp(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}

// use-counted ring positions: the n-th use of an NST-deep ring is slot n % NST; its "full" phase has parity (n / NST) & 1,
// and the slot may be refilled for use n >= NST once the "empty" phase (n / NST - 1) has completed
struct RingPos {
  int slot;
  uint32_t round;   // n / NST
  __device__ __forceinline__ void next(int nst) {
    if (++slot == nst) { slot = 0; ++round; }
  }
};

// -DTCT_PROFILE (variant builds for profiles/tma_agents.py): CTA (0, 0, g) records in g_tc_timeline[g * 32 + i] clock64 stamps
// (0 entry, 1 set-up done, 2 / 3 MMA loop start / end, 4 / 5 epilogue start / end) and the cycles its agents spent inside their
// barrier waits (6 MMA warp on b_full, 7 on a_full, 8 TMA thread on b_empty, 9 on raw_empty, 10 converter warp 0 on raw_full,
// 11 on a_empty, 12 B warp 0 on b_raw) and in their whole loops (13 converter warp 0, 14 B warp 0, 15 TMA thread); 16 / 17: the
// MMA warp's cycles issuing tcgen05.mma / tcgen05.commit
#ifdef TCT_PROFILE
#define TCT_TIMED(acc, stmt) do { const long long _c0 = clock64(); stmt; acc += clock64() - _c0; } while (0)
#define TCT_PUT(i, v) do { if (prof) g_tc_timeline[blockIdx.z * 32 + (i)] = (v); } while (0)
#else
#define TCT_TIMED(acc, stmt) do { stmt; } while (0)
#define TCT_PUT(i, v) do { } while (0)
#endif
// B is K-major (lines = batch rows). EPI: EPI_FWD (A K-major: forward) / EPI_BWD_DATA (A MN-major: data gradient; saved == NULL
// stores the plain product). LWB = loader warps (multiple of 4).
// PAIR (needs BT): two CTAs of a cluster (adjacent M tiles of one group) run every MMA together (cta_group::2, issued by the CTA
// of cluster rank 0): each stages, splits and holds only HALF of the batch lines (n_per / 2 rows, accumulator columns
// [0, bn/2) and [bn/2, bn)), which takes 49 of the 130 KB per k-block off each CTA's shared-memory pipe -- what bounds the
// one-CTA kernel (profiles/tma_ablate_r2.log). The peer's MMA warp forwards its full barriers to the leader (one remote
// arrive each); the commits arrive on the empty barriers of both CTAs. An odd number of M tiles adds a CTA without rows.
template <bool A_KMAJOR, int EPI, int LWB, bool BT, bool PAIR>
__global__ void __launch_bounds__((LWB + 11) * 32, 1)
tc_tma_gemm_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapAt,
                   const __grid_constant__ CUtensorMap tmapB0, const __grid_constant__ CUtensorMap tmapB1) {
  constexpr int CW0 = LWB;             // first converter warp
  constexpr int MMAW = LWB + 8;
  constexpr int TMAW = LWB + 9;
  constexpr int MMAW2 = LWB + 10;      // second issuing warp (dual issue: the k-blocks alternate between the two)
  constexpr int LT = LWB * 32;         // loader threads
  constexpr int ET = (LWB + 8) * 32;   // epilogue threads (B warps + A converters)
  constexpr int BKT = 32;
  constexpr int NBW = 32 / LWB;        // B patches (4 lines x 32 k) per loader warp: bn <= 128 -> <= 32 patches
  extern __shared__ __align__(1024) char tc_smem[];
  __shared__ __align__(8) unsigned long long bar_raw_full[TCT_NRAW];
  __shared__ __align__(8) unsigned long long bar_raw_empty[TCT_NRAW];
  __shared__ __align__(8) unsigned long long bar_a_full[2 * TCT_MAX_AS];   // ring of abars barriers over the nas stages (below)
  __shared__ __align__(8) unsigned long long bar_a_empty[TCT_MAX_AS];
  __shared__ __align__(8) unsigned long long bar_b_raw[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_b_full[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_b_empty[TC_MAX_STAGES];
  __shared__ __align__(8) unsigned long long bar_peer_a[TCT_MAX_AS];     // PAIR, leader: the peer's A stage is full
  __shared__ __align__(8) unsigned long long bar_peer_b[TC_MAX_STAGES];  // PAIR, leader: the peer's B stage is full
  __shared__ __align__(8) unsigned long long bar_done;
  __shared__ __align__(8) unsigned long long bar_tok[2];                 // dual issue: warp i may issue its next k-block
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.z;
#ifdef TCT_PROFILE
  const bool prof = g_tc_timeline != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
  long long w0 = 0, w1 = 0, w2 = 0, w3 = 0, loop0 = 0;
  if (tid == 0) TCT_PUT(0, clock64());
#endif
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // (a kernel that uses cta_group::2 is only accepted with an even cluster width in x: the pair lies along x)
  const int m0 = (PAIR ? blockIdx.x : blockIdx.y) * TC_BM;
  const int n0 = (PAIR ? blockIdx.y : blockIdx.x) * p.n_per;  // batch tiles start at multiples of n_per <= bn (BT: never inside two sources)
  const int bn = p.bn;
  const int hp = p.n_per >> 1, bh = bn >> 1;   // PAIR: batch rows / accumulator columns per CTA
  const int n_valid = (p.N - n0 < p.n_per) ? (p.N - n0) : p.n_per;   // rows of this tile that are stored
  const int nsb = p.n_stages;          // B stages in shared memory
  const int n_main = p.n_main;
  const int nas = p.tmem_cols;         // (re-used field) A stages in TMEM
  const int stride = bn;               // accumulator regions are packed
  const uint32_t a_col0 = (uint32_t)((1 + n_main) * bn);   // first A-stage column

  const uint32_t b_bytes = (uint32_t)bn * BKT * 4;          // bn % 8 == 0 -> multiple of 1024
  const uint32_t bstage_bytes = 2 * b_bytes;
  char* smem = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);
  char* smem_raw = smem;                                    // TCT_NRAW raw A tiles
  char* smem_b = smem + TCT_NRAW * TCT_RAW_BYTES;           // nsb x [B hi | B lo]

  const int nkb = (p.K + BKT - 1) / BKT;
  const int nks = (p.K + 7) >> 3;
  const int rowid = p.A.idx0 ? p.A.idx0[g] : g;             // bank row of this group's weights
  // Dual issue (experiment, tune bit 4): the issuing warp is ONE dependent instruction chain per k-block -- three barrier
  // waits (~130-170 clk each, even on a completed phase), three tcgen05.fence, twelve tcgen05.mma (~45 clk each to issue) and
  // three tcgen05.commit (~55 clk). Two warps take the k-blocks in turn; a token barrier hands the right to issue from one to
  // the other, so the MMAs still enter the tensor pipe in k order (bit-identical results) while each warp's waits, fences and
  // commits run under the other's MMAs. Measured (profiles/tma_dual_r2.log): correct, and 15 - 20 % SLOWER with two A stages
  // (the wait for both halves of a k-block before the token lengthens the A-stage round trip), 2 % faster with four.
  const bool dual_issue = !PAIR && !(p.tune & 1048576) && (p.tune & 4) && p.n_stages >= 2;
  // The "A stage full" barriers form a ring of abars = 2 * nas barriers over the nas stages (use u of the ring: stage u % nas,
  // barrier u % abars): with two issuing warps the waits for k-block kb + 1 start before those for kb have returned, and with
  // ONE barrier per stage a wait on the next phase of a barrier whose current phase is still open returns at once (a parity
  // wait cannot tell phase n + 1 from phase n - 1). Consecutive k-blocks never share a barrier this way (nas >= 2 stages of
  // 16 k = at least one k-block; the B ring has n_stages >= 2 barriers of its own).
  const int abars = PAIR ? p.tmem_cols : 2 * p.tmem_cols;

  if (tid == 0) {
    for (int i = 0; i < TCT_NRAW; ++i) {
      mbar_init(smem_u32(&bar_raw_full[i]), 1);
      mbar_init(smem_u32(&bar_raw_empty[i]), 8);
    }
    for (int i = 0; i < 2 * TCT_MAX_AS; ++i) mbar_init(smem_u32(&bar_a_full[i]), 4);
    for (int i = 0; i < TCT_MAX_AS; ++i) {
      mbar_init(smem_u32(&bar_a_empty[i]), 1);
      mbar_init(smem_u32(&bar_peer_a[i]), 1);
    }
    for (int i = 0; i < TC_MAX_STAGES; ++i) {
      mbar_init(smem_u32(&bar_b_raw[i]), 1);
      mbar_init(smem_u32(&bar_b_full[i]), LT);
      mbar_init(smem_u32(&bar_b_empty[i]), 1);
      mbar_init(smem_u32(&bar_peer_b[i]), 1);
    }
    mbar_init(smem_u32(&bar_done), dual_issue ? 2 : 1);
    mbar_init(smem_u32(&bar_tok[0]), 1);
    mbar_init(smem_u32(&bar_tok[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMAW) {
    if (PAIR) tmem_alloc_pair(smem_u32(&tmem_slot), TC_TMEM_COLS);
    else tmem_alloc(smem_u32(&tmem_slot), TC_TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // barriers initialised and TMEM allocated in both CTAs before anything is signalled
  else __syncthreads();
  tc_fence_after();
#ifdef TCT_PROFILE
  if (tid == 0) TCT_PUT(1, clock64());
  loop0 = clock64();
#endif
  const uint32_t tmem_d = tmem_slot;
  // bring-up ablations (CGL_TUNE): 4096 = MMAs only (nothing is fed, the MMA thread never waits: what the tensor pipe
  // alone costs), 8192 = no MMAs (what feeding alone costs)
  const bool mma_only = (p.tune & 4096) != 0, no_mma = (p.tune & 8192) != 0;
  const int nkb_feed = mma_only ? 0 : ((p.K + BKT - 1) / BKT);

  if (warp == TMAW) {
    // ===== TMA producer: one thread, raw fp32 tiles of W =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
      // B source of this tile: rows [n0, n0 + n_per) lie in one source (host-checked)
      const bool src1 = BT && n0 >= p.B.rows0;
      const CUtensorMap* tmapB = src1 ? &tmapB1 : &tmapB0;
      int b_row = n0 + (PAIR ? (int)rank * hp : 0), b_grp = 0;
      const uint32_t b_tx = (uint32_t)(PAIR ? hp : p.n_per) * 128u;   // the box is n_per (PAIR: n_per / 2) rows: always inside its source
      if (BT) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmapB)) : "memory");
        if (src1) { b_row -= p.B.rows0; b_grp = p.B.idx1 ? p.B.idx1[g] : g; }
        else b_grp = p.B.idx0 ? p.B.idx0[g] : g;
      }
      RingPos r = {0, 0}, rb = {0, 0};
      // bring-up ablations (results are garbage): 67108864 = no B transfers, 134217728 = no A transfers -- what the feed costs
      // without the bytes of one operand crossing L2 -> SM
      const bool skip_b = (p.tune & 67108864) != 0, skip_a = (p.tune & 134217728) != 0;
      for (int kb = 0; kb < nkb_feed; ++kb) {
        if (BT) {
          if (rb.round > 0) TCT_TIMED(w0, mbar_wait(smem_u32(&bar_b_empty[rb.slot]), (rb.round - 1) & 1u));   // the MMAs that read it are done
          const uint32_t bbar = smem_u32(&bar_b_raw[rb.slot]);
          if (skip_b) {                      // bring-up ablation: the stage is declared full without a transfer
            mbar_arrive(bbar);
          } else {
            mbar_arrive_expect_tx(bbar, b_tx);
            tma_load_3d(smem_u32(smem_b + (size_t)rb.slot * bstage_bytes), tmapB, kb * BKT, b_row, b_grp, bbar);
          }
          rb.next(nsb);
        }
        if (r.round > 0) TCT_TIMED(w1, mbar_wait(smem_u32(&bar_raw_empty[r.slot]), (r.round - 1) & 1u));
        const uint32_t bar = smem_u32(&bar_raw_full[r.slot]);
        const uint32_t dst = smem_u32(smem_raw + (size_t)r.slot * TCT_RAW_BYTES);
        // A box never reaches beyond the last ROW of the weight matrix (the ragged last tile uses the tail map, whose box
        // has exactly the rows that exist): TMA was measured to touch the addresses of out-of-bounds box rows
        // (profiles/tma_repro.py faults with the operand at the end of its allocation), and behind the last matrix of the
        // last bank row there may be nothing mapped. Rows of the stage the tail box leaves unwritten only reach
        // accumulator rows m >= M (forward) or k-steps that are never issued (data gradient, K % 8 == 0).
        if ((PAIR && m0 >= p.M) || skip_a) {   // the CTA that completes an odd number of M tiles: no rows, nothing to fetch
          mbar_arrive(bar);
        } else if (A_KMAJOR) {               // box (32 k, 128 | M % 128 rows m): 128-byte rows, SWIZZLE_128B
          const bool tail = m0 + TC_BM > p.M;
          mbar_arrive_expect_tx(bar, tail ? (uint32_t)(p.M - m0) * 128u : (uint32_t)TCT_RAW_BYTES);
          tma_load_3d(dst, tail ? &tmapAt : &tmapA, kb * BKT, m0, rowid, bar);
        } else {                             // box (128 m, 32 | K % 32 rows k): 512-byte rows, no swizzle
          const bool tail = kb * BKT + BKT > p.K;
          mbar_arrive_expect_tx(bar, tail ? (uint32_t)(p.K - kb * BKT) * 512u : (uint32_t)TCT_RAW_BYTES);
          tma_load_3d(dst, tail ? &tmapAt : &tmapA, m0, kb * BKT, rowid, bar);
        }
        r.next(TCT_NRAW);
      }
#ifdef TCT_PROFILE
      TCT_PUT(8, w0); TCT_PUT(9, w1); TCT_PUT(15, clock64() - loop0);
#endif
    }
    __syncwarp();
  } else if (warp == MMAW || warp == MMAW2) {
    // ===== MMA issuer: the whole warp runs the loop converged, the elected lane issues (see umma_tf32_ts_warp) =====
    // The loop is the critical instruction stream of the kernel (one warp, dependent issue): the first k-block (accumulate
    // flags) and a ragged last one (k-steps beyond K) are peeled off so that the blocks in between are branch-free.
    if (warp == MMAW2 && !dual_issue) {
      // (idle: one issuing warp)
    } else if (!PAIR && (p.tune & 1048576)) {
      // (comparison: the one-thread issue loop this kernel started with -- ~135 clk per MMA)
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_tf32(false, false, bn);
        RingPos rb = {0, 0}, ra = {0, 0}, rf = {0, 0};   // B stage, A stage, "A stage full" barrier
        int ks = 0, reg = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&bar_b_full[rb.slot]), rb.round & 1u);
          tc_fence_after();
          const uint32_t sb_hi = smem_u32(smem_b + (size_t)rb.slot * bstage_bytes);
          for (int h = 0; h < 2; ++h) {
            if (ks < nks) {
              mbar_wait(smem_u32(&bar_a_full[rf.slot]), rf.round & 1u);
              rf.next(abars);
              tc_fence_after();
              const uint32_t ta = tmem_d + a_col0 + (uint32_t)(ra.slot * 32);
              for (int jj = 0; jj < 2; ++jj) {
                if (ks < nks) {
                  const int j = h * 2 + jj;
                  const uint64_t dbh = umma_desc(sb_hi + j * 32u, 16u, 1024u, UMMA_LAYOUT_SW128);
                  const uint64_t dbl = umma_desc(sb_hi + b_bytes + j * 32u, 16u, 1024u, UMMA_LAYOUT_SW128);
                  const uint32_t a_hi = ta + (uint32_t)(jj * 8), a_lo = a_hi + 16u;
                  const uint32_t main_col = (uint32_t)((1 + reg) * stride);
                  if (++reg == n_main) reg = 0;
                  umma_tf32_ts(tmem_d, a_lo, dbh, idesc, ks > 0 ? 1u : 0u);
                  umma_tf32_ts(tmem_d, a_hi, dbl, idesc, 1u);
                  umma_tf32_ts(tmem_d + main_col, a_hi, dbh, idesc, ks >= n_main ? 1u : 0u);
                  ++ks;
                }
              }
              umma_commit(smem_u32(&bar_a_empty[ra.slot]));
              ra.next(nas);
            }
          }
          umma_commit(smem_u32(&bar_b_empty[rb.slot]));
          rb.next(nsb);
        }
        umma_commit(smem_u32(&bar_done));
      }
    } else if (PAIR && rank != 0) {
      // ===== forwarder (the peer's MMA warp): this CTA's stages are full -> one cluster-scope arrive each on the leader =====
      if (lane == 0) {
        const uint32_t peer_b0 = mapa_shared(smem_u32(&bar_peer_b[0]), 0), peer_a0 = mapa_shared(smem_u32(&bar_peer_a[0]), 0);
        int sb = 0, sa = 0;
        uint32_t pb = 0, pa = 0;
        for (int kb = 0; kb < nkb_feed; ++kb) {
          // one remote arrive per k-block (a release at cluster scope costs the forwarding thread ~0.5 us): the B stage and
          // both A stages of the block are full
          mbar_wait(smem_u32(&bar_b_full[sb]), pb);
          for (int h = 0; h < 2; ++h) {
            if (kb * 4 + h * 2 < nks) {
              mbar_wait(smem_u32(&bar_a_full[sa]), pa);
              if (++sa == nas) { sa = 0; pa ^= 1u; }
            }
          }
          tc_fence_after();
          tc_fence_before();
          mbar_arrive_cluster(peer_b0 + 8u * (uint32_t)sb);
          if (++sb == nsb) { sb = 0; pb ^= 1u; }
          (void)peer_a0;
        }
      }
    } else {
      const uint32_t idesc = PAIR ? umma_idesc_tf32_pair(false, false, bn) : umma_idesc_tf32(false, false, bn);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_d, 0);
      const uint32_t tmem_a0 = tmem_u + a_col0;
      const uint32_t main_lo = tmem_u + (uint32_t)stride, main_hi = tmem_u + (uint32_t)(n_main * stride);   // first / last main region
      const uint32_t sb0 = smem_u32(smem_b);
      const uint32_t bar_bf = smem_u32(&bar_b_full[0]), bar_be = smem_u32(&bar_b_empty[0]);
      const uint32_t bar_af = smem_u32(&bar_a_full[0]), bar_ae = smem_u32(&bar_a_empty[0]);
      const uint32_t bar_pa = smem_u32(&bar_peer_a[0]), bar_pb = smem_u32(&bar_peer_b[0]);
      uint32_t d_main = main_lo;           // main region of the next k-step
      int sb = 0, sa = 0, fa = 0;          // B stage, A stage, "A stage full" barrier of the next use
      uint32_t pb = 0, pa = 0;             // parities of the current rounds (pa: of the a_full barrier ring)
      const bool probe = !PAIR && !dual_issue && (p.tune & 2);    // tune bit 2: look-ahead barrier tests (measured: the issuing warp's wait time
                                                   // halves, the k-block period does not move -- profiles/tma_agents_r2.log)
      uint32_t a_ok = 0, b_ok = 0;                 // the next A / B stage was already full when it was tested
      const int me = (warp == MMAW) ? 0 : 1;       // dual issue: this warp issues the k-blocks kb % 2 == me
      const uint32_t bar_tk = smem_u32(&bar_tok[0]);
      uint32_t tok_par = 0;
      bool waited = false, pass_token = false;
#ifdef TCT_PROFILE
      long long w4 = 0;
#endif
      // one 32-wide k-block: FIRST = accumulate flags of the first k-steps, KS = k-steps that carry data (1..4)
      auto block = [&](auto first_c, int nks_here, int ks0) {
        constexpr bool FIRST = decltype(first_c)::value;
        if (!mma_only && !waited) {
          if (!__all_sync(0xffffffffu, b_ok)) TCT_TIMED(w0, mbar_wait(bar_bf + 8u * sb, pb));
          if (PAIR) mbar_wait_cluster(bar_pb + 8u * sb, pb);
        }
        tc_fence_after();
        const uint32_t sb_hi = sb0 + (uint32_t)sb * bstage_bytes;
        const uint64_t dbh0 = umma_desc(sb_hi, 16u, 1024u, UMMA_LAYOUT_SW128);
        const uint64_t dbl0 = umma_desc(sb_hi + b_bytes, 16u, 1024u, UMMA_LAYOUT_SW128);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h * 2 < nks_here) {
            if (!mma_only && !waited) {
              if (!__all_sync(0xffffffffu, a_ok)) TCT_TIMED(w1, mbar_wait(bar_af + 8u * fa, pa));
            }
            tc_fence_after();
            if (probe && !mma_only) {
              // (experiment) The issuing warp is one dependent chain (wait -> MMAs -> commit -> wait ...), and a barrier wait costs
              // it ~100 clk even when the phase completed long ago. Here the NEXT barriers are tested now, without blocking, and
              // the answers arrive under the MMAs issued below: a stage that was already full is then entered without a wait. The answers are only looked at where the wait would be (a use right here would stall the
              // warp for the test's latency); warp-uniform through a vote there: a lane that saw "not yet" sends all into the wait.
              const int fan = (fa + 1 == abars) ? 0 : fa + 1;
              a_ok = mbar_test(bar_af + 8u * fan, (fa + 1 == abars) ? (pa ^ 1u) : pa);
              if (h == 1) {
                const int sbn = (sb + 1 == nsb) ? 0 : sb + 1;
                b_ok = mbar_test(bar_bf + 8u * sbn, (sb + 1 == nsb) ? (pb ^ 1u) : pb);
              }
            }
            const uint32_t ta = tmem_a0 + (uint32_t)(sa * 32);
#ifdef TCT_PROFILE
            const long long cm0 = clock64();
#endif
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = h * 2 + jj;
              if (j < nks_here) {
                if (!no_mma) {
                  if (PAIR)
                    umma_kstep_ts_warp_pair(tmem_u, d_main, ta + (uint32_t)(jj * 8), ta + (uint32_t)(jj * 8 + 16),
                                            dbh0 + (uint64_t)(2 * j), dbl0 + (uint64_t)(2 * j), idesc,
                                            FIRST ? (ks0 + j > 0 ? 1u : 0u) : 1u, FIRST ? (ks0 + j >= n_main ? 1u : 0u) : 1u);
                  else
                    umma_kstep_ts_warp(tmem_u, d_main, ta + (uint32_t)(jj * 8), ta + (uint32_t)(jj * 8 + 16),
                                       dbh0 + (uint64_t)(2 * j), dbl0 + (uint64_t)(2 * j), idesc,
                                       FIRST ? (ks0 + j > 0 ? 1u : 0u) : 1u, FIRST ? (ks0 + j >= n_main ? 1u : 0u) : 1u);
                }
                d_main = (d_main == main_hi) ? main_lo : d_main + (uint32_t)stride;
              }
            }
#ifdef TCT_PROFILE
            const long long cm1 = clock64();
            w2 += cm1 - cm0;
#endif
            if (PAIR) umma_commit_pair_warp(bar_ae + 8u * sa);
            else umma_commit_warp(bar_ae + 8u * sa);
#ifdef TCT_PROFILE
            w3 += clock64() - cm1;
#endif
            if (++sa == nas) sa = 0;
            if (++fa == abars) { fa = 0; pa ^= 1u; }
          }
        }
        if (pass_token) {                  // dual issue: this block's MMAs are in the pipe, the other warp may issue the next block
          tc_fence_before();
          mbar_arrive_elect(bar_tk + 8u * (uint32_t)(me ^ 1));
        }
        if (PAIR) umma_commit_pair_warp(bar_be + 8u * sb);
        else TCT_TIMED(w3, umma_commit_warp(bar_be + 8u * sb));
        if (++sb == nsb) { sb = 0; pb ^= 1u; }
      };
      // dual issue: the ring positions move over a k-block the other warp issues
      auto skip = [&](int nks_here) {
        for (int h = 0; h < (nks_here > 2 ? 2 : 1); ++h) {
          if (++sa == nas) sa = 0;
          if (++fa == abars) { fa = 0; pa ^= 1u; }
        }
        for (int j = 0; j < nks_here; ++j) d_main = (d_main == main_hi) ? main_lo : d_main + (uint32_t)stride;
        if (++sb == nsb) { sb = 0; pb ^= 1u; }
      };
      // dual issue: everything this block needs is waited for BEFORE the token is taken, so that its twelve MMAs are issued
      // back to back while this warp holds the right to issue
      auto take = [&](int nks_here, bool need_token) {
        if (!mma_only) {
          TCT_TIMED(w0, mbar_wait(bar_bf + 8u * sb, pb));
          TCT_TIMED(w1, mbar_wait(bar_af + 8u * fa, pa));
          if (nks_here > 2) {
            const int fan = (fa + 1 == abars) ? 0 : fa + 1;
            TCT_TIMED(w1, mbar_wait(bar_af + 8u * fan, (fa + 1 == abars) ? (pa ^ 1u) : pa));
          }
        }
        if (need_token) {
          TCT_TIMED(w4, mbar_wait(bar_tk + 8u * (uint32_t)me, tok_par));
          tok_par ^= 1u;
        }
      };
      const int nkb_full = nks >> 2;       // k-blocks whose four k-steps all carry data
#ifdef TCT_PROFILE
      if (warp == MMAW) TCT_PUT(2, clock64());
#endif
      int kb = 0;
      if (dual_issue) {
        waited = true;
        for (; kb < nkb; ++kb) {
          const int nks_here = (nks - kb * 4 < 4) ? (nks - kb * 4) : 4;
          if ((kb & 1) != me) { skip(nks_here); continue; }
          take(nks_here, kb > 0);
          pass_token = kb + 1 < nkb;
          if (kb == 0) block(std::true_type{}, nks_here, 0);
          else if (nks_here == 4) block(std::false_type{}, 4, kb * 4);
          else block(std::false_type{}, nks_here, kb * 4);
        }
      } else {
        if (nkb_full > 0) { block(std::true_type{}, 4, 0); kb = 1; }
        for (; kb < nkb_full; ++kb) block(std::false_type{}, 4, kb * 4);
        if (kb < nkb) {
          if (kb == 0) block(std::true_type{}, nks - kb * 4, 0);
          else block(std::false_type{}, nks - kb * 4, kb * 4);
        }
      }
      if (PAIR) umma_commit_pair_warp(smem_u32(&bar_done));
      else umma_commit_warp(smem_u32(&bar_done));     // (dual issue: one arrival per issuing warp, each for its own MMAs)
#ifdef TCT_PROFILE
      if (warp == MMAW) { TCT_PUT(3, clock64()); TCT_PUT(6, w0); TCT_PUT(7, w1); TCT_PUT(16, w2); TCT_PUT(17, w3); TCT_PUT(18, w4); }
#endif
    }
    __syncwarp();
  } else {
    if (warp >= CW0) {
      // ===== A converter warps: raw tile row of this thread's TMEM lane -> hi / lo -> tcgen05.st =====
      // eight warps: lane quarter cq = warp % 4 (the TMEM lanes a warp may access), half ch of every k-block (16 k = one A stage)
      const int cw = warp - CW0;
      const int cq = cw & 3, ch = cw >> 2;
      const int row = cq * 32 + lane;
      const uint32_t t_lane = tmem_d + ((uint32_t)(cq * 32) << 16) + a_col0;
      const uint32_t bar_rf = smem_u32(&bar_raw_full[0]), bar_re = smem_u32(&bar_raw_empty[0]);
      const uint32_t bar_af = smem_u32(&bar_a_full[0]), bar_ae = smem_u32(&bar_a_empty[0]);
      RingPos rr = {0, 0};
      const bool abl_a = (p.tune & 268435456) != 0;   // bring-up ablation: barrier protocol only (no LDS / split / tcgen05.st)
      int a_slot = ch;               // A-stage use n = 2 kb + ch -> slot n % nas, round n / nas (nas >= 2)
      int a_bar = ch;                // ... and "full" barrier n % abars
      uint32_t a_round = 0;
      uint32_t hi[16], lo[16];
      // raw tile of the next k-block -> this thread's hi / lo registers; the raw slot goes back to the TMA thread (8 arrivals)
      auto load_split = [&]() {
        TCT_TIMED(w0, mbar_wait(bar_rf + 8u * rr.slot, rr.round & 1u));
        const char* raw = smem_raw + (size_t)rr.slot * TCT_RAW_BYTES;
        float x[16];
        if (abl_a) {
#pragma unroll
          for (int k = 0; k < 16; ++k) x[k] = 0.f;
        } else if (A_KMAJOR) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(raw + row * 128 + (((ch * 4 + c) ^ (row & 7)) << 4));
            x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) x[k] = *reinterpret_cast<const float*>(raw + (ch * 16 + k) * 512 + row * 4);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float hv = tf32_hi(x[j]);
          hi[j] = __float_as_uint(hv);
          lo[j] = __float_as_uint(x[j] - hv);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_re + 8u * rr.slot);
        rr.next(TCT_NRAW);
      };
      // tune bit 262144: the loads and the split of the NEXT k-block run between the TMEM stores of this one and their
      // tcgen05.wait::st (the converter is a serial chain per k-block; measured: profiles/tma_agents_r2.log)
      const bool pipe_a = (p.tune & 262144) != 0;
      if (pipe_a && nkb_feed > 0) load_split();
      for (int kb = 0; kb < nkb_feed; ++kb) {
        if (!pipe_a) load_split();
        const bool has = kb * 4 + ch * 2 < nks;      // (uniform) this 16-k half carries data
        const int bar_now = a_bar;
        if (has) {
          if (a_round > 0) {
            TCT_TIMED(w1, mbar_wait(bar_ae + 8u * a_slot, (a_round - 1) & 1u));   // the MMAs that read this stage are done
            tc_fence_after();
          }
          const uint32_t ta = t_lane + (uint32_t)(a_slot * 32);
          if (!abl_a) {
            tmem_st16(ta, hi);
            tmem_st16(ta + 16u, lo);
          }
          a_slot += 2;
          while (a_slot >= nas) { a_slot -= nas; ++a_round; }
          a_bar += 2;
          while (a_bar >= abars) a_bar -= abars;
        }
        if (pipe_a && kb + 1 < nkb_feed) load_split();
        if (has) {
          if (!abl_a) tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_af + 8u * bar_now);
        }
      }
#ifdef TCT_PROFILE
      if (warp == CW0) { TCT_PUT(10, w0); TCT_PUT(11, w1); TCT_PUT(13, clock64() - loop0); }
#endif
    } else if (BT) {
      // ===== B warps: the raw tile TMA wrote (SWIZZLE_128B) is split in place: hi over the raw bytes, lo into the twin =====
      const int f4_used = (((PAIR ? hp : n_valid) + 7) & ~7) * 8;   // whole 8-row swizzle atoms that hold stored rows
      constexpr int IT = (TC_BM * 8 + LT - 1) / LT;   // float4 items per thread (bn <= 128 rows x 8)
      RingPos rb = {0, 0};
      const bool trunc_b = (p.tune & 33554432) != 0, abl_b = (p.tune & 536870912) != 0;
      for (int kb = 0; kb < nkb_feed; ++kb) {
        TCT_TIMED(w0, mbar_wait(smem_u32(&bar_b_raw[rb.slot]), rb.round & 1u));
        char* b_hi = smem_b + (size_t)rb.slot * bstage_bytes;
        char* b_lo = b_hi + b_bytes;
        float4 v[IT];
        if (abl_b) {                       // bring-up ablation: barrier protocol only
          mbar_arrive(smem_u32(&bar_b_full[rb.slot]));
          rb.next(nsb);
          continue;
        }
#pragma unroll
        for (int i = 0; i < IT; ++i) {
          const int f = tid + i * LT;
          if (f < f4_used) v[i] = *reinterpret_cast<const float4*>(b_hi + f * 16);
        }
        if (trunc_b) {
          // tune bit 33554432: the raw fp32 tile stays where TMA put it and serves as hi -- the tensor core reads the upper 19
          // bits of a tf32 operand, i.e. hi = x truncated -- and only lo = tf32(x - trunc(x)) is written (exact difference,
          // |lo| < 2^-10 |x|, rounded to the 11 bits the tensor core keeps so that no one-sided error is left). The weight
          // operand keeps its ROUNDED split, so the dropped lo_w * lo_x term (< 2^-21 relative) has no preferred sign.
#pragma unroll
          for (int i = 0; i < IT; ++i) {
            const int f = tid + i * LT;
            if (f < f4_used) {
              float4 l;
              l.x = tf32_hi(v[i].x - __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u));
              l.y = tf32_hi(v[i].y - __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u));
              l.z = tf32_hi(v[i].z - __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u));
              l.w = tf32_hi(v[i].w - __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u));
              *reinterpret_cast<float4*>(b_lo + f * 16) = l;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < IT; ++i) {
            const int f = tid + i * LT;
            if (f < f4_used) tc_split_store(b_hi, b_lo, (uint32_t)(f * 16), v[i]);
          }
        }
        // every thread orders its own generic-proxy stores before the async proxy and arrives (one arrival per warp after a
        // __syncwarp was measured equal, profiles/tma_feed_abl_r2.log: the 256 arrivals are not what the loop costs)
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bar_b_full[rb.slot]));
        rb.next(nsb);
      }
#ifdef TCT_PROFILE
      if (warp == 0) { TCT_PUT(12, w0); TCT_PUT(14, clock64() - loop0); }
#endif
    } else {
      // ===== loader warps: B, global -> registers (TCT_DEPTH k-blocks in flight) -> split -> shared =====
      const Rows RB = resolve(p.B, g);
      const int npb = (n_valid + 3) >> 2;           // patches that hold data (the other lines are never read by the epilogue)
      constexpr int DEPTH = TCT_DEPTH;
      float4 rbuf[DEPTH][NBW];
      auto load_block = [&](int kb, float4 (&qb)[NBW]) {
        const int k0 = kb * BKT;
#pragma unroll
        for (int i = 0; i < NBW; ++i) {
          const int pp = warp + LWB * i;
          qb[i] = (pp < npb) ? tc_patch_load<true, BKT>(RB, pp, lane, n0, p.N, k0, p.K) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      RingPos rb = {0, 0};
      auto store_block = [&](const float4 (&qb)[NBW]) {
        if (rb.round > 0) mbar_wait(smem_u32(&bar_b_empty[rb.slot]), (rb.round - 1) & 1u);
        char* b_hi = smem_b + (size_t)rb.slot * bstage_bytes;
        char* b_lo = b_hi + b_bytes;
#pragma unroll
        for (int i = 0; i < NBW; ++i) {
          const int pp = warp + LWB * i;
          if (pp < npb) tc_split_store(b_hi, b_lo, tc_patch_offset<true, BKT>(pp, lane), qb[i]);
        }
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bar_b_full[rb.slot]));
        rb.next(nsb);
      };
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
        if (d < nkb_feed) load_block(d, rbuf[d]);
      for (int kb0 = 0; kb0 < nkb_feed; kb0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const int kb = kb0 + d;
          if (kb < nkb_feed) {
            store_block(rbuf[d]);
            if (kb + DEPTH < nkb_feed) load_block(kb + DEPTH, rbuf[d]);
          }
        }
      }
    }

    // ===== epilogue (loader + converter warps): this CTA's 128 x bn tile =====
    const int q = warp & 3;
    const int part = warp >> 2;                      // 0 .. (LWB + 8) / 4 - 1
    constexpr int PARTS = (LWB + 8) / 4;
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
#ifdef TCT_PROFILE
    if (tid == 0) TCT_PUT(4, clock64());
#endif
    const int crow = p.cidx ? p.cidx[g] : g;
    float* C = p.cbase + (long long)crow * p.c_gstride + p.c_off;
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
    // PAIR: batch row r sits in accumulator column r (r < hp, the leader's lines) or bh + r - hp (the peer's lines)
    auto chunk_ok = [&](int c) { return c < nch && (PAIR || c * 16 < n_valid); };
    // accumulators to shared ([n][m], through the idle operand stages), then float4 rows of the output
    float* T = reinterpret_cast<float*>(smem);
    for (int c = part; chunk_ok(c); c += PARTS) {
      float gv[16];
      tmem_chunk(c, gv);
#pragma unroll
      for (int j = 0; j < 16; ++j) T[(c * 16 + j) * TC_BM + q * 32 + lane] = gv[j];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");
    int m_rows = p.M - m0;
    m_rows = m_rows > TC_BM ? TC_BM : m_rows;
    const int m4_valid = m_rows >> 2;                // M % 4 == 0 (c_vec)
    const int items = n_valid * (TC_BM / 4);
    const int mq = tid & 31;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == EPI_FWD && p.bias_base && mq < m4_valid) {
      const int brow = p.bias_idx ? p.bias_idx[g] : g;
      b4 = __ldg(reinterpret_cast<const float4*>(p.bias_base + (long long)brow * p.bias_gstride + p.bias_off + m0 + mq * 4));
    }
    constexpr int UNR = 4;
    for (int i0 = tid; i0 < items; i0 += ET * UNR) {
      float4 s4[UNR];
      if (EPI == EPI_BWD_DATA && S) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int i = i0 + u * ET;
          if (i < items && mq < m4_valid)
            s4[u] = __ldg(reinterpret_cast<const float4*>(S + (long long)(n0 + (i >> 5)) * p.ldc + m0 + mq * 4));
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int i = i0 + u * ET;
        if (i < items && mq < m4_valid) {
          const int r = i >> 5;
          const int col = (PAIR && r >= hp) ? r - hp + bh : r;
          float4 o = *reinterpret_cast<const float4*>(T + col * TC_BM + mq * 4);
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

#ifdef TCT_PROFILE
  if (tid == 0) TCT_PUT(5, clock64());
#endif
  tc_fence_before();
  if (PAIR) {
    cluster_sync_all();          // neither CTA leaves (or frees TMEM) while the other may still be read or signalled
    if (warp == MMAW) tmem_dealloc_pair(tmem_d, TC_TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == MMAW) tmem_dealloc(tmem_d, TC_TMEM_COLS);
  }
}

// ---- host side ----
typedef CUresult (*tct_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline tct_encode_fn tct_encoder() {
  static tct_encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<tct_encode_fn>(sym);
  }
  return fn;
}

// 3-D tensor map over a bank of dense row-major matrices: dims (cols, rows, groups), strides (4, ld * 4, gstride * 4) bytes
static inline bool tct_make_map(CUtensorMap* map, const float* base, int cols, int rows, long long ld, long long gstride,
                                int box_cols, int box_rows, bool swizzle128) {
  tct_encode_fn enc = tct_encoder();
  if (!enc || gstride <= 0 || (ld % 4) || (gstride % 4) || !aligned16(base)) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)1 << 20};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)gstride * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// N tile and TMEM plan of the TMA kernel: tiles start every n_per rows (n_per divides the first source's row count when the
// batch is concatenated from two sources), regions of exactly bn = round16(n_per) columns, at least two 32-column A stages
static inline bool tct_plan(int N, int K, int rows0, int* bn_out, int* per_out, int* n_main_out, int* nas_out) {
  const int nks = (K + 7) / 8;
  const int need = (nks + TC_MAX_ACCUM - 1) / TC_MAX_ACCUM;
  const int span = rows0 < N ? rows0 : N;          // tiles may not straddle a multiple of span (the source boundary)
  for (int parts = (span + 127) / 128; parts <= 64; ++parts) {
    const int per = (span + parts - 1) / parts;
    if (rows0 < N && span % per) continue;
    const int bn = (per + 15) / 16 * 16;
    if (TC_TMEM_COLS - (1 + need) * bn >= 64 || ((tc_tune() & 1073741824) && TC_TMEM_COLS - 3 * bn >= 64)) {
      // as many hi*hi regions as still leave two A stages (up to 3, the rotation of the shared-memory-operand kernels: fewer
      // truncating accumulations per region than the cap asks for); tune bit 2097152: only the regions the cap needs
      int n_main = need;
      if ((tc_tune() & 1073741824) && n_main > 2) n_main = 2;   // timing experiment only: more than TC_MAX_ACCUM accumulations per region
      if (!(tc_tune() & 2097152))
        while (n_main < 3 && TC_TMEM_COLS - (2 + n_main) * bn >= 64) ++n_main;
      const int nas = (TC_TMEM_COLS - (1 + n_main) * bn) / 32;
      *bn_out = bn; *per_out = per; *n_main_out = n_main; *nas_out = nas > TCT_MAX_AS ? TCT_MAX_AS : nas;
      return true;
    }
    if (bn <= 16) break;
  }
  return false;
}

template <bool A_KMAJOR, int EPI, int LWB, bool BT, bool PAIR>
static inline void launch_tc_tma_inst(const TcParams& p, const CUtensorMap& mapA, const CUtensorMap& mapAt, const CUtensorMap& mapB0,
                                      const CUtensorMap& mapB1, dim3 grid, size_t smem, cudaStream_t stream, cudaError_t* err) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    *err = cudaFuncSetAttribute(tc_tma_gemm_kernel<A_KMAJOR, EPI, LWB, BT, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)TC_SMEM_BUDGET);
    if (*err != cudaSuccess) return;
  }
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((LWB + 11) * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    *err = cudaLaunchKernelEx(&cfg, tc_tma_gemm_kernel<A_KMAJOR, EPI, LWB, BT, PAIR>, p, mapA, mapAt, mapB0, mapB1);
    count_launch();
    if (*err == cudaSuccess) *err = cudaGetLastError();
    return;
  }
  tc_tma_gemm_kernel<A_KMAJOR, EPI, LWB, BT, PAIR><<<grid, (LWB + 11) * 32, smem, stream>>>(p, mapA, mapAt, mapB0, mapB1);
  count_launch();
  *err = cudaGetLastError();
}

// true: launched (or failed with *err set); false: not applicable (the caller uses the shared-memory-operand kernels).
// A must be the weight matrix of a packed bank: one dense [lines][ld] matrix per group (single RowMap).
// No TMA box may reach beyond the rows of its matrix (see the kernel): ragged last tiles of A use a tail map, B goes by TMA only
// when its boxes (n_per rows x 32 k) tile every source exactly, and K is a multiple of 16.
template <bool A_KMAJOR, int EPI>
static inline bool launch_tc_tma(TcParams p, int G, cudaStream_t stream, cudaError_t* err) {
  if (!(tc_tune() & 131072)) return false;
  if (G <= 0 || p.M <= 0 || p.N <= 0) return false;
  if (!p.c_vec || !p.A.vec || !p.B.vec || p.A.rows0 != 0x7fffffff) return false;
  if (p.K < 2 * TC_BK || p.M < 64) return false;
  // short K (<= 43 k-steps: the 128 x 256 layer of the 2DMG discriminator, the generator trunks, the data gradient of 512 <- 256):
  // a tile is 2 - 8 k-blocks between its set-up and its epilogue, and the two-CTAs-per-SM variant of tc_gemm.cuh (one accumulation
  // region, 256 TMEM columns) overlaps those phases -- measured 0.64 against 0.70 ms on the 512 <- 256 data gradient
  if ((tc_tune() & 1024) && (p.K + 7) / 8 <= TC_MAX_ACCUM) return false;
  // forward: box columns beyond K (K % 32 != 0: the last k-block of K = 784) are zero-filled; measured safe with the operands
  // at the very end of their allocations for K % 16 == 0 (profiles/tma_repro.py), while K = 100 faults there -- the
  // generators' first layer stays on the shared-memory-operand kernels. Data gradient: the k-steps cover K exactly; box columns beyond `in` (in % 128 != 0) belong
  // to the next row of W, for the last row to the bias that follows W in the packed row (out >= 128 floats cover them)
  if (A_KMAJOR ? (p.K % 16 != 0) : (p.K % 8 != 0 || (p.M % TC_BM != 0 && p.K < TC_BM))) return false;
  // the weight matrix as the tensor map sees it: forward A = W[out = M][in = K]; data gradient A = W[out = K][in = M]
  const int w_in = A_KMAJOR ? p.K : p.M, w_out = A_KMAJOR ? p.M : p.K;
  if (p.A.ld != w_in) return false;
  const bool dual = p.B.rows0 < p.N;
  int bn, per, n_main, nas;
  if (!tct_plan(p.N, p.K, dual ? p.B.rows0 : 0x7fffffff, &bn, &per, &n_main, &nas)) return false;
  CUtensorMap mapA, mapAt, mapB0, mapB1;
  const int box_rows = A_KMAJOR ? TC_BM : TC_BK, box_cols = A_KMAJOR ? TC_BK : TC_BM;
  const int tail_rows = w_out % box_rows;
  if (!tct_make_map(&mapA, p.A.base0, w_in, w_out, w_in, p.A.gstride0, box_cols, w_out < box_rows ? w_out : box_rows, A_KMAJOR)) return false;
  if (tail_rows && w_out > box_rows) {
    if (!tct_make_map(&mapAt, p.A.base0, w_in, w_out, w_in, p.A.gstride0, box_cols, tail_rows, A_KMAJOR)) return false;
  } else {
    mapAt = mapA;
  }
  // B by TMA: every source a dense [rows][ld] matrix per group that the n_per-row boxes tile exactly, K % 32 == 0
  // (tune bit 524288 keeps the register path for comparisons)
  bool bt = !(tc_tune() & 524288) && (p.K % 16 == 0);
  const int rows_b0 = dual ? p.B.rows0 : p.N, rows_b1 = dual ? p.N - p.B.rows0 : 0;
  if (rows_b0 % per || rows_b1 % per) bt = false;
  // CTA pairs (tune bit 4194304 switches them off): B by TMA in two halves, at least two M tiles
  // (opt-in, tune bit 8388608: measured slower than the one-CTA kernel -- profiles/tma_pair_r2.log)
  const bool pair = bt && (tc_tune() & 8388608) && (per % 2 == 0) && p.M > TC_BM;
  const int box_b = pair ? per / 2 : per;
  if (bt) bt = tct_make_map(&mapB0, p.B.base0, p.K, rows_b0, p.B.ld, p.B.gstride0, TC_BK, box_b, true);
  if (bt && dual) bt = tct_make_map(&mapB1, p.B.base1, p.K, rows_b1, p.B.ld, p.B.gstride1, TC_BK, box_b, true);
  if (bt && !dual) mapB1 = mapB0;
  if (!bt) { mapB0 = mapA; mapB1 = mapA; }
  p.bn = bn;
  p.n_per = per;
  p.n_main = n_main;
  p.tmem_cols = nas;                   // (re-used field) A stages in TMEM
  p.tune = tc_tune();
  const size_t bstage = 2 * (size_t)bn * TC_BK * 4;
  int nsb = (int)((TC_SMEM_BUDGET - 1024 - (size_t)TCT_NRAW * TCT_RAW_BYTES) / bstage);
  p.n_stages = nsb > TC_MAX_STAGES ? TC_MAX_STAGES : nsb;
  size_t smem = (size_t)TCT_NRAW * TCT_RAW_BYTES + (size_t)p.n_stages * bstage;
  const size_t t_bytes = (size_t)bn * TC_BM * 4;
  if (smem < t_bytes) smem = t_bytes;
  smem += 1024;
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM;
  dim3 grid((p.N + per - 1) / per, m_tiles, G);
  if (pair && bt) grid = dim3((m_tiles + 1) / 2 * 2, (p.N + per - 1) / per, G);   // the pair lies along x
  *err = cudaSuccess;
  static const bool dbg = getenv("CGL_DEBUG_TMA") != nullptr;
  if (dbg)
    fprintf(stderr, "tc_tma<%d,%d> G=%d M=%d N=%d K=%d bn=%d per=%d n_main=%d nas=%d nsb=%d bt=%d dual=%d rows0=%d tail=%d\n",
            (int)A_KMAJOR, EPI, G, p.M, p.N, p.K, bn, per, n_main, nas, p.n_stages, (int)bt + (int)(bt && pair), (int)dual, p.B.rows0, tail_rows);
  if (bt && pair) launch_tc_tma_inst<A_KMAJOR, EPI, 8, true, true>(p, mapA, mapAt, mapB0, mapB1, grid, smem, stream, err);
  else if (bt) launch_tc_tma_inst<A_KMAJOR, EPI, 8, true, false>(p, mapA, mapAt, mapB0, mapB1, grid, smem, stream, err);
  else launch_tc_tma_inst<A_KMAJOR, EPI, 8, false, false>(p, mapA, mapAt, mapB0, mapB1, grid, smem, stream, err);
  return true;
}

}  // namespace cgl
