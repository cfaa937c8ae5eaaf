// Backward of the full-softmax cross-entropy over all item rows (RecBLR.py:99-103, what autograd does through
// nn.CrossEntropyLoss + the logits GEMM) WITHOUT materialising the [users, items] logits or probabilities.
//
//   P = exp(Q E^T - lse) - onehot(pos),   dQ = scale * P E,   dE = scale * P^T Q.
//
// A third instantiation (MODE_FWD) is the FUSED FORWARD of the training step: with a per-user reference m_b (the row maximum,
// or any value within ~80 of it: softmax is shift invariant and bf16 / fp32 keep their relative precision over that range)
// it produces in ONE pass over the table both the softmax denominator s_b = sum_j exp(l_bj - m_b) and the unnormalised
// A_b = sum_j exp(l_bj - m_b) E_j, from which  lse_b = m_b + log s_b  and  dQ_b = scale * (A_b / s_b - E_pos_b)  follow
// without a separate dQ pass: one exponential pass over the [users, items] logits instead of two.
//
// One kernel template serves all of them:  "X rows against a stream of Y tiles"
//   dX[128 rows, D] = sum over Y tiles  P_tile[128, NT] * Y_tile[NT, D],   P_tile from S = X Y_tile^T.
//   dQ: X = Q (rows = users), Y = E (cols = items), softmax statistics per ROW   (TRANSPOSED = false)
//   dE: X = E (rows = items), Y = Q (cols = users), softmax statistics per COLUMN (TRANSPOSED = true)
// so each gradient recomputes the logits once on the tensor cores (4 GEMM passes in total instead of the 3 of a
// materialising implementation; no atomics, deterministic).
//
// Per CTA: warp 0 issues GEMM1, warp 1 issues GEMM2, warp 2 streams Y tiles by TMA (128-byte-swizzled slabs), warps 3-10
// are two groups of softmax warps (thread == row, TMEM lane quarter == warp % 4; group g owns accumulator stage g).  TMEM columns:
//   X  (A of GEMM1, packed bf16 pairs, written once per row block with tcgen05.st)        D/2
//   dX (fp32 accumulator of GEMM2, lives across the whole column loop)                   D
//   S  x NSTG (GEMM1 accumulator, [128, NT] fp32)                                          NSTG * NT
//   P  x NSTG (A of GEMM2, packed bf16 pairs written by the softmax warps)                 NSTG * NT/2
// GEMM1: S = X Y^T       M=128, N=NT, K=D : A from TMEM, B = Y slab as K-major (channels contiguous).
// GEMM2: dX += P Y       M=128, N=D, K=NT : A from TMEM, B = THE SAME shared-memory bytes described as MN-major
//                                            (N = channels contiguous inside a swizzled 128-byte row, K = tile rows).
// GEMM1 of tile t+1 is in flight while the other softmax group turns S(t) into P(t) and GEMM2 of tile t-1 runs.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc05.cuh"
#include "tma_host.cuh"

namespace bdlru {

constexpr int kRows = 128;
constexpr int kBwdMaxStages = 6;
constexpr int kBwdMaxAcc = 3;     // accumulator / P stages = softmax warp groups
constexpr int kBwdSmem = 200 * 1024;
// Which elements of each 32-column chunk take exp2 on the FMA pipe instead of the SFU (bit i = column i; common.cuh
// ex2_mixed).  Overridable at build time for tuning (tools/ce_variants.py).  Measured at 8192 x 1 M x 128
// (profiles/r1_ce_exp_offload.md): the dQ pass only gets slower with any offload (it runs at 0.92 of the sustained
// tensor rate; its softmax warps are issue-bound, not SFU-bound), the dE pass is unchanged within run-to-run noise.
#ifndef BDLRU_CE_DQ_POLY_MASK
#define BDLRU_CE_DQ_POLY_MASK 0u
#endif
#ifndef BDLRU_CE_DE_POLY_MASK
#define BDLRU_CE_DE_POLY_MASK 0u
#endif
constexpr uint32_t kPolyMaskDQ = BDLRU_CE_DQ_POLY_MASK, kPolyMaskDE = BDLRU_CE_DE_POLY_MASK;
// Measured at 8 192 x 10 M x 128 (B200): three groups 46.1 ms vs two groups with 96-column tiles 41.6 ms — the narrower
// GEMM1 tiles cost more than the extra group hides; kept as a build-time variant, off.
#ifndef BDLRU_CE_DE_THREE_GROUPS
#define BDLRU_CE_DE_THREE_GROUPS 0
#endif
constexpr bool kDeThreeGroups = BDLRU_CE_DE_THREE_GROUPS != 0;
#ifndef BDLRU_CE_DE_AUGMENT
#define BDLRU_CE_DE_AUGMENT 1
#endif
constexpr bool kDeAugment = BDLRU_CE_DE_AUGMENT != 0;
// The augmented slab as 32-byte rows (16 K elements, 32-byte swizzle) instead of a 128-byte-swizzled 64-channel slab of
// which one K step was used: 3 KB instead of 12 KB of TMA traffic per tile.
#ifndef BDLRU_CE_DE_AUG_NARROW
#define BDLRU_CE_DE_AUG_NARROW 1
#endif
constexpr bool kAugNarrow = BDLRU_CE_DE_AUG_NARROW != 0;
// Softmax warps: all of S(t) is pulled into registers first and the accumulator stage released BEFORE the exponentials
// (instead of after the last 32-column chunk was loaded, two thirds into them), and P(t) goes back chunk by chunk.
#ifndef BDLRU_CE_EARLY_S_RELEASE
#define BDLRU_CE_EARLY_S_RELEASE 2   // bit per MODE: dE pass only (-3 %); the dQ and fused-forward passes, whose softmax
                                     // warps carry the row statistics, measured 5-10 % slower with it
#endif
constexpr bool kEarlySRelease = BDLRU_CE_EARLY_S_RELEASE != 0;
#ifndef BDLRU_CE_DE_EARLY_X
#define BDLRU_CE_DE_EARLY_X 1
#endif
constexpr bool kDeEarlyX = BDLRU_CE_DE_EARLY_X != 0;

struct BwdParams {
  const void* X;        // [n_x, D] bf16 rows owned by CTAs
  const void* Y;        // [n_y, D] bf16 (also behind the tensor map); read directly for the onehot row of the dQ pass
  long n_x, n_y;        // rows of X, rows of Y (columns of S)
  int D, stages, splits;
  long row_blocks, tiles_total;
  const float* lse;     // [n_users] global logsumexp (MODE_FWD: the per-user reference m_b)
  float* sum_parts;     // MODE_FWD: [splits * NSTG][n_x] partial sums of exp(l - m) (one per column split and softmax group)
  const int64_t* pos;   // [n_users] global item ids
  long n_users;
  long id_offset;       // global id of item row 0 of this shard
  float scale;
  const float* scale_dev;  // optional device scalar multiplied into `scale` (the upstream dL/dloss: no host sync, no extra pass)
  int early_x;          // dE pass: X rows by TMA, early hand-over, dX out by bulk stores (host: D <= 128, two softmax groups, one split)
  int dbg;              // BDLRU_FS_DEBUG & 8: print per-tile phase timings of one softmax warp
  float* out;           // [splits][n_x][D] fp32 (splits == 1: the final gradient)
};

enum { MODE_DQ = 0, MODE_DE = 1, MODE_FWD = 2 };

template <int MODE, int NT, int NSTG>
__global__ void __launch_bounds__(96 + 128 * NSTG, 1)
ce_bwd_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmL,
              const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO, const BwdParams p) {
  constexpr bool TRANSPOSED = MODE == MODE_DE;
  constexpr bool FWD = MODE == MODE_FWD;
  // dE pass: the per-COLUMN statistic -lse_u rides the recompute GEMM as 16 extra K elements — X gets the constant columns
  // (1, 1, 1, 0, ...), every user row of Y the exact three-way bf16 split of -lse_u (one more 64-channel slab per stage from
  // the auxiliary matrix behind tmL, of which one K step is used) — so the accumulator already holds l - lse and the softmax
  // loop needs neither a shared-memory read nor an FFMA per element (it was latency-bound on exactly those).
  constexpr bool AUG = kDeAugment && MODE == MODE_DE;
  // dE pass: a row block is only ~86 tiles long, so the hand-over between row blocks matters.  Measured per block (cycles of a
  // softmax warp, 8192 users): thread == row loads of X 2400 and the thread == row float4 drain of dX 4100 — both touch 32
  // different lines per warp instruction, ~1 cycle per line in the L1 — out of ~112 000 for the whole block.
  // With EARLY_X (D <= 128, two softmax groups):
  //   * the TMA producer fetches the NEXT block's X rows into a swizzled shared-memory stage during the current block;
  //   * the softmax warps copy them into TMEM as soon as the last GEMM1 of the current block has completed — BEFORE they
  //     drain dX — so the first GEMM1s of the next block overlap the drain (dx_empty orders its first GEMM2 after it);
  //   * the drain only moves dX from TMEM into a 64 KB swizzled stage (which contains the X stage) and frees the
  //     accumulator; the TMA producer writes it to global memory with bulk tensor stores while the next block runs (even
  //     as full lines, 64 KB of st.global per block kept the softmax warps ~2900 cycles: the SM's store path).
  constexpr bool EARLY_X = kDeEarlyX && MODE == MODE_DE;
  constexpr bool ESR = kEarlySRelease && ((BDLRU_CE_EARLY_S_RELEASE >> MODE) & 1);   // bit per MODE
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slab = p.D >> 6;
  // ring stage: the 64-channel slabs of a Y tile, then (AUG) NT rows x 32 bytes of -lse splits in 32-byte swizzle
  constexpr uint32_t kAugB = kAugNarrow ? NT * 32 : NT * 128;
  const uint32_t stage_b = (uint32_t)n_slab * (NT * 128) + (AUG ? kAugB : 0u);
  constexpr uint32_t kSlabB = NT * 128;
  constexpr int NCH = NT / 32;
  constexpr int NG = NSTG;                 // softmax warp groups: group g owns accumulator / P stage g
  uint8_t* sY = smem;
  constexpr uint32_t kXSlabB = kRows * 128;   // one 64-channel slab of an X row block
  uint8_t* x_stage = sY + (size_t)p.stages * stage_b;   // early_x: X slabs (1024-byte aligned: stages are)
  // early_x: the drain stage = 4 boxes of 128 rows x 32 fp32 columns; its first half is the X stage
  float* col_lse = reinterpret_cast<float*>(x_stage + (p.early_x ? 4 * kXSlabB : 0));  // [4 * NSTG warps][NT]
  int* col_pos = reinterpret_cast<int*>(col_lse + 4 * NSTG * NT);                       // [4 * NSTG warps][NT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(col_pos + 4 * NSTG * NT);
  uint64_t* y_full = bars;
  uint64_t* y_empty = y_full + kBwdMaxStages;
  uint64_t* s_full = y_empty + kBwdMaxStages;
  uint64_t* s_empty = s_full + kBwdMaxAcc;
  uint64_t* p_full = s_empty + kBwdMaxAcc;
  uint64_t* p_empty = p_full + kBwdMaxAcc;
  uint64_t* x_full = p_empty + kBwdMaxAcc;
  uint64_t* dx_full = x_full + 1;
  uint64_t* dx_empty = dx_full + 1;
  uint64_t* x_ready = dx_empty + 1;
  uint64_t* stage_free = x_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);
  const bool use_pf = EARLY_X && p.early_x;

  // The 8 augmented X columns sit BEHIND the P stages: in front of dX they shifted every accumulator off its 32-column
  // alignment, and each MMA then took ~20-35 % longer (measured issue time per MMA: 70 / 79 cycles instead of 58 / 59).
  const uint32_t x_cols = (uint32_t)(p.D >> 1);
  const uint32_t dx_col = x_cols;
  const uint32_t s_col = dx_col + (uint32_t)p.D;
  const uint32_t p_col = s_col + NSTG * NT;
  const uint32_t xaug_col = p_col + NSTG * (NT / 2);

  if (warp == 2 && lane == 0) {
    tc::prefetch_tensormap(&tmY);
    if (AUG) tc::prefetch_tensormap(&tmL);
    if (EARLY_X) tc::prefetch_tensormap(&tmX);
    if (EARLY_X) tc::prefetch_tensormap(&tmO);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(&y_full[s], 1);
      tc::mbar_init(&y_empty[s], 1);
    }
    for (int b = 0; b < NSTG; ++b) {
      tc::mbar_init(&s_full[b], 1);
      tc::mbar_init(&s_empty[b], 4);
      tc::mbar_init(&p_full[b], 4);
      tc::mbar_init(&p_empty[b], 1);
    }
    tc::mbar_init(x_full, 4 * NG);
    tc::mbar_init(dx_full, 1);
    tc::mbar_init(dx_empty, 4 * NG);
    tc::mbar_init(x_ready, 1);
    tc::mbar_init(stage_free, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const long n_work = p.row_blocks * p.splits;
  // every role walks the same (work item, tile) sequence; g counts tiles globally for ring / stage parities
  if (warp == 2) {
    // ===================================================================== TMA producer
    long g = 0;
    uint32_t wi = 0;
    auto load_x = [&](long wq) {   // X rows of work item wq -> x_stage (rows past n_x arrive as zeros)
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(x_ready, (uint32_t)n_slab * kXSlabB);
        for (int sl = 0; sl < n_slab; ++sl)
          tc::tma_load_2d(x_stage + (size_t)sl * kXSlabB, &tmX, x_ready, sl * 64, (int)((wq % p.row_blocks) * kRows));
      }
      __syncwarp();
    };
    auto store_dx = [&](long wq) {   // the staged dX of work item wq -> global (rows past n_x are clipped by the TMA)
      if (tc::elect_one()) {
        for (int c = 0; c < (p.D >> 5); ++c)
          tc::tma_store_2d(&tmO, x_stage + (size_t)c * kXSlabB, c * 32, (int)((wq % p.row_blocks) * kRows));
        tc::bulk_commit();
        tc::bulk_wait_read();
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(stage_free);
    };
    if (use_pf && blockIdx.x < n_work) load_x(blockIdx.x);
    for (long w = blockIdx.x; w < n_work; w += gridDim.x, ++wi) {
      const int split = (int)(w / p.row_blocks);
      const long t0 = p.tiles_total * split / p.splits, t1 = p.tiles_total * (split + 1) / p.splits;
      // the next block's X is requested once the ring has been filled for this block (so the Y prefetch never waits for
      // it) and the stage is free: X of THIS block copied to TMEM (x_full) and the previous block's drain, which
      // is staged there and has been read by the bulk stores issued right here (dx_empty; it follows that copy)
      const long x_at = t0 + ((t1 - t0) < p.stages ? (t1 - t0) : p.stages) - 1;
      for (long t = t0; t < t1; ++t, ++g) {
        const int s = (int)(g % p.stages);
        const uint32_t ph = (uint32_t)(g / p.stages) & 1u;
        tc::mbar_wait(&y_empty[s], ph ^ 1u);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&y_full[s], stage_b);
          for (int sl = 0; sl < n_slab; ++sl)
            tc::tma_load_2d(sY + (size_t)s * stage_b + (size_t)sl * kSlabB, &tmY, &y_full[s], sl * 64, (int)(t * NT));
          if (AUG) tc::tma_load_2d(sY + (size_t)s * stage_b + (size_t)n_slab * kSlabB, &tmL, &y_full[s], 0, (int)(t * NT));
        }
        __syncwarp();
        if (use_pf && t == x_at) {
          if (wi == 0) {
            tc::mbar_wait(x_full, 0u);
          } else {   // the previous block's dX has been staged (and its X copied out before that)
            tc::mbar_wait(dx_empty, (wi - 1) & 1u);
            store_dx(w - gridDim.x);
          }
          if (w + gridDim.x < n_work) load_x(w + gridDim.x);
        }
      }
    }
    if (use_pf && wi > 0) {   // the last block of this CTA
      tc::mbar_wait(dx_empty, (wi - 1) & 1u);
      store_dx((long)blockIdx.x + (long)(wi - 1) * gridDim.x);
    }
  } else if (warp == 0) {
    // ===================================================================== GEMM1 issuer:  S(t) = X Y(t)^T
    // GEMM1 and GEMM2 are issued by DIFFERENT warps: a tcgen05.mma blocks its issuing thread for about its execution
    // time and every mbarrier wait costs ~300 cycles, so one issuer serialises waits and MMAs (see fullsort.cu).
    constexpr uint32_t idesc1 = tc::idesc_bf16_f32(kRows, NT, 0, 0);
    long g = 0;
    uint32_t wi = 0;
    long long ti_y = 0, ti_s = 0, ti_issue = 0, ti_n = 0;
    for (long w = blockIdx.x; w < n_work; w += gridDim.x, ++wi) {
      const int split = (int)(w / p.row_blocks);
      const long t0 = p.tiles_total * split / p.splits, t1 = p.tiles_total * (split + 1) / p.splits;
      tc::mbar_wait(x_full, wi & 1u);
      tc::fence_after_sync();
      for (long t = t0; t < t1; ++t, ++g) {
        const int s = (int)(g % p.stages);
        const uint32_t ph = (uint32_t)(g / p.stages) & 1u;
        const int b = (int)(g % NSTG);
        const uint32_t bph = (uint32_t)(g / NSTG) & 1u;
        const long long i0 = FS_CLOCK();
        tc::mbar_wait(&y_full[s], ph);
        const long long i1 = FS_CLOCK();
        tc::mbar_wait(&s_empty[b], bph ^ 1u);
        tc::fence_after_sync();
        const long long i2 = FS_CLOCK();
        ti_y += i1 - i0; ti_s += i2 - i1; ++ti_n;
        if (tc::elect_one()) {
          const uint64_t bd0 = tc::smem_desc_sw128(tc::smem_u32(sY + (size_t)s * stage_b), 16, 1024);
          const uint32_t d_tmem = tmem_base + s_col + (uint32_t)b * NT;
          for (int sl = 0; sl < n_slab; ++sl) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              tc::umma_bf16_ts(d_tmem, tmem_base + (uint32_t)(sl * 4 + k4) * 8,
                               bd0 + (uint64_t)((uint32_t)sl * (kSlabB >> 4) + (uint32_t)k4 * 2), idesc1,
                               (uint32_t)((sl | k4) != 0));
          }
          if (AUG)   // the 16 augmented K elements: X columns D/2 .. D/2+7, first K step of the extra slab
            tc::umma_bf16_ts(d_tmem, tmem_base + xaug_col,
                             kAugNarrow ? tc::smem_desc_sw32(tc::smem_u32(sY + (size_t)s * stage_b + (size_t)n_slab * kSlabB), 16, 256)
                                        : bd0 + (uint64_t)((uint32_t)n_slab * (kSlabB >> 4)),
                             idesc1, 1u);
          tc::umma_commit(&s_full[b]);
        }
        __syncwarp();
        ti_issue += FS_CLOCK() - i2;
      }
    }
    if ((FS_DBG(p) & 8) && blockIdx.x == 0 && lane == 0 && ti_n > 0)
      printf("[ce_bwd GEMM1 issuer, mode %d] tiles %lld: wait y_full %lld + wait s_empty %lld + issue %lld cycles per tile\n",
             MODE, ti_n, ti_y / ti_n, ti_s / ti_n, ti_issue / ti_n);
  } else if (warp == 1) {
    // ===================================================================== GEMM2 issuer:  dX += P(t) Y(t)
    const uint32_t idesc2 = tc::idesc_bf16_f32(kRows, p.D, 0, 1);  // B is MN-major in GEMM2
    long g = 0;
    uint32_t wi = 0;
    long long tj_p = 0, tj_issue = 0, tj_n = 0;
    for (long w = blockIdx.x; w < n_work; w += gridDim.x, ++wi) {
      const int split = (int)(w / p.row_blocks);
      const long t0 = p.tiles_total * split / p.splits, t1 = p.tiles_total * (split + 1) / p.splits;
      // the softmax warps of an early-X pass (below) hand the next row block to GEMM1 BEFORE they drain dX: the first
      // GEMM2 of that block (accumulate = 0) must not start until the drain has read the accumulator
      if (EARLY_X && wi > 0) tc::mbar_wait(dx_empty, (wi - 1) & 1u);
      for (long t = t0; t < t1; ++t, ++g) {
        const int s = (int)(g % p.stages);
        const int b = (int)(g % NSTG);
        const uint32_t bph = (uint32_t)(g / NSTG) & 1u;
        const long long j0 = FS_CLOCK();
        tc::mbar_wait(&p_full[b], bph);
        tc::fence_after_sync();
        const long long j1 = FS_CLOCK();
        tj_p += j1 - j0; ++tj_n;
        if (tc::elect_one()) {
          // Y tile as the MN-major B operand: leading (MN) stride = one 64-channel slab, K groups of 8 rows are
          // 1024 bytes apart; one MMA consumes 16 rows = 2048 bytes
          const uint64_t bd0 = tc::smem_desc_sw128(tc::smem_u32(sY + (size_t)s * stage_b), kSlabB, 1024);
          const uint32_t a0 = tmem_base + p_col + (uint32_t)b * (NT / 2);
#pragma unroll
          for (int kk = 0; kk < NT / 16; ++kk)
            tc::umma_bf16_ts(tmem_base + dx_col, a0 + (uint32_t)kk * 8, bd0 + (uint64_t)(kk * (2048 >> 4)), idesc2,
                             (uint32_t)((t != t0) || kk != 0));
          // these commits also cover GEMM1(t): it completed before P(t) could be produced
          tc::umma_commit(&y_empty[s]);
          tc::umma_commit(&p_empty[b]);
          if (t + 1 == t1) tc::umma_commit(dx_full);
        }
        __syncwarp();
        tj_issue += FS_CLOCK() - j1;
      }
    }
    if ((FS_DBG(p) & 8) && blockIdx.x == 0 && lane == 0 && tj_n > 0)
      printf("[ce_bwd GEMM2 issuer, mode %d] tiles %lld: wait p_full %lld + issue %lld cycles per tile\n", MODE, tj_n,
             tj_p / tj_n, tj_issue / tj_n);
  } else {
    // ===================================================================== softmax warps: thread == row
    // NG groups of 4 warps; group g handles the tiles whose accumulator stage is g, so the fixed latencies of one
    // group's hand-offs (barrier waits, TMEM load / store round trips) overlap the other group's arithmetic.
    const int q = warp & 3;
    const int ew = warp - 3;                 // private scratch index
    const int grp = ew >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* my_lse = col_lse + ew * NT;
    int* my_pos = col_pos + ew * NT;
    constexpr float kLog2e = 1.4426950408889634f;
    long g = 0;
    uint32_t wi = 0;
    long long tq_wait_s = 0, tq_math = 0, tq_wait_p = 0, tq_st = 0, tq_n = 0;
    long long tb_x = 0, tb_wait_dx = 0, tb_drain = 0, tb_head = 0, tb_n = 0;
    const long long tq_begin = FS_CLOCK();
    const int xj0 = ((p.D >> 4) * grp) / NG, xj1 = ((p.D >> 4) * (grp + 1)) / NG;  // 8-column groups of X of this warp group
    // this thread's row of X -> TMEM (packed bf16 pairs, channel 2j in the low half); columns split over the groups
    auto x_publish = [&]() {
      if (AUG && grp == 0) {  // augmented K elements D .. D+15 of every X row: 1, 1, 1, 0, ... (bf16 1.0 = 0x3F80)
        const uint32_t ones[8] = {0x3F803F80u, 0x00003F80u, 0u, 0u, 0u, 0u, 0u, 0u};
        tc::tmem_st_32x32_x8(lane_addr + xaug_col, ones);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(x_full);
    };
    auto x_direct = [&](long xr_row) {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.X) + xr_row * p.D);
      for (int j = xj0; j < xj1; ++j) {
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (xr_row < p.n_x) {
          lo = src[2 * j];
          hi = src[2 * j + 1];
        }
        const uint32_t wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tc::tmem_st_32x32_x8(lane_addr + (uint32_t)j * 8, wv);
      }
      x_publish();
    };
    // early_x: this thread's row of the staged X block (TMA 128-byte swizzle: 16-byte chunk c of row r sits at c ^ (r & 7))
    uint32_t xk = 0;   // X loads consumed so far (phase of x_ready)
    auto x_store = [&]() {
      tc::mbar_wait(x_ready, xk & 1u);
      ++xk;
      for (int j = xj0; j < xj1; ++j) {
        const uint8_t* rowp = x_stage + (size_t)(j >> 2) * kXSlabB + (size_t)row * 128;
        const int c0 = (2 * j) & 7;
        const uint4 lo = *reinterpret_cast<const uint4*>(rowp + ((c0 ^ (row & 7)) << 4));
        const uint4 hi = *reinterpret_cast<const uint4*>(rowp + (((c0 + 1) ^ (row & 7)) << 4));
        const uint32_t wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tc::tmem_st_32x32_x8(lane_addr + (uint32_t)j * 8, wv);
      }
      x_publish();
    };
    for (long w = blockIdx.x; w < n_work; w += gridDim.x, ++wi) {
      const long rb = w % p.row_blocks;
      const int split = (int)(w / p.row_blocks);
      const long t0 = p.tiles_total * split / p.splits, t1 = p.tiles_total * (split + 1) / p.splits;
      const long xrow = rb * kRows + row;
      const long w_next = w + gridDim.x;
      const long long b_head0 = FS_CLOCK();
      if (!use_pf) x_direct(xrow);
      else if (wi == 0) x_store();
      const long long b_head1 = FS_CLOCK();
      // row statistics (dQ) / row identity (dE)
      float row_lse2 = 0.f;
      long row_pos = -1;   // dQ: local column of the positive item;  dE: this row's local item index
      if (!TRANSPOSED) {
        if (xrow < p.n_x) {
          row_lse2 = p.lse[xrow] * kLog2e;
          if (!FWD) row_pos = p.pos[xrow] - p.id_offset;
        }
      } else {
        row_pos = xrow;
      }
      float pre_lse[NCH];
      long pre_pos[NCH];
      bool pre_valid = false;
      float rsum[2] = {0.f, 0.f};   // MODE_FWD: this thread's share of sum_j exp(l - m) (two chains for ILP)
      for (long t = t0; t < t1; ++t, ++g) {
        const int b = (int)(g % NSTG);
        if (b != grp) continue;
        const uint32_t bph = (uint32_t)(g / NSTG) & 1u;
        const long cbase = t * NT;
        bool onehot_here = false;
        // With the augmented GEMM (AUG) the dE pass needs NO per-column data in its softmax loop: -lse comes out of the
        // accumulator and the "- onehot" term is applied afterwards by onehot_sub_kernel (dE[pos_b] -= scale * q_b, exact
        // in fp32) — no staging through shared memory, no vote, no second code path per tile.
        if (TRANSPOSED && !AUG) {
          // Column statistics of this tile (lse*log2e, local positive row) -> per-warp scratch.  The values were
          // prefetched into registers while the previous own tile was processed (first own tile: loaded here).
          // (the prefetched values stay RAW in registers — lse unscaled, pos as loaded — and are converted only here, one
          // tile later: converting at prefetch time made the very next instruction wait for the global load, ncu
          // long-scoreboard stalls on the scaling FMUL)
          if (!pre_valid) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              const long u = cbase + lane + 32 * k;
              const bool ok = u < p.n_users;
              if (!AUG) pre_lse[k] = ok ? p.lse[u] : INFINITY;
              pre_pos[k] = ok ? p.pos[u] : -1;
            }
          }
          __syncwarp();
          const int row_lo = (int)(rb * kRows) + q * 32;  // this warp's 32 rows
          bool mine = false;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            const int lp = pre_pos[k] < 0 ? -1 : (int)(pre_pos[k] - p.id_offset);
            if (!AUG) my_lse[lane + 32 * k] = pre_lse[k] * kLog2e;
            my_pos[lane + 32 * k] = lp;
            mine |= (lp >= row_lo && lp < row_lo + 32);
          }
          onehot_here = __any_sync(0xffffffffu, mine);  // rare: some column's positive item is one of this warp's rows
          __syncwarp();
          // prefetch for the next own tile of this work item (latency hidden behind this tile's arithmetic)
          const long nt = t + NG;
          pre_valid = nt < t1;
          if (pre_valid) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              const long u = nt * NT + lane + 32 * k;
              const bool ok = u < p.n_users;
              if (!AUG) pre_lse[k] = ok ? p.lse[u] : INFINITY;
              pre_pos[k] = ok ? p.pos[u] : -1;
            }
          }
        }
        const bool tail = !TRANSPOSED && (cbase + NT > p.n_y);  // last tile: columns beyond the table contribute 0
        uint32_t packed[NT / 2];
        const long long k0 = FS_CLOCK();
        tc::mbar_wait(&s_full[b], bph);
        tc::fence_after_sync();
        const long long k1 = FS_CLOCK();
        uint32_t raw_all[ESR ? NCH : 1][32];
        if (ESR) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc::tmem_ld_32x32(lane_addr + s_col + (uint32_t)b * NT + c * 32, raw_all[c]);
          tc::tmem_ld_wait();
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&s_empty[b]);   // S(t) is in registers: GEMM1 of tile t + NSTG may overwrite the stage
          tc::mbar_wait(&p_empty[b], bph ^ 1u);          // GEMM2 of the tile that used this P stage has completed
          tc::fence_after_sync();
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t (&raw)[32] = raw_all[ESR ? c : 0];
          if (!ESR) {
            tc::tmem_ld_32x32(lane_addr + s_col + (uint32_t)b * NT + c * 32, raw);
            tc::tmem_ld_wait();
            if (c == NCH - 1) {
              tc::fence_before_sync();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(&s_empty[b]);
            }
          }
          // P = exp(S - lse) as packed bf16 pairs.  dQ pass: the "- onehot" term is NOT applied here (it would put
          // per-element 64-bit compares into this MUFU-bound loop); it is subtracted as one row of Y when the row block
          // is written out.  dE pass: applied here, but only in the rare tiles that contain one of this warp's rows.
          auto chunk_math = [&](auto check_tag) {
            constexpr bool CHECK = decltype(check_tag)::value;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float pr[4];
              // dE pass: the four columns' statistics in ONE 128-bit broadcast read (the per-element LDS was a quarter
              // of this loop's issue slots)
              float cl[4] = {0.f, 0.f, 0.f, 0.f};
              if (TRANSPOSED && !AUG) {
                const float4 l4 = *reinterpret_cast<const float4*>(my_lse + c * 32 + i);
                cl[0] = l4.x; cl[1] = l4.y; cl[2] = l4.z; cl[3] = l4.w;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int col = c * 32 + i + u;
                const float sv = __uint_as_float(raw[i + u]);
                if (TRANSPOSED) {
                  // AUG: the accumulator already holds l - lse (columns past n_users hold 0 and multiply zero rows of Y)
                  pr[u] = ex2_mixed<TRANSPOSED ? kPolyMaskDE : kPolyMaskDQ, 3>(AUG ? sv * kLog2e : fmaf(sv, kLog2e, -cl[u]), i + u);
                  if (CHECK && my_pos[col] == (int)row_pos) pr[u] -= 1.f;
                } else {
                  pr[u] = ex2_mixed<TRANSPOSED ? kPolyMaskDE : kPolyMaskDQ, 3>(fmaf(sv, kLog2e, -row_lse2), i + u);
                  if (CHECK && cbase + col >= p.n_y) pr[u] = 0.f;
                }
              }
              if (FWD) {
                rsum[0] += pr[0] + pr[2];
                rsum[1] += pr[1] + pr[3];
              }
              const __nv_bfloat162 h0 = __floats2bfloat162_rn(pr[0], pr[1]);
              const __nv_bfloat162 h1 = __floats2bfloat162_rn(pr[2], pr[3]);
              packed[(c * 32 + i) >> 1] = *reinterpret_cast<const uint32_t*>(&h0);
              packed[((c * 32 + i) >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            }
          };
          if (TRANSPOSED ? onehot_here : tail) chunk_math(std::true_type{});
          else chunk_math(std::false_type{});
          if (ESR) {
#pragma unroll
            for (int j = 2 * c; j < 2 * c + 2; ++j) {
              const uint32_t wv[8] = {packed[8 * j], packed[8 * j + 1], packed[8 * j + 2], packed[8 * j + 3],
                                      packed[8 * j + 4], packed[8 * j + 5], packed[8 * j + 6], packed[8 * j + 7]};
              tc::tmem_st_32x32_x8(lane_addr + p_col + (uint32_t)b * (NT / 2) + (uint32_t)j * 8, wv);
            }
          }
        }
        const long long k2 = FS_CLOCK();
        if (!ESR) {
          tc::mbar_wait(&p_empty[b], bph ^ 1u);   // GEMM2 of the tile that used this P stage has completed
          tc::fence_after_sync();
        }
        const long long k3 = FS_CLOCK();
#pragma unroll
        for (int j = 0; j < (ESR ? 0 : NT / 16); ++j) {
          const uint32_t wv[8] = {packed[8 * j], packed[8 * j + 1], packed[8 * j + 2], packed[8 * j + 3],
                                  packed[8 * j + 4], packed[8 * j + 5], packed[8 * j + 6], packed[8 * j + 7]};
          tc::tmem_st_32x32_x8(lane_addr + p_col + (uint32_t)b * (NT / 2) + (uint32_t)j * 8, wv);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&p_full[b]);
        const long long k4 = FS_CLOCK();
        tq_wait_s += k1 - k0; tq_math += k2 - k1; tq_wait_p += k3 - k2; tq_st += k4 - k3; ++tq_n;
      }
      const long long b0 = FS_CLOCK();
      if (use_pf && w_next < n_work) {
        // every GEMM1 of this row block has completed (the last tile's accumulator is full): X may be replaced
        const long gl = g - 1;
        tc::mbar_wait(&s_full[(int)(gl % NSTG)], (uint32_t)(gl / NSTG) & 1u);
        tc::fence_after_sync();
        x_store();
      }
      // row block finished: dX accumulator -> global (scaled); 32-column chunks split over the groups
      // the X stage doubles as the drain's transpose buffer: every softmax warp has finished copying X out of it
      if (use_pf) asm volatile("bar.sync 1, %0;" ::"n"(128 * NG) : "memory");
      const long long b1 = FS_CLOCK();
      tc::mbar_wait(dx_full, wi & 1u);
      tc::fence_after_sync();
      const long long b2 = FS_CLOCK();
      float* orow = p.out + ((size_t)split * p.n_x + xrow) * p.D;
      const float out_scale = FWD ? 1.f : p.scale * (p.scale_dev ? __ldg(p.scale_dev) : 1.f);
      // dQ pass: "- onehot" = minus the positive item's row of Y, applied once, by the split that owns that column
      if (FWD && xrow < p.n_x) p.sum_parts[((size_t)split * NG + grp) * p.n_x + xrow] = rsum[0] + rsum[1];
      const bool sub_pos = MODE == MODE_DQ && xrow < p.n_x && row_pos >= t0 * NT && row_pos < t1 * NT && row_pos < p.n_y;
      const __nv_bfloat16* yrow = reinterpret_cast<const __nv_bfloat16*>(p.Y) + (sub_pos ? row_pos : 0) * p.D;
      if (use_pf) {
        // thread == row moves its 32-column chunks into the stage (box c = 128 rows x 128 bytes, swizzled like a TMA
        // box: conflict-free); the producer warp stores the boxes once every softmax warp has arrived on dx_empty
        if (wi > 0) tc::mbar_wait(stage_free, (wi - 1) & 1u);   // the previous block's bulk stores have read the stage
        for (int c = grp; c < (p.D >> 5); c += NG) {
          uint32_t acc[32];
          tc::tmem_ld_32x32(lane_addr + dx_col + (uint32_t)c * 32, acc);
          tc::tmem_ld_wait();
          uint8_t* box = x_stage + (size_t)c * kXSlabB + (size_t)row * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(box + ((i ^ (row & 7)) << 4)) =
                make_float4(__uint_as_float(acc[4 * i]) * out_scale, __uint_as_float(acc[4 * i + 1]) * out_scale,
                            __uint_as_float(acc[4 * i + 2]) * out_scale, __uint_as_float(acc[4 * i + 3]) * out_scale);
        }
        tc::fence_proxy_async();   // generic writes -> visible to the bulk stores (async proxy)
      } else
      for (int c = grp; c < (p.D >> 5); c += NG) {
        uint32_t acc[32];
        tc::tmem_ld_32x32(lane_addr + dx_col + (uint32_t)c * 32, acc);
        tc::tmem_ld_wait();
        if (xrow < p.n_x) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              o[e] = __uint_as_float(acc[i + e]);
              if (sub_pos) o[e] -= __bfloat162float(yrow[c * 32 + i + e]);
              o[e] *= out_scale;
            }
            *reinterpret_cast<float4*>(orow + c * 32 + i) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
      tc::fence_before_sync();
      if (EARLY_X) {   // dX has been read: the next row block's first GEMM2 may overwrite it
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(dx_empty);
      }
      const long long b3 = FS_CLOCK();
      tb_x += b1 - b0; tb_wait_dx += b2 - b1; tb_drain += b3 - b2; tb_head += b_head1 - b_head0; ++tb_n;
    }
    if ((FS_DBG(p) & 8) && blockIdx.x == 0 && warp == 3 && lane == 0 && tq_n > 0)
      printf("[ce_bwd softmax warp, transposed=%d] own tiles %lld: total/own-tile %lld = wait s_full %lld + ld+math %lld + "
             "wait p_empty %lld + st+arrive %lld cycles\n", (int)TRANSPOSED, tq_n, (FS_CLOCK() - tq_begin) / tq_n,
             tq_wait_s / tq_n, tq_math / tq_n, tq_wait_p / tq_n, tq_st / tq_n);
    if ((FS_DBG(p) & 8) && blockIdx.x == 0 && (warp == 3 || warp == 7) && lane == 0 && tb_n > 0)
      printf("[ce_bwd row-block hand-over, transposed=%d, warp %d] blocks %lld: X fetch/store at head %lld, wait last GEMM1 + X "
             "store %lld, wait dx_full %lld, drain %lld cycles per block\n", (int)TRANSPOSED, warp, tb_n, tb_head / tb_n,
             tb_x / tb_n, tb_wait_dx / tb_n, tb_drain / tb_n);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// out[r][d] = sum_s part[s][r][d]
__global__ void sum_partials_kernel(const float4* __restrict__ part, long n4, int splits, float4* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = part[i];
  for (int s = 1; s < splits; ++s) {
    const float4 b = part[(size_t)s * n4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = a;
}

// dE[pos_b - id_offset, :] -= scale * q_b for every user whose positive item lives in this shard: the "- onehot" term of
// dE = scale * (P - onehot)^T Q, applied after the augmented dE pass has written scale * P^T Q.  One warp per user,
// vector red.global.add (several users may share a positive item).
__global__ void __launch_bounds__(256) onehot_sub_kernel(const __nv_bfloat16* __restrict__ Q, const int64_t* __restrict__ pos,
                                                         long n_users, int D, long id_offset, long n_rows, float scale,
                                                         const float* __restrict__ scale_dev, float* __restrict__ dE) {
  const long u = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (u >= n_users) return;
  const long r = pos[u] - id_offset;
  if (r < 0 || r >= n_rows) return;
  const float sc = -scale * (scale_dev ? __ldg(scale_dev) : 1.f);
  for (int v = lane; v < D / 4; v += 32) {
    float q[4];
    IO<__nv_bfloat16>::load(Q + u * D + v * 4, q);
    atomicAdd(reinterpret_cast<float4*>(dE + r * D + v * 4), make_float4(sc * q[0], sc * q[1], sc * q[2], sc * q[3]));
  }
}

// L_aug[u, 0..2] = exact three-way bf16 split of -lse[u], the rest of the row zero; a row is kAugNarrow ? 32 : 128 bytes
__global__ void lse_aug_kernel(const float* __restrict__ lse, long n_users, __nv_bfloat16* __restrict__ out) {
  constexpr int CH = kAugNarrow ? 2 : 8;   // 16-byte chunks per row
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (user, chunk)
  if (i >= n_users * CH) return;
  const long u = i / CH;
  uint4 w = make_uint4(0, 0, 0, 0);
  if (i % CH == 0) {
    const float v = -lse[u];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    w.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
    w.y = (uint32_t)__bfloat16_as_ushort(lo);
  }
  reinterpret_cast<uint4*>(out)[i] = w;
}

// ----------------------------------------------------------------------------- host side
struct BwdPlan {
  int NT, NSTG, stages, splits, grid;
  bool early_x;
  long row_blocks, tiles;
  size_t smem;
};

// three_groups: the dE pass at D <= 128 runs 64-column tiles with THREE accumulator stages / softmax warp groups (TMEM
// D/2 + D + 3*64 + 3*32 = 480 columns): its softmax warps are latency-bound (ncu: issue slots 35 % busy, tensor and XU
// pipes ~52 %), so a third group in flight hides more of the per-tile hand-off latency than wider tiles gain.
static void bwd_plan(long n_x, long n_y, int D, BwdPlan* pl, bool three_groups = false, bool aug = false,
                     bool early_x = false) {
  pl->NT = D <= 128 ? 96 : 64;
  pl->NSTG = D <= 192 ? 2 : 1;
  if (three_groups && D <= 128) {
    pl->NT = 64;
    pl->NSTG = 3;
  }
  pl->row_blocks = (n_x + kRows - 1) / kRows;
  pl->tiles = (n_y + pl->NT - 1) / pl->NT;
  // column splits per row block: the smallest count whose work items fill the persistent grid to >= 95 % in whole waves
  // (e.g. 64 row blocks on 148 SMs: 2 splits leave 20 SMs idle, 9 splits = 576 items = 3.9 waves), at least 4 tiles each
  const long sm = sm_count();
  const long max_s = pl->tiles / 4 > 0 ? pl->tiles / 4 : 1;
  long splits = 1;
  double best = 0.0;
  for (long s = 1; s <= max_s && s <= 64; ++s) {
    const long n = pl->row_blocks * s;
    const double util = (double)n / (double)(((n + sm - 1) / sm) * sm);
    if (util > best + 1e-9) { best = util; splits = s; }
    if (util >= 0.95) break;
  }
  // aug: the -lse columns ride along (32-byte rows, or one more 64-channel slab)
  const size_t stage = (size_t)(D / 64) * pl->NT * 128 + (aug ? (size_t)pl->NT * (kAugNarrow ? 32 : 128) : 0);
  // early_x (dE pass): 64 KB stage for the next X block / the drained dX, out of the full 227 KB instead of the 200 KB
  // budget (4 ring stages instead of 5: no measurable difference, 3 cost 2.5 %).  The bulk stores clip at n_x, which is
  // only the right bound of the output when there is a single column split.
  pl->early_x = early_x && pl->NSTG == 2 && D <= 128 && splits == 1;
  const size_t x_stage = pl->early_x ? (size_t)4 * kRows * 128 : 0;   // 4 boxes of 128 rows x 128 bytes
  const size_t budget = x_stage ? (size_t)226 * 1024 : (size_t)kBwdSmem;
  int stages = (int)((budget - 1024 - 8 * pl->NSTG * pl->NT * 4 - 512 - x_stage) / stage);
  pl->stages = stages > kBwdMaxStages ? kBwdMaxStages : stages;
#ifdef BDLRU_CE_DE_MAX_STAGES
  if (early_x && pl->stages > BDLRU_CE_DE_MAX_STAGES) pl->stages = BDLRU_CE_DE_MAX_STAGES;
#endif
  pl->splits = (int)splits;
  const long n_work = pl->row_blocks * splits;
  pl->grid = (int)(n_work < sm_count() ? n_work : sm_count());
  pl->smem = 1024 + (size_t)pl->stages * stage + 8 * pl->NSTG * pl->NT * 4 + 512 + x_stage;
}

template <int TR>
static int bwd_launch(const BwdPlan& pl, const CUtensorMap& my, const CUtensorMap& ml, const CUtensorMap& mx,
                      const CUtensorMap& mo, const BwdParams& p, cudaStream_t st) {
#define BWD_CASE(NTv, NSv)                                                                                      \
  if (pl.NT == NTv && pl.NSTG == NSv) {                                                                         \
    BDLRU_CUDA(cudaFuncSetAttribute(ce_bwd_kernel<TR, NTv, NSv>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                    (int)pl.smem));                                                             \
    ce_bwd_kernel<TR, NTv, NSv><<<pl.grid, 96 + 128 * NSv, pl.smem, st>>>(my, ml, mx, mo, p);                                       \
    BDLRU_LAUNCHED();                                                                                           \
    return BDLRU_OK;                                                                                            \
  }
  BWD_CASE(96, 2) BWD_CASE(64, 2) BWD_CASE(64, 1)
  if constexpr (TR == MODE_DE) { BWD_CASE(64, 3) }
#undef BWD_CASE
  set_error("fullsort_ce_bwd: no kernel for NT=%d NSTG=%d", pl.NT, pl.NSTG);
  return BDLRU_ERR_UNSUPPORTED;
}

// one gradient: X rows against Y columns; result (scaled) in `grad` [n_x, D]
template <int TR>
static int bwd_one(const void* X, long n_x, const void* Y, long n_y, int D, const float* lse, const int64_t* pos,
                   long n_users, long id_offset, float scale, const float* scale_dev, float* grad, float* scratch,
                   __nv_bfloat16* laug, cudaStream_t st) {
  BwdPlan pl;
  constexpr bool AUG = TR == MODE_DE && kDeAugment;
  bwd_plan(n_x, n_y, D, &pl, TR == MODE_DE && kDeThreeGroups, AUG, TR == MODE_DE && kDeEarlyX);
  CUtensorMap my, ml;
  int rc = make_rows_map(&my, Y, n_y, D, pl.NT);
  if (rc) return rc;
  ml = my;
  if (AUG) {   // Y = Q here: one row of -lse splits per user, streamed as a third slab of every stage
    lse_aug_kernel<<<(unsigned)((n_users * (kAugNarrow ? 2 : 8) + 255) / 256), 256, 0, st>>>(lse, n_users, laug);
    BDLRU_LAUNCHED();
    if ((rc = kAugNarrow ? make_rows16_map(&ml, laug, n_users, pl.NT) : make_rows_map(&ml, laug, n_users, 64, pl.NT)))
      return rc;
  }
  BwdParams p = {};
  p.X = X; p.Y = Y; p.n_x = n_x; p.n_y = n_y; p.D = D; p.stages = pl.stages; p.splits = pl.splits;
  p.row_blocks = pl.row_blocks; p.tiles_total = pl.tiles;
  p.lse = lse; p.pos = pos; p.n_users = n_users; p.id_offset = id_offset; p.scale = scale; p.scale_dev = scale_dev;
  p.dbg = tuning_env("BDLRU_FS_DEBUG");
  p.out = pl.splits > 1 ? scratch : grad;
  p.early_x = pl.early_x ? 1 : 0;
  CUtensorMap mx = my, mo = my;
  if (pl.early_x && (rc = make_rows_map(&mx, X, n_x, D, kRows))) return rc;
  if (pl.early_x && (rc = make_rows_map_f32(&mo, p.out, n_x, D, kRows))) return rc;
  if ((rc = bwd_launch<TR>(pl, my, ml, mx, mo, p, st))) return rc;
  if (pl.splits > 1) {
    const long n4 = n_x * D / 4;
    sum_partials_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(scratch), n4,
                                                                      pl.splits, reinterpret_cast<float4*>(grad));
    BDLRU_LAUNCHED();
  }
  if (AUG) {   // here X = E (n_x item rows), Y = Q (n_users rows)
    onehot_sub_kernel<<<(unsigned)((n_users * 32 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(Y), pos, n_users, D, id_offset, n_x, scale, scale_dev, grad);
    BDLRU_LAUNCHED();
  }
  return BDLRU_OK;
}

// out[u] = sum_j parts[j][u]
__global__ void rowsum_merge_kernel(const float* __restrict__ parts, long n, int n_parts, float* __restrict__ out) {
  const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  float a = 0.f;
  for (int j = 0; j < n_parts; ++j) a += parts[(size_t)j * n + u];
  out[u] = a;
}

static size_t round128(size_t x) { return (x + 127) & ~(size_t)127; }

size_t ce_bwd_workspace_bytes(long n_users, long n_rows, int D) {
  BwdPlan a, b;
  bwd_plan(n_users, n_rows, D, &a);
  bwd_plan(n_rows, n_users, D, &b, kDeThreeGroups, kDeAugment, kDeEarlyX);
  const size_t wa = a.splits > 1 ? (size_t)a.splits * n_users * D * 4 : 0;
  const size_t wb = b.splits > 1 ? (size_t)b.splits * n_rows * D * 4 : 0;
  // + the [n_users, 64] bf16 matrix of -lse splits the dE pass streams next to Q
  return round128(wa > wb ? wa : wb) + (kDeAugment ? (size_t)n_users * 128 : 0);
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_fullsort_ce_bwd(const void* Q, const void* E, const int64_t* pos, const float* lse,
                                               float scale, const float* scale_dev, int64_t n_users, int64_t n_rows, int D,
                                               int64_t id_offset,
                                               float* dQ, float* dE, void* workspace, size_t workspace_bytes,
                                               void* stream) {
  BDLRU_REQUIRE(Q && E && pos && lse, "fullsort_ce_bwd: null input");
  BDLRU_REQUIRE(dQ || dE, "fullsort_ce_bwd: both gradients null");
  BDLRU_REQUIRE(n_users >= 1 && n_rows >= 1, "fullsort_ce_bwd: bad sizes n_users=%ld n_rows=%ld", (long)n_users, (long)n_rows);
  BDLRU_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256, "fullsort_ce_bwd: D=%d must be a multiple of 64 in [64, 256]", D);
  BDLRU_REQUIRE(aligned(Q, 16) && aligned(E, 16) && aligned(dQ, 16) && aligned(dE, 16),
                "fullsort_ce_bwd: Q/E/dQ/dE must be 16-byte aligned");
  const size_t need = ce_bwd_workspace_bytes(n_users, n_rows, D);
  BDLRU_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), "fullsort_ce_bwd: workspace %zu < %zu bytes",
                workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc;
  __nv_bfloat16* laug = kDeAugment ? reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(workspace) + need -
                                                                   (size_t)n_users * 128)
                                   : nullptr;
  if (dQ && (rc = bwd_one<MODE_DQ>(Q, n_users, E, n_rows, D, lse, pos, n_users, id_offset, scale, scale_dev, dQ,
                                   reinterpret_cast<float*>(workspace), nullptr, st)))
    return rc;
  if (dE && (rc = bwd_one<MODE_DE>(E, n_rows, Q, n_users, D, lse, pos, n_users, id_offset, scale, scale_dev, dE,
                                   reinterpret_cast<float*>(workspace), laug, st)))
    return rc;
  return BDLRU_OK;
}

extern "C" BDLRU_API size_t bdlru_fullsort_ce_fwd_dq_workspace_bytes(int64_t n_users, int64_t n_rows, int D) {
  if (D % 64 != 0 || D < 64 || D > 256 || n_users < 1 || n_rows < 1) return 0;
  BwdPlan a;
  bwd_plan(n_users, n_rows, D, &a);
  const size_t acc = a.splits > 1 ? (size_t)a.splits * n_users * D * 4 : 0;
  return acc + (size_t)a.splits * a.NSTG * n_users * 4;
}

extern "C" BDLRU_API int bdlru_fullsort_ce_fwd_dq(const void* Q, const void* E, const float* ref, int64_t n_users,
                                                  int64_t n_rows, int D, float* acc, float* row_sumexp, void* workspace,
                                                  size_t workspace_bytes, void* stream) {
  BDLRU_REQUIRE(Q && E && ref && acc && row_sumexp, "fullsort_ce_fwd_dq: null pointer");
  BDLRU_REQUIRE(n_users >= 1 && n_rows >= 1, "fullsort_ce_fwd_dq: bad sizes n_users=%ld n_rows=%ld", (long)n_users, (long)n_rows);
  BDLRU_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256, "fullsort_ce_fwd_dq: D=%d must be a multiple of 64 in [64, 256]", D);
  BDLRU_REQUIRE(aligned(Q, 16) && aligned(E, 16) && aligned(acc, 16), "fullsort_ce_fwd_dq: Q/E/acc must be 16-byte aligned");
  const size_t need = bdlru_fullsort_ce_fwd_dq_workspace_bytes(n_users, n_rows, D);
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "fullsort_ce_fwd_dq: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  BwdPlan pl;
  bwd_plan(n_users, n_rows, D, &pl);
  CUtensorMap my;
  int rc = make_rows_map(&my, E, n_rows, D, pl.NT);
  if (rc) return rc;
  const size_t acc_bytes = pl.splits > 1 ? (size_t)pl.splits * n_users * D * 4 : 0;
  float* scratch = reinterpret_cast<float*>(workspace);
  BwdParams p = {};
  p.X = Q; p.Y = E; p.n_x = n_users; p.n_y = n_rows; p.D = D; p.stages = pl.stages; p.splits = pl.splits;
  p.row_blocks = pl.row_blocks; p.tiles_total = pl.tiles;
  p.lse = ref; p.pos = nullptr; p.n_users = n_users; p.id_offset = 0; p.scale = 1.f; p.scale_dev = nullptr;
  p.dbg = tuning_env("BDLRU_FS_DEBUG");
  p.sum_parts = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + acc_bytes);
  p.out = pl.splits > 1 ? scratch : acc;
  if ((rc = bwd_launch<MODE_FWD>(pl, my, my, my, my, p, st))) return rc;
  if (pl.splits > 1) {
    const long n4 = n_users * D / 4;
    sum_partials_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(scratch), n4,
                                                                      pl.splits, reinterpret_cast<float4*>(acc));
    BDLRU_LAUNCHED();
  }
  rowsum_merge_kernel<<<(unsigned)((n_users + 255) / 256), 256, 0, st>>>(p.sum_parts, n_users, pl.splits * pl.NSTG,
                                                                        row_sumexp);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
