// Full-sort scoring (RecBLR.py:114-122 + RecBole's mask/top-k) and the forward of the full-softmax CE
// (RecBLR.py:99-103) as ONE warp-specialised tcgen05 kernel family: S = Q E^T is accumulated in TMEM and consumed
// straight out of TMEM by the epilogue (streaming per-user top-k, or online max / sum-exp), so the [users, items]
// logits never exist in memory.
//
// Tiling.  A CTA owns UB (1 or 2) blocks of 128 users, kept resident in shared memory as the A operand
// (K-major, 128-byte swizzle, one 16 KB slab per 64 channels), and walks a contiguous range of 128-item tiles of E
// streamed by TMA through a multi-stage ring (B operand, same layout).  Per tile it issues UB x D/16
// tcgen05.mma (M=128, N=128, K=16, bf16 x bf16 -> fp32) into one of two TMEM accumulator sets, so the epilogue of
// tile t overlaps the MMAs of tile t+1.  Using both user blocks against each E tile halves the L2->SM operand
// traffic per flop (at D=128 one user block alone would need ~10 TB/s of L2 bandwidth to keep the tensor pipe fed).
//
// Warp roles: warps 0-1 = MMA issuers (alternate tiles; warp 0 also owns the TMEM allocation), warp 2 = TMA producer,
// warps 3.. = epilogue,
// one warp per 32 TMEM lanes per user block (a warp may only touch TMEM lanes 32*(warp%4)..+31): thread == user.
//
// Top-k epilogue: every thread keeps its user's K best (score, id) sorted in REGISTERS (K = 10/16/20/32 by template)
// and the K-th score as threshold.  A 32-score chunk is first reduced to its maximum; only if some lane's maximum beats
// its threshold does the warp enter insert rounds: each lane with a candidate extracts its chunk maximum (first index
// among equals), inserts it with a branch-free unrolled shift, knocks it out and recomputes its maximum, until no lane
// has a candidate left.  Lanes insert together, so a round costs ~100 + 5K issue slots for up to 32 insertions (the
// expected number of insertions per user over n items is ~K ln(n/K); a one-lane-at-a-time list in shared memory made
// the epilogue 10x slower than the MMAs).  An item is inserted only if STRICTLY better than the K-th and after all
// entries with an equal score; chunks arrive in increasing id order, so ties resolve to the lowest id.
// The grid splits the item range S ways; a second kernel merges the S (or, multi-GPU, G) sorted lists per user by
// (score desc, id asc).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"
#include "tma_host.cuh"

namespace bdlru {

constexpr int kTile = 128;                 // users per block (TMEM lanes) and items per E tile (MMA N)
constexpr int kMaxSmem = 227 * 1024;
constexpr int kMaxStages = 8;
// CE forward: elements of each 32-column chunk whose exp2 runs on the FMA pipe (common.cuh ex2_mixed, degree 4)
#ifndef BDLRU_CE_FWD_POLY_MASK
#define BDLRU_CE_FWD_POLY_MASK 0u
#endif
constexpr uint32_t kPolyMaskFwd = BDLRU_CE_FWD_POLY_MASK;

enum { MODE_TOPK = 0, MODE_CE = 1, MODE_MAX = 2 };  // MAX: row maxima only (the reference of the fused CE forward)

struct FsParams {
  int D, k, stages, splits, n_ug;
  int dbg;  // BDLRU_FS_DEBUG, honoured ONLY in -DBDLRU_TUNING builds (never in the shipped library): 1 = epilogue skips
            // loads+arithmetic, 2 = MMA warp skips the MMAs, 4 = no TMA loads, 8 = clock64 phase print, 16 = no threshold sharing
  const void* Q;              // [n_users, D] bf16 row-major
  long n_users, n_rows;       // rows of Q, rows of this E shard
  long id_offset, mask_local; // global id of E row 0; LOCAL row to exclude (-1: none)
  long tiles_total;
  int tile_stride;            // MODE_MAX: walk every tile_stride-th tile only (sampled maximum); 1 elsewhere
  // top-k partials [n_users][splits][k]
  float* part_scores;
  int* part_ids;
  unsigned* shared_thr;       // [n_users] order-preserving encoding of the best K-th score seen by ANY split (0 = none)
  // CE partials [n_users][ce_parts = splits * softmax warp sets] and direct outputs
  int ce_parts;
  float* part_max;
  float* part_sum;
  const int64_t* pos;
  float* pos_logit;
};

struct FsPlan {
  int UB, NT, NSTG, stages, splits, n_ug, threads;
  size_t smem;
  long tiles_total;
};

// One plan for (n_users, n_rows, D): used by the launch AND by the workspace query, so both agree.
// TMEM budget (512 columns): UB*D/2 for Q + 2*UB*NT for the double-buffered accumulators -> NT = 96 up to D = 128,
// 64 above.
static bool fs_plan(long n_users, long n_rows, int D, int k, int mode, FsPlan* pl) {
  (void)k;
  const int n_slab = D / 64;
  const int UB = n_users > kTile ? 2 : 1;
  // BDLRU_FS_NT=64 (tuning): 64-item tiles with 3 accumulator stages instead of 96-item tiles with 2
  static const int force_nt = tuning_env("BDLRU_FS_NT");
  int NT = D <= 128 ? 96 : 64;
  if (force_nt == 64) NT = 64;
  int NSTG = 2;
  if (NT == 64 && UB * (D / 2) + 3 * UB * 64 <= 512) NSTG = 3;
  pl->NSTG = NSTG;
  const long tiles = (n_rows + NT - 1) / NT;
  const size_t stage = (size_t)n_slab * NT * 128;
  int stages = (int)(((size_t)kMaxSmem - 1024 - 256) / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return false;
  pl->UB = UB;
  pl->NT = NT;
  pl->stages = stages;
  pl->n_ug = (int)((n_users + (long)kTile * UB - 1) / ((long)kTile * UB));
  // ONE wave of long item streams: the insert work per user grows only with the log of the stream length, so
  // fewer, longer streams keep the epilogue under the MMA time.  splits = floor(SMs / user groups), at least 4 tiles
  // per CTA; when there are more user groups than SMs every group gets one CTA (several waves).
  long max_s = tiles / 4 > 0 ? tiles / 4 : 1;
  long want = (long)sm_count() / pl->n_ug;
  if (want < 1) want = 1;
  pl->splits = (int)(want < max_s ? want : max_s);
  if (mode == MODE_CE || mode == MODE_MAX) {
    // The softmax statistics have no per-stream cost that grows with the split count (a (max, sum) pair per split), so
    // the CE grid need not be one wave: with 32 user groups one wave is 128 CTAs on 148 SMs (ncu: SMs active 85 % of
    // the time).  Take more, shorter streams (>= 32 tiles each) when that fills whole waves better.
    auto eff = [&](long sp) {
      const long ctas = (long)pl->n_ug * sp, sm = sm_count();
      return (double)ctas / (double)(((ctas + sm - 1) / sm) * sm);
    };
    long best = pl->splits;
    double best_eff = eff(best);
    const long cap = tiles / 32 < 256 ? tiles / 32 : 256;
    for (long sp = best + 1; sp <= cap; ++sp)
      if (eff(sp) > best_eff + 0.02) {
        best = sp;
        best_eff = eff(sp);
      }
    pl->splits = (int)best;
  }
  pl->threads = 96 + 128 * UB;
  pl->smem = 1024 + (size_t)stages * stage + 256;
  pl->tiles_total = tiles;
  return true;
}

// ----------------------------------------------------------------------------- per-thread sorted list in registers
// Inserts (v, id) after every entry with score >= v (branch-free, fully unrolled).  Caller has checked v > ls[K-1].
template <int K>
__device__ __forceinline__ void topk_insert(float (&ls)[K], int (&li)[K], float v, int id) {
#pragma unroll
  for (int j = K - 1; j >= 1; --j) {
    const bool prev = ls[j - 1] < v;  // entry j-1 moves down to j
    const bool cur = ls[j] < v;
    ls[j] = prev ? ls[j - 1] : (cur ? v : ls[j]);
    li[j] = prev ? li[j - 1] : (cur ? id : li[j]);
  }
  if (ls[0] < v) {
    ls[0] = v;
    li[0] = id;
  }
}

// Order-preserving float <-> unsigned (0 is below every float): lets the splits of a user group share their K-th best
// score through atomicMax.
__device__ __forceinline__ unsigned thr_encode(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float thr_decode(unsigned k) {
  if (k == 0u) return -INFINITY;
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// 3-input max (FMNMX3 on sm_100): 32 values in 16 instructions instead of 31 — the epilogue's threshold filter is
// bound by the issue rate of these min/max instructions (measured ~185 cycles per 32-column chunk before).
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max32(const float (&v)[32]) {
  float m[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) m[i] = max3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  const float a = max3(m[0], m[1], m[2]), b = max3(m[3], m[4], m[5]), c = max3(m[6], m[7], m[8]);
  return max3(max3(a, b, c), max3(m[9], v[30], v[31]), -INFINITY);
}

// First index i with v[i] == m, where m = max32(v): winners of the ten triples are found independently, then a 10-long
// select chain (instead of 31 dependent selects over the elements — the insert rounds are latency-bound).
__device__ __forceinline__ int first_argmax32(const float (&v)[32], float m) {
  int best = (v[30] == m) ? 30 : 31;
#pragma unroll
  for (int j = 9; j >= 0; --j) {
    const float t = max3(v[3 * j], v[3 * j + 1], v[3 * j + 2]);
    const int li = (v[3 * j] == t) ? 3 * j : ((v[3 * j + 1] == t) ? 3 * j + 1 : 3 * j + 2);
    best = (t == m) ? li : best;
  }
  return best;
}

// ES = epilogue warp SETS (CE only).  The softmax statistics need one MUFU.EX2 per logit (16/clk/SM: 1 536 cycles per
// 256 x 96 tile against 768 of MMA), and with one set the two warps of each scheduler run in lock step behind the same
// accumulator hand-off: they fight over the SFU while exponentiating and leave it idle while waiting / loading / taking
// maxima (measured ~2 600 cycles per tile).  With two sets, set s owns tiles it = s (mod 2) and walks them one 32-column
// chunk at a time (register budget at 19 warps), so some warp of every scheduler is always in its exp phase; each set
// keeps its own (max, sum) pair and the merge kernel combines splits x sets partials.
// Measured at 8192 x 1 M x 128: 3.38 -> 2.85 ms on the same box (2.77 with 64-item tiles and 3 accumulator stages, not
// adopted: the tiling plan is shared with the top-k).  Keeping the next chunk's tcgen05.ld in flight while a chunk is
// processed (double-buffered registers) gave 3.38 ms at 96-item tiles, i.e. nothing: dropped.
template <int UB, int MODE, int K, int NT, int NSTG, int ES>
__global__ void __launch_bounds__(96 + 128 * UB * ES, 1)
fullsort_kernel(const __grid_constant__ CUtensorMap tmE, const FsParams p) {
  static_assert(ES == 1 || MODE == MODE_CE, "a second epilogue set exists for the CE statistics only");
  static_assert(MODE != MODE_MAX || K == 1, "MODE_MAX keeps no list");
  constexpr int NSETS = ES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slab = p.D >> 6;
  constexpr uint32_t kSlabB = NT * 128;    // bytes of one [NT items x 64 channels] swizzled slab
  constexpr int NCH = NT / 32;             // 32-column chunks per accumulator tile
  uint8_t* sE = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sE + (size_t)p.stages * n_slab * kSlabB);
  uint64_t* q_full = bars;
  uint64_t* e_full = bars + 1;
  uint64_t* e_empty = e_full + kMaxStages;
  uint64_t* acc_full = e_empty + kMaxStages;
  uint64_t* acc_empty = acc_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 4);

  // work assignment: consecutive CTAs share an item range and differ in user group -> E tiles hit in L2
  const int ug = blockIdx.x % p.n_ug;
  const int split = blockIdx.x / p.n_ug;
  const long t_begin = p.tiles_total * split / p.splits;
  const long t_end = p.tiles_total * (split + 1) / p.splits;
  const int n_iter = (int)(t_end - t_begin);
  const long user0 = (long)ug * kTile * UB;
  // TMEM columns: [0, UB*D/2) the user block(s) Q as packed bf16 pairs (A operand), then NSTG x UB accumulators of NT
  const uint32_t q_cols = (uint32_t)(p.D >> 1);
  const uint32_t acc_col0 = (uint32_t)UB * q_cols;

  if (warp == 2 && lane == 0) {
    tc::prefetch_tensormap(&tmE);
    tc::mbar_init(q_full, 4 * UB);
    for (int s = 0; s < p.stages; ++s) {
      tc::mbar_init(&e_full[s], 1);
      tc::mbar_init(&e_empty[s], 1);
    }
    for (int b = 0; b < NSTG; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], 4 * UB);
    }
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

  if (warp == 2) {
    // ===================================================================== TMA producer (warp stays converged)
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      tc::mbar_wait(&e_empty[s], ph ^ 1u);
      if (tc::elect_one()) {
        if (FS_DBG(p) & 4) {
          tc::mbar_arrive(&e_full[s]);
        } else {
        tc::mbar_arrive_expect_tx(&e_full[s], (uint32_t)n_slab * kSlabB);
        for (int sl = 0; sl < n_slab; ++sl)
          tc::tma_load_2d(sE + (size_t)(s * n_slab + sl) * kSlabB, &tmE, &e_full[s], sl * 64,
                          (int)((t_begin + it) * p.tile_stride * NT));
        }
      }
      __syncwarp();
    }
  } else if (warp < 2) {
    // ===================================================================== MMA issuers (warps stay converged)
    // TWO issuing warps, tile it handled by warp it % 2.  A tcgen05.mma blocks its issuing thread for about the MMA's
    // execution time and each mbarrier wait costs ~300 cycles even when already satisfied (measured with clock64:
    // 409 + 328 cycles of waits against 870 of issue per tile), so a single issuer leaves the tensor pipe idle half the
    // time; with two, one warp's waits overlap the other's MMAs.
    constexpr uint32_t idesc = tc::idesc_bf16_f32(kTile, NT, 0, 0);
    tc::mbar_wait(q_full, 0);
    tc::fence_after_sync();
    long long t_acc = 0, t_ring = 0, t_issue = 0;
    for (int it = warp; it < n_iter; it += 2) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      const int b = it % NSTG;
      const uint32_t bph = (uint32_t)(it / NSTG) & 1u;
      const long long c0 = FS_CLOCK();
      tc::mbar_wait(&acc_empty[b], bph ^ 1u);
      const long long c1 = FS_CLOCK();
      tc::mbar_wait(&e_full[s], ph);
      const long long c2 = FS_CLOCK();
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t e0 = tc::smem_u32(sE + (size_t)s * n_slab * kSlabB);
        // descriptor of (slab 0, k step 0); a K step of 16 channels is +32 bytes (+2 in the >>4 address field) inside
        // the swizzle atom, a slab is +kSlabB
        const uint64_t bd0 = tc::smem_desc_sw128(e0, 16, 1024);
#pragma unroll
        for (int ub = 0; ub < UB; ++ub) {
          if (FS_DBG(p) & 2) break;
          const uint32_t d_tmem = tmem_base + acc_col0 + (uint32_t)((b * UB + ub) * NT);
          const uint32_t a_tmem = tmem_base + (uint32_t)ub * q_cols;
          for (int sl = 0; sl < n_slab; ++sl) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              // A: 16 channels = 8 packed TMEM columns per K step
              tc::umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(sl * 4 + k4) * 8,
                               bd0 + (uint64_t)((uint32_t)sl * (kSlabB >> 4) + (uint32_t)k4 * 2), idesc,
                               (uint32_t)((sl | k4) != 0));
            }
          }
        }
        tc::umma_commit(&e_empty[s]);   // smem stage reusable once these MMAs have read it
        tc::umma_commit(&acc_full[b]);  // accumulators complete
      }
      __syncwarp();
      const long long c3 = FS_CLOCK();
      t_acc += c1 - c0; t_ring += c2 - c1; t_issue += c3 - c2;
    }
    if ((FS_DBG(p) & 8) && lane == 0 && blockIdx.x == 0 && n_iter > 1)
      printf("[fs mma warp %d] tiles %d: wait acc_empty %lld, wait e_full %lld, issue %lld cycles/tile\n", warp, n_iter,
             t_acc / (n_iter / 2), t_ring / (n_iter / 2), t_issue / (n_iter / 2));
  } else {
    // ===================================================================== epilogue: thread == user
    const int e = warp - 3;
    const int ub = (e >> 2) % UB;
    const int eset = e / (4 * UB);          // epilogue set (CE with ES = 2: tiles it = eset mod 2), else 0
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // user row inside the block
    const long user = user0 + ub * kTile + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    if (ES == 1 || eset == 0) {
      // This thread's user row of Q -> TMEM lane `row`, columns ub*D/2 .. : bf16 pairs, channel 2j in the low half.
      // The A operand of every MMA of this CTA then comes from TMEM, which halves the shared-memory operand traffic
      // (with both operands in shared memory a 128x128x16 MMA needs 128 B/clk, all the SM has: measured half rate).
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.Q) + user * p.D);
      for (int j = 0; j < (p.D >> 4); ++j) {
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (user < p.n_users) {
          lo = src[2 * j];
          hi = src[2 * j + 1];
        }
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tc::tmem_st_32x32_x8(lane_addr + (uint32_t)ub * q_cols + (uint32_t)j * 8, w);
      }
      tc::tmem_st_wait();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(q_full);
    }
    float ls[K];                            // top-k: K best scores, descending
    int li[K];                              //        and their global ids
#pragma unroll
    for (int j = 0; j < K; ++j) {
      ls[j] = -INFINITY;
      li[j] = -1;
    }
    // top-k: the CTAs that stream different item ranges for the SAME users publish their K-th best score; an item below
    // the best published K-th cannot be in the final top-K, so every split filters with it (>=: ties are still admitted
    // and resolved by id in the merge).  This makes S short streams insert like ONE long stream (K ln(N/K) in total
    // instead of per split).  `sthr` lags by one tile so the L2 read is never waited for.
    float sthr = -INFINITY;
    const bool share = MODE == MODE_TOPK && p.shared_thr != nullptr && user < p.n_users;
    // loaded one tile ago, decoded (= first use) only now; the first value is the SEED a sampled pre-pass left there
    // (bdlru_fullsort_topk), fetched here so that its latency hides behind the wait for the first accumulator
    unsigned sthr_raw = share ? __ldcg(p.shared_thr + user) : 0u;
    float run_m = -INFINITY, run_s = 0.f;   // CE: online max / sum of exp
    long pos_local = -1;
    if (MODE == MODE_CE && user < p.n_users) {
      pos_local = p.pos[user] - p.id_offset;
      if (pos_local >= p.n_rows) pos_local = -1;  // positive lives in another shard (its tail rows are masked here)
    }
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr bool CLK = ES == 1;           // clock64 phase counters (BDLRU_FS_DEBUG & 8) only where registers allow
    long long te_wait = 0, te_ld = 0;
    const long long te_begin = CLK ? FS_CLOCK() : 0;
    for (int it = (ES == 1 ? 0 : eset); it < n_iter; it += NSETS) {
      const int b = it % NSTG;
      const uint32_t bph = (uint32_t)(it / NSTG) & 1u;
      const long base = (t_begin + it) * p.tile_stride * NT;  // local row index of the tile's first item
      const bool special = (base + NT > p.n_rows) || (p.mask_local >= base && p.mask_local < base + NT);
      const long long e0c = CLK ? FS_CLOCK() : 0;
      tc::mbar_wait(&acc_full[b], bph);
      const long long e1c = CLK ? FS_CLOCK() : 0;
      te_wait += e1c - e0c;
      tc::fence_after_sync();
      if (share) {
        sthr = fmaxf(sthr, thr_decode(sthr_raw));
        sthr_raw = __ldcg(p.shared_thr + user);
      }
      const uint32_t taddr = lane_addr + acc_col0 + (uint32_t)((b * UB + ub) * NT);
      // The accumulator tile is pulled into registers CH chunks (CH*32 columns) at a time; once the last group has
      // landed the TMEM buffer is handed back to the MMA warp BEFORE the scores are processed, so the tensor pipe
      // only ever waits for the loads, not for the top-k / softmax arithmetic.
      constexpr int CH = ((MODE == MODE_TOPK && K > 16) || ES > 1) ? 1 : NCH;
      if (FS_DBG(p) & 1) {
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[b]);
        continue;
      }
#pragma unroll 1
      for (int g = 0; g < NCH / CH; ++g) {
        uint32_t raw[CH][32];
#pragma unroll
        for (int c = 0; c < CH; ++c) tc::tmem_ld_32x32(taddr + (g * CH + c) * 32, raw[c]);
        tc::tmem_ld_wait();
        if (g == NCH / CH - 1) {
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_empty[b]);
          if (CLK) te_ld += FS_CLOCK() - e1c;
        }
        // all CH chunk maxima first (independent trees: ILP), one vote for the common case "nothing to insert"
        float v[CH][32];
        float m[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c][i] = __uint_as_float(raw[c][i]);
          if (special) {
            const long cb = base + (g * CH + c) * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (cb + i >= p.n_rows || cb + i == p.mask_local) v[c][i] = -INFINITY;
          }
          m[c] = max32(v[c]);
        }
        float mt = m[0];
#pragma unroll
        for (int c = 1; c < CH; ++c) mt = fmaxf(mt, m[c]);
        if (MODE == MODE_TOPK) {
          if (__any_sync(0xffffffffu, mt > ls[K - 1] && mt >= sthr)) {
            const float kth_before = ls[K - 1];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const int id0 = (int)(p.id_offset + base + (g * CH + c) * 32);
              // insert rounds: lanes with a candidate act together
              while (__any_sync(0xffffffffu, m[c] > ls[K - 1] && m[c] >= sthr)) {
                if (m[c] > ls[K - 1] && m[c] >= sthr) {
                  const int idx = first_argmax32(v[c], m[c]);  // first index among equal scores
                  topk_insert<K>(ls, li, m[c], id0 + idx);
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[c][i] = (i == idx) ? -INFINITY : v[c][i];
                  m[c] = max32(v[c]);
                }
              }
            }
            if (share && ls[K - 1] > kth_before) atomicMax(p.shared_thr + user, thr_encode(ls[K - 1]));
          }
        } else if (MODE == MODE_MAX) {
          run_m = fmaxf(run_m, mt);
        } else {
          if (mt > run_m) {  // rescale the running sum to the new maximum (mt is finite here)
            run_s *= ex2_ftz((run_m - mt) * kLog2e);
            run_m = mt;
          }
          if (run_m > -INFINITY) {
            const float mb = run_m * kLog2e;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) acc[u] += ex2_mixed<kPolyMaskFwd, 4>(fmaf(v[c][i + u], kLog2e, -mb), i + u);
              }
            run_s += (acc[0] + acc[1]) + (acc[2] + acc[3]);
          }
          const long gb = base + g * CH * 32;
          if (pos_local >= gb && pos_local < gb + CH * 32) {
            const int sel = (int)(pos_local - gb);
            float pv = 0.f;
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i == sel) pv = v[c][i];
            p.pos_logit[user] = pv;
          }
        }
      }
    }
    if (CLK && (FS_DBG(p) & 8) && lane == 0 && blockIdx.x == 0 && warp == 3 && n_iter > 0)
      printf("[fs epilogue warp] per tile: total %lld, wait acc_full %lld, ld+release %lld cycles\n",
             (FS_CLOCK() - te_begin) / n_iter, te_wait / n_iter, te_ld / n_iter);
    if (user < p.n_users) {
      if (MODE == MODE_TOPK) {
        float* os = p.part_scores + ((size_t)user * p.splits + split) * p.k;
        int* oi = p.part_ids + ((size_t)user * p.splits + split) * p.k;
#pragma unroll
        for (int j = 0; j < K; ++j)
          if (j < p.k) {
            os[j] = ls[j];
            oi[j] = li[j];
          }
      } else {
        p.part_max[(size_t)user * p.ce_parts + split * NSETS + eset] = run_m;
        if (MODE == MODE_CE) p.part_sum[(size_t)user * p.ce_parts + split * NSETS + eset] = run_s;
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------- merges
// One warp per user: k rounds of "best remaining candidate" by (score desc, id asc) over n_cand candidates staged in
// shared memory.  Used for the S per-CTA lists of one GPU and for the G per-shard lists of the NCCL merge.
// Candidate (list l, user u, slot j) lives at l * list_stride + u * user_stride + j (elements): user-major lists
// ([n_users][n_lists][k]: list_stride = k, user_stride = n_lists * k) and the rank-major layout an NCCL all-gather
// leaves behind ([n_lists][...][n_users][k]) are both read in place, no permute copy.
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* __restrict__ cs, const int* __restrict__ ci,
                                                         long n_users, int n_cand, int k, long list_stride,
                                                         long user_stride, float* __restrict__ os,
                                                         int* __restrict__ oi, unsigned* __restrict__ seed_out) {
  extern __shared__ uint8_t msm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long user = (long)blockIdx.x * 4 + warp;
  if (user >= n_users) return;
  float* s = reinterpret_cast<float*>(msm) + (size_t)warp * n_cand;
  int* id = reinterpret_cast<int*>(msm + (size_t)4 * n_cand * sizeof(float)) + (size_t)warp * n_cand;
  for (int j = lane; j < n_cand; j += 32) {
    const long src = (long)(j / k) * list_stride + user * user_stride + (j % k);
    s[j] = cs[src];
    id[j] = ci[src];
  }
  __syncwarp();
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY;
    int bi = 0x7fffffff, bj = -1;
    for (int j = lane; j < n_cand; j += 32) {
      const float x = s[j];
      const int y = id[j];
      if (y >= 0 && (x > bs || (x == bs && y < bi))) {
        bs = x; bi = y; bj = j;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float xs = __shfl_xor_sync(0xffffffffu, bs, o);
      const int xi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int xj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (xj >= 0 && (bj < 0 || xs > bs || (xs == bs && xi < bi))) {
        bs = xs; bi = xi; bj = xj;
      }
    }
    if (lane == 0) {
      os[user * k + r] = bj >= 0 ? bs : -INFINITY;
      oi[user * k + r] = bj >= 0 ? bi : -1;
      if (bj >= 0) id[bj] = -1;  // taken
      // pre-pass: the K-th best of the SAMPLE is a lower bound of the K-th best of the whole table -> threshold seed
      if (seed_out && r == k - 1 && bj >= 0) atomicMax(seed_out + user, thr_encode(bs));
    }
    __syncwarp();
  }
}

// CE: combine the per-split (max, sumexp) pairs of each user.
__global__ void ce_merge_kernel(const float* __restrict__ pm, const float* __restrict__ ps, long n_users, int splits,
                                float* __restrict__ row_max, float* __restrict__ row_sum) {
  const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  float m = -INFINITY;
  for (int j = 0; j < splits; ++j) m = fmaxf(m, pm[u * splits + j]);
  float s = 0.f;
  for (int j = 0; j < splits; ++j) {
    const float mj = pm[u * splits + j];
    if (mj > -INFINITY) s += ps[u * splits + j] * expf(mj - m);
  }
  row_max[u] = m;
  row_sum[u] = s;
}

__global__ void max_merge_kernel(const float* __restrict__ pm, long n_users, int splits, float* __restrict__ row_max) {
  const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  float m = -INFINITY;
  for (int j = 0; j < splits; ++j) m = fmaxf(m, pm[u * splits + j]);
  row_max[u] = m;
}

// ----------------------------------------------------------------------------- host side
static int fs_debug() {
  static const int v = tuning_env("BDLRU_FS_DEBUG");
  return v;
}

static int fs_check(const void* Q, const void* E, long n_users, long n_rows, int D) {
  BDLRU_REQUIRE(Q && E, "fullsort: null Q/E");
  BDLRU_REQUIRE(n_users >= 1 && n_rows >= 1, "fullsort: bad sizes n_users=%ld n_rows=%ld", n_users, n_rows);
  BDLRU_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256, "fullsort: D=%d must be a multiple of 64 in [64, 256]", D);
  BDLRU_REQUIRE(aligned(Q, 16) && aligned(E, 16), "fullsort: Q/E must be 16-byte aligned");
  BDLRU_REQUIRE(n_rows < (1L << 31) && n_users < (1L << 31), "fullsort: sizes must fit int32");
  return BDLRU_OK;
}

// Softmax warp sets of the CE forward (see fullsort_kernel).  BDLRU_CE_SETS=1 selects the single-set kernel (A/B runs).
static int ce_sets() {
  static const int v = tuning_env("BDLRU_CE_SETS") == 1 ? 1 : 2;
  return v;
}

template <int UB, int MODE, int K, int NT, int NSTG, int ES>
static int fs_launch_one(const FsPlan& pl, const CUtensorMap& me, const FsParams& p, cudaStream_t st) {
  BDLRU_CUDA(cudaFuncSetAttribute(fullsort_kernel<UB, MODE, K, NT, NSTG, ES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)pl.smem));
  fullsort_kernel<UB, MODE, K, NT, NSTG, ES><<<pl.n_ug * pl.splits, 96 + 128 * UB * ES, pl.smem, st>>>(me, p);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

template <int MODE, int K, int ES>
static int fs_launch_k(const FsPlan& pl, const CUtensorMap& me, const FsParams& p, cudaStream_t st) {
  if (pl.NT == 96)
    return pl.UB == 2 ? fs_launch_one<2, MODE, K, 96, 2, ES>(pl, me, p, st) : fs_launch_one<1, MODE, K, 96, 2, ES>(pl, me, p, st);
  if (pl.NSTG == 3)
    return pl.UB == 2 ? fs_launch_one<2, MODE, K, 64, 3, ES>(pl, me, p, st) : fs_launch_one<1, MODE, K, 64, 3, ES>(pl, me, p, st);
  return pl.UB == 2 ? fs_launch_one<2, MODE, K, 64, 2, ES>(pl, me, p, st) : fs_launch_one<1, MODE, K, 64, 2, ES>(pl, me, p, st);
}

template <int MODE>
static int fs_launch(const FsPlan& pl, const CUtensorMap& me, const FsParams& p, cudaStream_t st) {
  if (MODE == MODE_CE)
    return ce_sets() == 2 ? fs_launch_k<MODE_CE, 1, 2>(pl, me, p, st) : fs_launch_k<MODE_CE, 1, 1>(pl, me, p, st);
  if (MODE == MODE_MAX) return fs_launch_k<MODE_MAX, 1, 1>(pl, me, p, st);
  if (p.k <= 10) return fs_launch_k<MODE_TOPK, 10, 1>(pl, me, p, st);
  if (p.k <= 16) return fs_launch_k<MODE_TOPK, 16, 1>(pl, me, p, st);
  if (p.k <= 20) return fs_launch_k<MODE_TOPK, 20, 1>(pl, me, p, st);
  return fs_launch_k<MODE_TOPK, 32, 1>(pl, me, p, st);
}

}  // namespace bdlru

using namespace bdlru;

extern "C" BDLRU_API int bdlru_fullsort_available(void) { return 1; }

extern "C" BDLRU_API size_t bdlru_fullsort_topk_workspace_bytes(int64_t n_users, int64_t n_rows, int D, int k) {
  FsPlan pl;
  if (D % 64 != 0 || D < 64 || D > 256 || k < 1 || k > 32 || !fs_plan(n_users, n_rows, D, k, MODE_TOPK, &pl)) return 0;
  return (size_t)n_users * pl.splits * k * 8 + (size_t)n_users * 4;
}

extern "C" BDLRU_API int bdlru_fullsort_topk(const void* Q, const void* E, int64_t n_users, int64_t n_rows, int D, int k,
                                             int64_t id_offset, int64_t mask_id, float* out_scores, int32_t* out_ids,
                                             void* workspace, size_t workspace_bytes, void* stream) {
  int rc = fs_check(Q, E, n_users, n_rows, D);
  if (rc) return rc;
  BDLRU_REQUIRE(k >= 1 && k <= 32, "fullsort_topk: k=%d not in [1, 32]", k);
  BDLRU_REQUIRE(out_scores && out_ids, "fullsort_topk: null outputs");
  BDLRU_REQUIRE(id_offset >= 0 && id_offset + n_rows < (1L << 31), "fullsort_topk: item ids must fit int32");
  FsPlan pl;
  BDLRU_REQUIRE(fs_plan(n_users, n_rows, D, k, MODE_TOPK, &pl), "fullsort_topk: no tiling fits shared memory (D=%d k=%d)", D, k);
  const size_t need = (size_t)n_users * pl.splits * k * 8 + (size_t)n_users * 4;
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "fullsort_topk: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap me;
  if ((rc = make_rows_map(&me, E, n_rows, D, pl.NT))) return rc;
  FsParams p = {};
  p.Q = Q;
  p.D = D; p.k = k; p.stages = pl.stages; p.splits = pl.splits; p.n_ug = pl.n_ug;
  p.dbg = fs_debug();
  p.n_users = n_users; p.n_rows = n_rows; p.id_offset = id_offset;
  p.mask_local = (mask_id >= id_offset && mask_id < id_offset + n_rows) ? mask_id - id_offset : -1;
  p.tiles_total = pl.tiles_total;
  p.tile_stride = 1;
  p.part_scores = reinterpret_cast<float*>(workspace);
  p.part_ids = reinterpret_cast<int*>(p.part_scores + (size_t)n_users * pl.splits * k);
  p.shared_thr = nullptr;
  const int n_cand = pl.splits * k;
  const size_t msmem = (size_t)4 * n_cand * 8;
  BDLRU_REQUIRE(msmem <= 200 * 1024, "fullsort_topk: merge of %d candidates per user does not fit", n_cand);
  BDLRU_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  if (pl.splits > 1 && !(fs_debug() & 16)) {  // BDLRU_FS_DEBUG & 16: disable threshold sharing (tuning / A-B runs)
    p.shared_thr = reinterpret_cast<unsigned*>(p.part_ids + (size_t)n_users * pl.splits * k);
    BDLRU_CUDA(cudaMemsetAsync(p.shared_thr, 0, (size_t)n_users * 4, st));
    // Threshold SEED.  A stream's insert work is ~K ln(n/K) and almost all of it falls on its first tiles, which is what
    // keeps 1 M-item shards at 0.74 of the tensor rate against 0.95+ at 10 M.  A pre-pass runs the SAME kernel over a
    // strided sample of the tiles (>= 256 tiles, ~1/32 of the table) and its merged K-th best score per user — a lower
    // bound of the true K-th best, whatever the sample — is left in shared_thr, so the real streams start filtering at
    // the ~(32 K)-th best score instead of -inf.  The result is unchanged (ties at the threshold are admitted).
    const long n_samp = pl.tiles_total / 32 > 256 ? pl.tiles_total / 32 : 256;
    const long stride = pl.tiles_total / n_samp;
    if (stride >= 4 && !(fs_debug() & 32)) {  // BDLRU_FS_DEBUG & 32: no seed (A-B runs)
      FsParams ps = p;
      FsPlan pp = pl;
      ps.tile_stride = (int)stride;
      ps.tiles_total = (pl.tiles_total + stride - 1) / stride;
      const long max_s = ps.tiles_total / 4 > 0 ? ps.tiles_total / 4 : 1;
      if (pp.splits > max_s) pp.splits = (int)max_s;
      ps.splits = pp.splits;
      if ((rc = fs_launch<MODE_TOPK>(pp, me, ps, st))) return rc;
      const int nc = pp.splits * k;
      // the sample's merged lists land in the output buffers (overwritten by the real merge below)
      topk_merge_kernel<<<(unsigned)((n_users + 3) / 4), 128, (size_t)4 * nc * 8, st>>>(
          ps.part_scores, ps.part_ids, n_users, nc, k, k, (long)nc, out_scores, out_ids, p.shared_thr);
      BDLRU_LAUNCHED();
    }
  }
  if ((rc = fs_launch<MODE_TOPK>(pl, me, p, st))) return rc;
  topk_merge_kernel<<<(unsigned)((n_users + 3) / 4), 128, msmem, st>>>(p.part_scores, p.part_ids, n_users, n_cand, k, k,
                                                                      (long)n_cand, out_scores, out_ids, nullptr);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

static int merge_impl(const float* cand_scores, const int32_t* cand_ids, int64_t n_users, int n_lists, int k,
                      int64_t list_stride, int64_t user_stride, float* out_scores, int32_t* out_ids, void* stream) {
  BDLRU_REQUIRE(cand_scores && cand_ids && out_scores && out_ids, "topk_merge: null pointer");
  BDLRU_REQUIRE(n_users >= 1 && n_lists >= 1 && k >= 1 && k <= 32, "topk_merge: bad sizes");
  BDLRU_REQUIRE(list_stride >= 1 && user_stride >= k, "topk_merge: bad strides");
  const int n_cand = n_lists * k;
  const size_t msmem = (size_t)4 * n_cand * 8;
  BDLRU_REQUIRE(msmem <= 200 * 1024, "topk_merge: %d candidates per user do not fit", n_cand);
  BDLRU_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
  topk_merge_kernel<<<(unsigned)((n_users + 3) / 4), 128, msmem, reinterpret_cast<cudaStream_t>(stream)>>>(
      cand_scores, cand_ids, n_users, n_cand, k, list_stride, user_stride, out_scores, out_ids, nullptr);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

extern "C" BDLRU_API int bdlru_topk_merge(const float* cand_scores, const int32_t* cand_ids, int64_t n_users,
                                          int n_lists, int k, float* out_scores, int32_t* out_ids, void* stream) {
  return merge_impl(cand_scores, cand_ids, n_users, n_lists, k, k, (int64_t)n_lists * k, out_scores, out_ids, stream);
}

extern "C" BDLRU_API int bdlru_topk_merge_strided(const float* cand_scores, const int32_t* cand_ids, int64_t n_users,
                                                  int n_lists, int k, int64_t list_stride, int64_t user_stride,
                                                  float* out_scores, int32_t* out_ids, void* stream) {
  return merge_impl(cand_scores, cand_ids, n_users, n_lists, k, list_stride, user_stride, out_scores, out_ids, stream);
}

namespace bdlru { size_t ce_bwd_workspace_bytes(long n_users, long n_rows, int D); }

extern "C" BDLRU_API size_t bdlru_fullsort_ce_workspace_bytes(int64_t n_users, int64_t n_rows, int D) {
  FsPlan pl;
  if (D % 64 != 0 || D < 64 || D > 256 || !fs_plan(n_users, n_rows, D, 1, MODE_CE, &pl)) return 0;
  const size_t fwd = (size_t)n_users * pl.splits * ce_sets() * 8;
  const size_t bwd = ce_bwd_workspace_bytes(n_users, n_rows, D);
  return fwd > bwd ? fwd : bwd;
}

extern "C" BDLRU_API int bdlru_fullsort_ce_fwd(const void* Q, const void* E, const int64_t* pos, int64_t n_users,
                                               int64_t n_rows, int D, int64_t id_offset, float* row_max,
                                               float* row_sumexp, float* pos_logit, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  int rc = fs_check(Q, E, n_users, n_rows, D);
  if (rc) return rc;
  BDLRU_REQUIRE(pos && row_max && row_sumexp && pos_logit, "fullsort_ce_fwd: null pointer");
  FsPlan pl;
  BDLRU_REQUIRE(fs_plan(n_users, n_rows, D, 1, MODE_CE, &pl), "fullsort_ce_fwd: no tiling fits shared memory (D=%d)", D);
  const int parts = pl.splits * ce_sets();
  const size_t need = (size_t)n_users * parts * 8;
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "fullsort_ce_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap me;
  if ((rc = make_rows_map(&me, E, n_rows, D, pl.NT))) return rc;
  FsParams p = {};
  p.Q = Q;
  p.dbg = fs_debug();
  p.D = D; p.k = 1; p.stages = pl.stages; p.splits = pl.splits; p.n_ug = pl.n_ug;
  p.n_users = n_users; p.n_rows = n_rows; p.id_offset = id_offset; p.mask_local = -1;
  p.tiles_total = pl.tiles_total;
  p.tile_stride = 1;
  p.ce_parts = parts;
  p.part_max = reinterpret_cast<float*>(workspace);
  p.part_sum = p.part_max + (size_t)n_users * parts;
  p.pos = pos;
  p.pos_logit = pos_logit;
  if ((rc = fs_launch<MODE_CE>(pl, me, p, st))) return rc;
  ce_merge_kernel<<<(unsigned)((n_users + 255) / 256), 256, 0, st>>>(p.part_max, p.part_sum, n_users, parts,
                                                                     row_max, row_sumexp);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}

// Row maxima of the logits Q E^T over this shard (no exponentials: tensor-bound), or over every tile_stride-th 96-item
// tile only — a SAMPLED maximum, good enough as the softmax reference of bdlru_fullsort_ce_fwd_dq (any reference within
// ~80 of the true maximum gives the same result; see the header).
extern "C" BDLRU_API size_t bdlru_fullsort_rowmax_workspace_bytes(int64_t n_users, int64_t n_rows, int D) {
  FsPlan pl;
  if (D % 64 != 0 || D < 64 || D > 256 || !fs_plan(n_users, n_rows, D, 1, MODE_MAX, &pl)) return 0;
  return (size_t)n_users * pl.splits * 4;
}

extern "C" BDLRU_API int bdlru_fullsort_rowmax(const void* Q, const void* E, int64_t n_users, int64_t n_rows, int D,
                                               int tile_stride, float* row_max, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  int rc = fs_check(Q, E, n_users, n_rows, D);
  if (rc) return rc;
  BDLRU_REQUIRE(row_max, "fullsort_rowmax: null output");
  BDLRU_REQUIRE(tile_stride >= 1, "fullsort_rowmax: tile_stride=%d must be >= 1", tile_stride);
  FsPlan pl;
  BDLRU_REQUIRE(fs_plan(n_users, n_rows, D, 1, MODE_MAX, &pl), "fullsort_rowmax: no tiling fits shared memory (D=%d)", D);
  const long tiles_eff = (pl.tiles_total + tile_stride - 1) / tile_stride;
  // re-plan the splits for the tiles actually walked (same workspace bound: splits never grow)
  FsPlan pe = pl;
  if (tile_stride > 1) {
    long max_s = tiles_eff / 4 > 0 ? tiles_eff / 4 : 1;
    if (pe.splits > max_s) pe.splits = (int)max_s;
  }
  const size_t need = (size_t)n_users * pe.splits * 4;
  BDLRU_REQUIRE(workspace && workspace_bytes >= need, "fullsort_rowmax: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap me;
  if ((rc = make_rows_map(&me, E, n_rows, D, pe.NT))) return rc;
  FsParams p = {};
  p.Q = Q;
  p.dbg = fs_debug();
  p.D = D; p.k = 1; p.stages = pe.stages; p.splits = pe.splits; p.n_ug = pe.n_ug;
  p.n_users = n_users; p.n_rows = n_rows; p.id_offset = 0; p.mask_local = -1;
  p.tiles_total = tiles_eff;
  p.tile_stride = tile_stride;
  p.ce_parts = pe.splits;
  p.part_max = reinterpret_cast<float*>(workspace);
  p.part_sum = nullptr;
  if ((rc = fs_launch<MODE_MAX>(pe, me, p, st))) return rc;
  max_merge_kernel<<<(unsigned)((n_users + 255) / 256), 256, 0, st>>>(p.part_max, n_users, pe.splits, row_max);
  BDLRU_LAUNCHED();
  return BDLRU_OK;
}
