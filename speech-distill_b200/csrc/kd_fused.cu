// K1: fused LM head + KD loss on the Blackwell tensor pipeline (tcgen05 / TMEM / TMA).
//
// Replaces `logits = lm_head(hidden)` (transformers Qwen3ForCausalLM, called at train.py:54)
// followed by DistillationLoss.forward (distillation_loss.py:14-128) and their autograd, without
// ever materialising the [rows, V] logits.
//
// One persistent, warp-specialised kernel template serves every GEMM on the path (CTA pairs by default:
// tcgen05 cta_group::2, 256 x 256 tiles, see kd_umma_kernel):
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 3..8-stage smem ring, SWIZZLE_128B)
//   warp 1      : MMA issuer    (one thread, tcgen05.mma kind::f16, 256 x 256 x 16 per pair, fp32 in TMEM)
//   warp 2      : TMEM allocator (2 accumulator buffers x 256 columns = all 512 columns) + work-unit scheduler
//   warp 3      : TMA producer of the teacher-tile ring
//   warps 4..19 : epilogue      (tcgen05.ld 32x32b: one thread per accumulator row, 4 column groups)
// and the epilogue is a policy:
//   FwdEpi  : online soft-max statistics of the student tile + streamed teacher tile (forward)
//   GradEpi : recomputed tile -> gradient tile G (bf16) into a V-independent scratch (backward)
//   StoreEpi: dW chunk = G^T h (final rows, bf16) and dH += G W_chunk (fp32 accumulate -> bf16)
//
// Algorithmic FLOPs: 2 R H V (forward) + 2 R H V (dH) + 2 R H V (dW); executed: + 2 R H V
// (tile recompute in the backward, the price of not storing logits).
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "kd_common.cuh"
#include "kd_umma.cuh"

namespace kd {
namespace fused {

using namespace umma;

constexpr int BM = 128, BN = 256, BK = 64, UK = 16;
constexpr int kEpiWarps = 16;                    // 4 per TMEM lane quadrant = 4 per warp scheduler
constexpr int kEpiThreads = 32 * kEpiWarps;      // 512
constexpr int kThreads = 128 + kEpiThreads;      // 640 (<= 96 registers per thread)
constexpr int kColGroups = kEpiWarps / 4;        // column groups working on one step side by side
constexpr uint32_t kABytes = BM * BK * 2;        // 16 KB
constexpr uint32_t kBBytes = BN * BK * 2;        // 32 KB
constexpr uint32_t kBoxMnBytes = 64 * BK * 2;    // one 64(mn) x 64(k) MN-major box = 8 KB
constexpr uint32_t kTmemCols = 512;
constexpr int kStepCols = 32 * kColGroups;       // columns per epilogue step: 32 for each column group (128)
constexpr int kSteps = BN / kStepCols;           // 2
constexpr int kStepBoxes = kStepCols / 64;       // 64-column (128-byte row) TMA boxes per step
constexpr uint32_t kBoxBytes = BM * 64 * 2;      // 16 KB: [128 rows x 64 cols] of a 16-bit type, swizzled 128 B rows
constexpr uint32_t kStepBytes = kStepBoxes * kBoxBytes;  // 32 KB
constexpr int kSchedSlots = 2;  // work-unit queue depth per CTA (pair): the next unit is fetched while one runs
#ifndef KD_G_FP16_DEFAULT
#define KD_G_FP16_DEFAULT 1
#endif
// compile-time A/B switches (speech-distill_b200/build.py --variant NAME -DKD_OPT_...=0 builds a second library that
// tools/k1_abab.sh runs against the default one on the same box; box-to-box spread is larger than most of these)
#ifndef KD_OPT_FWD_ONE_MAX
#define KD_OPT_FWD_ONE_MAX 1  // forward epilogue: one maximum per 32-column piece (statistics + logit-cache reference)
#endif
#ifndef KD_OPT_GC_CURSOR
#define KD_OPT_GC_CURSOR 1    // cached-gradient kernel: per-lane cursors and L2 evict-first loads
#endif
constexpr int kRecFloats = 8;  // forward partial record: m, s1, st, mt, t1, tt, a, z_label
// per-row record one vocabulary slice hands to the cross-rank merge (vocab-parallel mode):
// m, s1, st, mt | t1, tt, a, z_label | y_label, sum p log p, hit value sum, hits
constexpr int kRankRecFloats = 12;

// shared-memory plan of one kernel instantiation: as many operand stages as fit beside the epilogue rings
template <int CG, int YSLOTS, int GSLOTS, int MAXSTAGES = 8>
struct SmemPlan {
  static constexpr uint32_t kBBytesL = kBBytes / CG;
  static constexpr uint32_t kStageL = kABytes + kBBytesL;
  static constexpr uint32_t kRing = (YSLOTS + GSLOTS) * kStepBytes;
  static constexpr uint32_t kBudget = 232448 - 1024 /*alignment slack*/ - 512 /*barriers*/;
  static constexpr int kStagesRaw = (int)((kBudget - kRing) / kStageL);
  static constexpr int kStages = kStagesRaw > MAXSTAGES ? MAXSTAGES : kStagesRaw;
  static constexpr uint32_t kYOff = kStages * kStageL;
  static constexpr uint32_t kGOff = kYOff + YSLOTS * kStepBytes;
  static constexpr uint32_t kBarOff = kGOff + GSLOTS * kStepBytes;
  static constexpr uint32_t kReqBar = 2 * kStages + 4 + 2 * (YSLOTS > 0 ? YSLOTS : 1);  // "draw the next unit"
  static constexpr uint32_t kSchedBar0 = kReqBar + 1;  // first of the queue's full / empty barriers
  static constexpr uint32_t kNumBars = kSchedBar0 + 2 * kSchedSlots;
  static constexpr uint32_t kBytes = kBarOff + 8 * kNumBars + 16 /*tmem slot*/ + 16 /*unit slots*/ + 1024;
  static_assert(kStages >= 2, "operand ring too shallow");  // 2 only for the single-CTA fallback of GradEpi
  static_assert(kBytes <= 232448, "shared memory plan exceeds 227 KB");
};

struct Geom {
  int num_m_blk, num_n_blk, num_k_blk;
  int a_m0, a_k0, b_n0, b_k0;  // element offsets added to the TMA coordinates
  int n_per_unit;              // consecutive n blocks handled by one work unit (same m block)
  int num_units;
  int dynamic;                 // 1: work units are handed out by an atomic counter (see UnitQueue), 0: static stride
  // valid-row compaction (kd_rows.cu): *n_rows = number of live rows (device memory, no host sync).
  // rows_dim = 1: the M dimension is rows -> units whose first row is >= *n_rows are skipped;
  // rows_dim = 2: the K dimension is rows (dW = G^T h) -> only the first ceil(*n_rows / 64) k-blocks are run.
  const int32_t* n_rows;
  int rows_dim;
  int ab_fp16;  // operands A and B are fp16 (dW / dH GEMMs with the fp16 gradient operand), else bf16
  // unit order.  0: the m block is the fast index (units that run side by side share the B tile: forward, dH);
  // 1: the n range is the fast index (they share the A tile: dW = G^T h, whose A tile is a 2 MB column block of
  // the gradient chunk - with m fast all 74 CTA pairs swept the whole 155 MB chunk once per h tile, 0.62 GB of
  // DRAM reads per chunk for 0.16 GB algorithmic, ncu profiles/r01g)
  int n_fast;
};

__host__ __device__ inline int num_ranges_of(const Geom& g) { return (g.num_n_blk + g.n_per_unit - 1) / g.n_per_unit; }
__host__ __device__ inline void decode_unit(const Geom& g, int u, int& m_blk, int& range, int& n_begin, int& n_end) {
  if (g.n_fast) {
    const int nr = num_ranges_of(g);
    m_blk = u / nr;
    range = u - m_blk * nr;
  } else {
    range = u / g.num_m_blk;
    m_blk = u - range * g.num_m_blk;
  }
  n_begin = range * g.n_per_unit;
  n_end = n_begin + g.n_per_unit < g.num_n_blk ? n_begin + g.n_per_unit : g.num_n_blk;
}

// ---------------------------------------------------------------------------------------------
// Epilogue policies.  Each epilogue thread owns accumulator row `row_in_tile` (= its TMEM lane) and
// the column half `half`; a 128 x 256 tile is visited in 4 steps of 64 columns, the thread taking
// columns [64 c + 32 half, +32) of step c.
// ---------------------------------------------------------------------------------------------
struct EpiThread {
  int row_in_tile;         // 0..127 (TMEM lane)
  int cgrp;                // column group 0..kColGroups-1: columns [32 cgrp, +32) of every step
  int lane;                // lane in warp
  int epi_tid;             // 0..kEpiThreads-1
  uint32_t tmem_lane_off;  // lane field of the TMEM address
  uint32_t y_base, g_base; // smem rings (teacher tile in, gradient tile out)
  uint32_t yfull0, yempty0;  // first barrier of each ring set (8 bytes apart)
  const CUtensorMap* tma_g;
};

// hand an accumulator buffer back to the MMA issuer (in the leader CTA for a CTA pair)
template <int CG>
__device__ __forceinline__ void release_tmem(uint32_t bar) {
  if (CG == 2) mbar_arrive_cluster(bar);
  else mbar_arrive(bar);
}

// address of this thread's q-th 16-byte piece (q = 0..3: its 32 columns) inside a step buffer made of
// kStepBoxes [128 x 64] 16-bit boxes with SWIZZLE_128B rows
__device__ __forceinline__ uint32_t step_piece_addr(uint32_t buf, const EpiThread& t, int q) {
  return buf + (uint32_t)(t.cgrp >> 1) * kBoxBytes + (uint32_t)t.row_in_tile * 128u +
         ((uint32_t)(((t.cgrp & 1) * 4 + q) ^ (t.row_in_tile & 7)) << 4);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <typename TY>
__device__ __forceinline__ void load_row16(const TY* __restrict__ p, bool vec_ok, int ncols, float (&f)[16]) {
  // direct global path; ncols = number of in-range columns (<= 16); out-of-range columns read as -inf
  if (vec_ok && ncols >= 16) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      Vec8<TY> v;
      v.load_global(p + 8 * q);
      float t8[8];
      v.unpack(t8);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[8 * q + j] = fmaxf(t8[j], kTeacherFloor);  // -inf teacher entries: see teacher_update
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = j < ncols ? fmaxf(Elem<TY>::to_f(p[j]), kTeacherFloor) : -CUDART_INF_F;
  }
}

// consumer side of the teacher-tile ring: wait for the step buffer, pull this thread's 64 bytes, release
struct YRing {
  int slot = 0;
  uint32_t phase = 0;
  template <int SLOTS>
  __device__ __forceinline__ void take(const EpiThread& t, uint4 (&pk)[4]) {
    mbar_wait(t.yfull0 + 8u * slot, phase);
    const uint32_t buf = t.y_base + (uint32_t)slot * kStepBytes;
#pragma unroll
    for (int q = 0; q < 4; ++q) pk[q] = lds128(step_piece_addr(buf, t, q));
    __syncwarp();
    if (t.lane == 0) mbar_arrive(t.yempty0 + 8u * slot);
    if (++slot == SLOTS) {
      slot = 0;
      phase ^= 1u;
    }
  }
};

// 16 teacher logits of sub-chunk `sub` (0/1) from the 64 packed bytes; columns >= ncols read as -inf
template <typename TY>
__device__ __forceinline__ void unpack_sub(const uint4 (&pk)[4], int sub, int ncols, float (&fy)[16]) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    Vec8<TY> v;
    v.a = sub == 0 ? pk[q] : pk[2 + q];
    if (std::is_same<TY, __nv_bfloat16>::value) {  // -inf -> most negative finite value, two elements per op
      v.a.x = clamp_neg_inf_bf16x2(v.a.x);
      v.a.y = clamp_neg_inf_bf16x2(v.a.y);
      v.a.z = clamp_neg_inf_bf16x2(v.a.z);
      v.a.w = clamp_neg_inf_bf16x2(v.a.w);
    }
    float f8[8];
    v.unpack(f8);
#pragma unroll
    for (int j = 0; j < 8; ++j) fy[8 * q + j] = f8[j];
  }
  if (ncols < 16) {  // ragged vocabulary edge: TMA zero-filled the out-of-range columns
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j >= ncols) fy[j] = -CUDART_INF_F;
  }
}

// Sparse teacher as the epilogues see it (built per call by kd_sparse_prepare_kernel): for every row the K
// entries sorted by vocabulary index, p = softmax(v / tau) over the K entries (distillation_loss.py:94), and
// off[row][t] = number of entries with index < 256 t, so a thread finds the entries of its 256-column tile with
// one 4-byte load (on average K / (V / 256) = 0.1 entries per row and tile at K = 64, V = 152,936).
struct SparseView {
  const int32_t* idx;   // [R][K] ascending; out-of-range indices are dropped to the end
  const float* p;       // [R][K] aligned with idx
  const uint16_t* off;  // [R][off_stride]
  int K, off_stride;
};

// fp16 gradient operand (KD_G_FP16): G is scaled by a power of two S so that its entries sit in fp16's normal
// range (|G| <= ~coef / N: 2.4e-4 at N = 4096 would be fine, a probability of 1e-5 / N would not), rounded to fp16
// (11 significand bits instead of bf16's 8) and multiplied with fp16 copies of h and W (exact for bf16 values inside
// fp16's exponent range); the dW / dH epilogues multiply by 1 / S.  S = 2^(12 + ceil(log2 N) - ceil(log2 max|coef|)):
// the largest entry is <= 2^13, the smallest normal one corresponds to a probability of ~1.5e-8.
__device__ __forceinline__ float g_operand_scale(const int32_t* n_norm, const float* coef, float tau) {
  const int nn = *n_norm;
  const float cmax = fmaxf(fmaxf(fabsf(coef[0]), fabsf(coef[1]) * tau), 1e-30f);
  const float e = 12.f + ceilf(log2f((float)(nn > 0 ? nn : 1))) - ceilf(log2f(cmax));
  return exp2f(fminf(fmaxf(e, -100.f), 100.f));
}

// label of a row in this call's column space: row_target - label_off when it falls inside [0, V), a far-away
// sentinel when another vocabulary slice owns it (vocab-parallel); `valid` is the row predicate itself
constexpr int kLabelElsewhere = -(1 << 30);
__device__ __forceinline__ int local_target(int row_target, int label_off, int V, bool& valid) {
  valid = row_target >= 0;
  const int t = row_target - label_off;
  return (valid && t >= 0 && t < V) ? t : kLabelElsewhere;
}

// ---- forward: online statistics -------------------------------------------------------------
struct FwdParams {
  const int32_t* row_target;
  const void* y;
  int64_t y_stride;
  int y_vec_ok;
  int R, V;
  float inv_tau;
  float* partials;  // [num_ranges * kColGroups][R][kRecFloats]
  int label_off;    // vocab-parallel: this call's columns are vocabulary [label_off, label_off + V)
  SparseView sp;    // sparse teacher (index-sorted entries + per-tile offsets), unused otherwise
  int debug_skip_math;  // KD_DEBUG_SKIP_MATH=1: bring-up knob that measures the pipeline without the epilogue math
  // logit cache (see LogitCache): the first zc_tiles 256-column tiles of every row are kept, encoded, for the
  // backward; zc_tiles = 0 switches the store path off
  int zc_tiles;
  int16_t* zc_ref;  // [zc_tiles * 8][R] piece references
};

template <typename TY, bool DENSE, bool TAU2, bool Y_TMA, bool SPARSE = false>
struct FwdEpi {
  static_assert(!(DENSE && SPARSE), "one teacher kind per instantiation");
  using Params = FwdParams;
  static constexpr bool kUseYRing = DENSE && Y_TMA;
  // two teacher-tile slots + one staging buffer for the logit-cache store (32 KB each) + four operand stages.
  // (Round 1, without the cache store: three teacher slots + four stages measured 4 % faster than 2 + 5.)
  static constexpr int kYSlots = kUseYRing ? 2 : 0;
  static constexpr int kGSlots = 1;
  static constexpr int kMaxStages = 8;
  static constexpr int kBoundThreads = kThreads;
  const Params& p;
  EpiThread t;
  int row, target, range, m0;
  bool valid;
  float m, s1, st, mt, t1, tt, a, zl;
  YRing ring;

  __device__ FwdEpi(const Params& p_, const EpiThread& t_) : p(p_), t(t_) {}

  __device__ void begin_unit(const Geom&, int m0_, int range_) {
    m0 = m0_;
    row = m0 + t.row_in_tile;
    range = range_;
    target = local_target(row < p.R ? p.row_target[row] : -1, p.label_off, p.V, valid);
    m = mt = -CUDART_INF_F;
    s1 = st = t1 = tt = a = zl = 0.f;
  }

  // 16 columns of one row: label pick-up, student statistics, teacher statistics
  __device__ __forceinline__ void sub_chunk(const uint32_t (&raw)[16], float (&fy)[16], int col0, int ncols) {
    float fz[16];
    if (ncols >= 16) {  // common case: no per-element predicates in the hot loop
#pragma unroll
      for (int j = 0; j < 16; ++j) fz[j] = __uint_as_float(raw[j]);
    } else {            // ragged vocabulary edge (last column tile only)
#pragma unroll
      for (int j = 0; j < 16; ++j) fz[j] = j < ncols ? __uint_as_float(raw[j]) : -CUDART_INF_F;
    }
    const unsigned d = (unsigned)(target - col0);
    if (d < 16u) {  // the label column lives in one sub-chunk per row: rare
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j == (int)d) zl = fz[j];
    }
#if KD_OPT_FWD_ONE_MAX
    student_add<TAU2, 16>(fz, 16, p.inv_tau, m, s1, st);  // tile() has raised m over the whole 32-column piece
#else
    student_update<TAU2, 16>(fz, 16, p.inv_tau, m, s1, st);
#endif
    if (DENSE) {
      if (ncols >= 16) {
        teacher_update<TAU2, 16, false>(fy, fz, 16, p.inv_tau, mt, t1, tt, a);
      } else {  // masked columns hold -inf on both sides: keep their 0 * inf out of the cross term
        teacher_update<TAU2, 16, true>(fy, fz, 16, p.inv_tau, mt, t1, tt, a);
      }
    }
  }

  // logit cache: 16 logits -> fp16 of (z - ref), two 16-byte pieces
  __device__ __forceinline__ void encode16(const uint32_t (&raw)[16], float ref, uint4& lo, uint4& hi) {
    float t8[8];
    Vec8<__half> v;
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = __uint_as_float(raw[j]) - ref;
    v.pack(t8);
    lo = v.a;
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = __uint_as_float(raw[8 + j]) - ref;
    v.pack(t8);
    hi = v.a;
  }

  template <int CG>
  __device__ void tile(const Geom& g, int n_blk, uint32_t tmem_acc, uint32_t tempty_bar) {
    const int col_base = g.b_n0 + n_blk * BN + t.cgrp * 32;
    const bool zc_on = n_blk < p.zc_tiles;  // uniform over the CTA pair
    int e_beg = 0, e_end = 0;
    if (SPARSE && valid) {  // this row's teacher entries inside the tile
      const uint16_t* o = p.sp.off + (size_t)row * p.sp.off_stride + (g.b_n0 / BN + n_blk);
      e_beg = o[0];
      e_end = o[1];
    }
#pragma unroll 1
    for (int c = 0; c < kSteps; ++c) {
      uint32_t raw0[16], raw1[16];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent row predicates
      const uint32_t taddr = tmem_acc + t.tmem_lane_off + (uint32_t)(c * kStepCols + t.cgrp * 32);
      tmem_ld16(taddr, raw0);
      tmem_ld16(taddr + 16u, raw1);
      uint4 pk[4];
      if (kUseYRing) ring.template take<kYSlots>(t, pk);
      tmem_ld_wait();
      if (c == kSteps - 1) {  // accumulator fully read: hand the TMEM buffer back before the math
        fence_before_sync();
        __syncwarp();
        if (t.lane == 0) release_tmem<CG>(tempty_bar);
      }
      const int col0 = col_base + c * kStepCols;
      const int nrem = p.V - col0;
      const bool live = valid && nrem > 0 && !p.debug_skip_math;
      float ref = 0.f;
#if !KD_OPT_FWD_ONE_MAX
      float piece_max = 0.f;
#endif
      // logit-cache staging: every epilogue warp owns a [32 rows x 32 columns] fp16 box (2 KB, 64-byte swizzled rows)
      // of the 32 KB slot and stores it with its own TMA store - no CTA-wide barrier on the store path
      const uint32_t wbuf = t.g_base + (uint32_t)(t.epi_tid >> 5) * 2048u;
      const uint32_t wrow = wbuf + (uint32_t)t.lane * 64u;
      const uint32_t wswz = ((uint32_t)t.lane >> 1) & 3u;
      if (live) {
        // maximum of this thread's 32 columns, taken once: it raises the running maximum of the online statistics
        // and, rounded up to an integer, is the piece reference of the logit cache - the smallest integer >= every
        // logit this thread has seen in its columns of this row so far.  It is <= row max + 1, so
        // |z - ref| <= (row max - z) + 1: the fp16 rounding error of an entry shrinks with its probability.
        float vm = -CUDART_INF_F;
        if (nrem >= 32) {
#pragma unroll
          for (int j = 0; j < 16; ++j) vm = fmaxf(vm, fmaxf(__uint_as_float(raw0[j]), __uint_as_float(raw1[j])));
        } else {  // ragged vocabulary edge
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < nrem) vm = fmaxf(vm, __uint_as_float(raw0[j]));
            if (16 + j < nrem) vm = fmaxf(vm, __uint_as_float(raw1[j]));
          }
        }
#if KD_OPT_FWD_ONE_MAX
        student_raise_max(vm, p.inv_tau, m, s1, st);
#else
        piece_max = fmaxf(vm, m);
#endif
      }
      if (zc_on) {
        // the previous step's store (issued by lane 0 a whole TMEM load + teacher-tile wait ago) must have read the box
        if (t.lane == 0) bulk_wait_read_all();
        __syncwarp();
        if (live) {
#if KD_OPT_FWD_ONE_MAX
          ref = fminf(fmaxf(ceilf(m), -32000.f), 32000.f);
#else
          ref = fminf(fmaxf(ceilf(piece_max), -32000.f), 32000.f);
#endif
          p.zc_ref[(size_t)(n_blk * 8 + c * kColGroups + t.cgrp) * p.R + row] = (int16_t)ref;
        } else {  // rows that are not scored (and columns past V) store zeros
#pragma unroll
          for (int q = 0; q < 4; ++q) sts128(wrow + (((uint32_t)q ^ wswz) << 4), make_uint4(0, 0, 0, 0));
        }
      }
      if (live) {
        if (SPARSE) {  // cross term sum_k p_k z[i_k] (distillation_loss.py:101-106): rare, predicated pick-up
          for (int e = e_beg; e < e_end; ++e) {
            const unsigned d = (unsigned)(p.sp.idx[(size_t)row * p.sp.K + e] - col0);
            if (d < 32u) {
              float z = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j == (int)(d & 15u)) z = __uint_as_float(d < 16u ? raw0[j] : raw1[j]);
              a = fmaf(p.sp.p[(size_t)row * p.sp.K + e], z, a);
            }
          }
        }
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int nc = nrem - 16 * sub < 16 ? nrem - 16 * sub : 16;
          if (nc <= 0) {  // (ragged edge) nothing to score, but the cache piece must still be defined
            if (zc_on) {
              sts128(wrow + ((2u ^ wswz) << 4), make_uint4(0, 0, 0, 0));
              sts128(wrow + ((3u ^ wswz) << 4), make_uint4(0, 0, 0, 0));
            }
            break;
          }
          float fy[16];
          if (DENSE) {
            if (kUseYRing) {
              unpack_sub<__nv_bfloat16>(pk, sub, nc, fy);
            } else {
              const TY* yp = reinterpret_cast<const TY*>(p.y) + (int64_t)row * p.y_stride + col0 + 16 * sub;
              load_row16<TY>(yp, p.y_vec_ok != 0, nc, fy);
            }
          }
          if (sub == 0) {
            if (zc_on) {
              uint4 lo, hi;
              encode16(raw0, ref, lo, hi);
              sts128(wrow + ((0u ^ wswz) << 4), lo);
              sts128(wrow + ((1u ^ wswz) << 4), hi);
            }
            sub_chunk(raw0, fy, col0, nc);
          } else {
            if (zc_on) {
              uint4 lo, hi;
              encode16(raw1, ref, lo, hi);
              sts128(wrow + ((2u ^ wswz) << 4), lo);
              sts128(wrow + ((3u ^ wswz) << 4), hi);
            }
            sub_chunk(raw1, fy, col0 + 16, nc);
          }
        }
      }
      if (zc_on) {
        fence_proxy_async_smem();
        __syncwarp();
        if (t.lane == 0) {
          tma_store_2d(t.tma_g, wbuf, n_blk * BN + c * kStepCols + t.cgrp * 32, m0 + (t.row_in_tile & ~31));
          bulk_commit();
        }
      }
    }
  }

  __device__ void end_unit(const Geom&) {
    if (row < p.R) {
      float* rec = p.partials + ((size_t)(range * kColGroups + t.cgrp) * p.R + row) * kRecFloats;
      *reinterpret_cast<float4*>(rec) = make_float4(m, s1, st, mt);
      *reinterpret_cast<float4*>(rec + 4) = make_float4(t1, tt, a, zl);
    }
  }
  __device__ void finish() {
    if (t.lane == 0) bulk_wait_all();  // every warp issued its own stores
  }
};

// ---- backward: gradient tile ------------------------------------------------------------------
struct GradParams {
  const int32_t* row_target;
  const float* row_stats;  // [R][4] = LSE1, LSE_tau, LSEteacher_tau, valid
  const void* y;
  int64_t y_stride;
  int y_vec_ok;
  int R, V;
  float tau;
  int use_kl;  // 0: CE only (no teacher)
  const int32_t* n_norm;
  const float* coef;  // device float[2]: weight of d(sum CE) and of tau^2 d(sum KL) in the returned gradient
  int v0;             // first vocabulary column of this chunk; scratch column j <-> vocabulary index v0 + j
  int label_off;      // vocab-parallel: column v of this call is vocabulary index label_off + v
  int g_fp16;         // write G as fp16 scaled by g_operand_scale() instead of bf16
  SparseView sp;      // sparse teacher, see FwdParams
};

// COPY = true turns the policy into "logits tile -> bf16 -> scratch" (no soft-max arithmetic): the teacher LM head
// of kd_linear_bf16, which shares the swizzled smem staging + TMA store path of the gradient tile.
template <typename TY, bool DENSE, bool TAU2, bool Y_TMA, bool SPARSE = false, bool COPY = false>
struct GradEpi {
  static_assert(!(DENSE && SPARSE), "one teacher kind per instantiation");
  static_assert(!COPY || (!DENSE && !SPARSE), "the copy policy has no teacher");
  using Params = GradParams;
  static constexpr bool kUseYRing = DENSE && Y_TMA;
  static constexpr int kYSlots = kUseYRing ? 2 : 0;
  static constexpr int kGSlots = 1;
  static constexpr int kMaxStages = 8;
  static constexpr int kBoundThreads = kThreads;
  const Params& p;
  EpiThread t;
  int row, target, m0;
  bool valid;
  float c1, c2, c_tau, off1, offt, offy, half_off1, k_tau;
  YRing ring;

  __device__ GradEpi(const Params& p_, const EpiThread& t_) : p(p_), t(t_) {
    if (COPY) return;
    const int nn = *p.n_norm;
    float inv_n = nn > 0 ? 1.0f / (float)nn : 0.f;
    if (p.g_fp16) inv_n *= g_operand_scale(p.n_norm, p.coef, p.tau);  // power of two: exact
    c1 = p.coef[0] * inv_n;
    c2 = p.use_kl ? p.coef[1] * p.tau * inv_n : 0.f;
    c_tau = kLog2e / p.tau;
  }

  __device__ void begin_unit(const Geom&, int m0_, int) {
    m0 = m0_;
    row = m0 + t.row_in_tile;
    if (COPY) {
      valid = row < p.R;
      target = kLabelElsewhere;
      return;
    }
    target = local_target(row < p.R ? p.row_target[row] : -1, p.label_off, p.V, valid);
    if (valid) {
      const float4 rs = *reinterpret_cast<const float4*>(p.row_stats + (size_t)row * 4);
      off1 = rs.x * kLog2e;
      offt = rs.y * kLog2e;
      offy = rs.z * kLog2e;
      half_off1 = 0.5f * off1;
      k_tau = c2 * ex2(half_off1 - offt);
    }
  }

  // gradient of 16 columns -> two packed 16-byte pieces
  __device__ __forceinline__ void sub_chunk(const uint32_t (&raw)[16], const float (&fy)[16], int col0, int ncols,
                                            int e_beg, int e_end, uint4& lo, uint4& hi) {
    float gq[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float z = __uint_as_float(raw[j]);
      float gi;
      if (COPY) {
        gi = z;
      } else if (TAU2) {
        const float e = ex2(fmaf(z, c_tau, -half_off1));
        gi = e * fmaf(e, c1, k_tau);
      } else {
        gi = c1 * ex2(fmaf(z, kLog2e, -off1)) + c2 * ex2(fmaf(z, c_tau, -offt));
      }
      if (DENSE) gi = fmaf(-c2, ex2(fmaf(fy[j], c_tau, -offy)), gi);
      gq[j] = gi;
    }
    if (ncols < 16) {  // ragged vocabulary edge: columns beyond V contribute nothing
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j >= ncols) gq[j] = 0.f;
    }
    if (SPARSE) {  // - c2 P with P = scatter(i_k, p_k); duplicate indices accumulate
      for (int e = e_beg; e < e_end; ++e) {
        const unsigned ds = (unsigned)(p.sp.idx[(size_t)row * p.sp.K + e] - col0);
        if (ds < 16u) {
          const float pk = c2 * p.sp.p[(size_t)row * p.sp.K + e];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j == (int)ds) gq[j] -= pk;
        }
      }
    }
    const unsigned d = (unsigned)(target - col0);
    if (d < 16u) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j == (int)d) gq[j] -= c1;
    }
    float t8[8];
    if (!COPY && p.g_fp16) {
      Vec8<__half> v;
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = gq[j];
      v.pack(t8);
      lo = v.a;
#pragma unroll
      for (int j = 0; j < 8; ++j) t8[j] = gq[8 + j];
      v.pack(t8);
      hi = v.a;
      return;
    }
    Vec8<__nv_bfloat16> v;
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = gq[j];
    v.pack(t8);
    lo = v.a;
#pragma unroll
    for (int j = 0; j < 8; ++j) t8[j] = gq[8 + j];
    v.pack(t8);
    hi = v.a;
  }

  template <int CG>
  __device__ void tile(const Geom& g, int n_blk, uint32_t tmem_acc, uint32_t tempty_bar) {
    const int jtile = n_blk * BN;  // first scratch column of this tile
    int e_beg = 0, e_end = 0;
    if (SPARSE && valid) {
      const uint16_t* o = p.sp.off + (size_t)row * p.sp.off_stride + ((p.v0 + jtile) / BN);
      e_beg = o[0];
      e_end = o[1];
    }
#pragma unroll 1
    for (int c = 0; c < kSteps; ++c) {
      uint32_t raw0[16], raw1[16];
      __syncwarp();
      const uint32_t taddr = tmem_acc + t.tmem_lane_off + (uint32_t)(c * kStepCols + t.cgrp * 32);
      tmem_ld16(taddr, raw0);
      tmem_ld16(taddr + 16u, raw1);
      uint4 pk[4];
      if (kUseYRing) ring.template take<kYSlots>(t, pk);
      tmem_ld_wait();
      if (c == kSteps - 1) {
        fence_before_sync();
        __syncwarp();
        if (t.lane == 0) release_tmem<CG>(tempty_bar);
      }
      const int j0 = jtile + c * kStepCols + t.cgrp * 32;
      const int col0 = p.v0 + j0;
      const int nrem = p.V - col0;
      uint4 out[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) out[q] = make_uint4(0, 0, 0, 0);
      if (valid && nrem > 0) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int nc = nrem - 16 * sub < 16 ? nrem - 16 * sub : 16;
          if (nc <= 0) break;
          float fy[16];
          if (DENSE) {
            if (kUseYRing) {
              unpack_sub<__nv_bfloat16>(pk, sub, nc, fy);
            } else {
              const TY* yp = reinterpret_cast<const TY*>(p.y) + (int64_t)row * p.y_stride + col0 + 16 * sub;
              load_row16<TY>(yp, p.y_vec_ok != 0, nc, fy);
            }
          }
          if (sub == 0) sub_chunk(raw0, fy, col0, nc, e_beg, e_end, out[0], out[1]);
          else sub_chunk(raw1, fy, col0 + 16, nc, e_beg, e_end, out[2], out[3]);
        }
      }
      // stage the [128 x kStepCols] bf16 step in shared memory (swizzled rows) and hand it to TMA stores.
      // One buffer: the elected thread first waits until the previous step's stores have finished
      // reading it, a barrier publishes that, everybody writes, a second barrier publishes the writes.
      if (t.epi_tid == 0) bulk_wait_read_all();
      named_bar_sync(1, kEpiThreads);
#pragma unroll
      for (int q = 0; q < 4; ++q) sts128(step_piece_addr(t.g_base, t, q), out[q]);
      fence_proxy_async_smem();
      named_bar_sync(2, kEpiThreads);
      if (t.epi_tid == 0) {
#pragma unroll
        for (int b = 0; b < kStepBoxes; ++b)
          tma_store_2d(t.tma_g, t.g_base + b * kBoxBytes, jtile + c * kStepCols + 64 * b, m0);
        bulk_commit();
      }
    }
  }

  __device__ void end_unit(const Geom&) {}
  __device__ void finish() {
    if (t.epi_tid == 0) bulk_wait_all();
  }
};

// ---- teacher head feeding the top-k compaction (train.py:60-94, extract_teacher_logits.py:109-129) -------------
// logits tile -> bf16 -> row-block scratch (per-warp TMA stores, as the forward's logit cache) and, from the SAME
// rounded values, what the selection kernel (kd_topk.cu, kd_head_select_kernel) needs so that it never re-reads the
// row: the maximum of every 32-column piece (bf16, exact) and this thread's online (max, sum exp) over its pieces
// of the work unit.  The k-th largest piece maximum bounds the k-th logit from below, so the selection scans only
// the ~k pieces (64 bytes each) whose maximum reaches it instead of the row's 2 V bytes.
struct HeadParams {
  int R, V;
  __nv_bfloat16* pmax;  // [R][pmax_stride], piece j = columns [32 j, 32 j + 32); pieces past V hold -inf
  int pmax_stride;      // = 8 * number of 256-column tiles
  float2* part;         // [R][part_stride]: (m, s) of (vocabulary range, column group), s = sum 2^((x - m) log2 e)
  int part_stride;      // = kColGroups * number of ranges
};

struct HeadEpi {
  using Params = HeadParams;
  static constexpr int kYSlots = 0;
  static constexpr int kGSlots = 1;
#ifndef KD_HEAD_STAGES
#define KD_HEAD_STAGES 8  // as many operand stages as fit (6); 4 and 5 measured slower (tools/head_parts.py, r2i)
#endif
  static constexpr int kMaxStages = KD_HEAD_STAGES;
  static constexpr int kBoundThreads = kThreads;
  const Params& p;
  EpiThread t;
  int row, range, m0;
  bool valid;
  float m, s;

  __device__ HeadEpi(const Params& p_, const EpiThread& t_) : p(p_), t(t_) {}

  __device__ void begin_unit(const Geom&, int m0_, int range_) {
    m0 = m0_;
    row = m0 + t.row_in_tile;
    range = range_;
    valid = row < p.R;
    m = -CUDART_INF_F;
    s = 0.f;
  }

  template <int CG>
  __device__ void tile(const Geom& g, int n_blk, uint32_t tmem_acc, uint32_t tempty_bar) {
#pragma unroll 1
    for (int c = 0; c < kSteps; ++c) {
      uint32_t raw0[16], raw1[16];
      __syncwarp();
      const uint32_t taddr = tmem_acc + t.tmem_lane_off + (uint32_t)(c * kStepCols + t.cgrp * 32);
      tmem_ld16(taddr, raw0);
      tmem_ld16(taddr + 16u, raw1);
      tmem_ld_wait();
      if (c == kSteps - 1) {
        fence_before_sync();
        __syncwarp();
        if (t.lane == 0) release_tmem<CG>(tempty_bar);
      }
      const int col0 = n_blk * BN + c * kStepCols + t.cgrp * 32;
      const int nrem = p.V - col0;
      // round to bf16 once; everything below (store, maximum, exponentials) sees the rounded values
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(raw0[2 * j]), __uint_as_float(raw0[2 * j + 1]));
        __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(raw1[2 * j]), __uint_as_float(raw1[2 * j + 1]));
        w[j] = *reinterpret_cast<uint32_t*>(&h0);
        w[8 + j] = *reinterpret_cast<uint32_t*>(&h1);
      }
      // per-warp [32 rows x 32 columns] box (2 KB, 64-byte swizzled rows) -> TMA store; no CTA-wide barrier
      const uint32_t wbuf = t.g_base + (uint32_t)(t.epi_tid >> 5) * 2048u;
      const uint32_t wrow = wbuf + (uint32_t)t.lane * 64u;
      const uint32_t wswz = ((uint32_t)t.lane >> 1) & 3u;
      if (t.lane == 0) bulk_wait_read_all();  // the previous step's store has read the box
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sts128(wrow + (((uint32_t)q ^ wswz) << 4), make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
      fence_proxy_async_smem();
      __syncwarp();
      if (t.lane == 0 && nrem > 0) {  // rows >= R and columns >= V are clipped by the tensor map
        tma_store_2d(t.tma_g, wbuf, col0, m0 + (t.row_in_tile & ~31));
        bulk_commit();
      }
      if (!valid) continue;
      __nv_bfloat16* pm = p.pmax + (size_t)row * p.pmax_stride + (col0 >> 5);
      if (nrem <= 0) {  // piece past the vocabulary (last tile only)
        *pm = __float2bfloat16(-CUDART_INF_F);
        continue;
      }
      float f[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f[2 * j] = __uint_as_float(w[j] << 16);
        f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
      if (nrem < 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j >= nrem) f[j] = -CUDART_INF_F;
      }
      float vm = f[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) vm = fmaxf(vm, f[j]);
      *pm = __float2bfloat16(vm);  // exact: vm is one of the bf16 values
      if (vm > m) {
        s *= exp_diff(m, vm, kLog2e);
        m = vm;
      }
      if (m != -CUDART_INF_F) {
        const float off = m * kLog2e;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          p0 += ex2(fmaf(f[j], kLog2e, -off));
          p1 += ex2(fmaf(f[j + 1], kLog2e, -off));
          p2 += ex2(fmaf(f[j + 2], kLog2e, -off));
          p3 += ex2(fmaf(f[j + 3], kLog2e, -off));
        }
        s += (p0 + p1) + (p2 + p3);
      }
    }
  }

  __device__ void end_unit(const Geom&) {
    if (valid) p.part[(size_t)row * p.part_stride + range * kColGroups + t.cgrp] = make_float2(m, s);
  }
  __device__ void finish() {
    if (t.lane == 0) bulk_wait_all();
  }
};

// ---- backward from the logit cache: gradient chunk without the recompute GEMM ---------------------
// Logit cache (written by FwdEpi, read here).  The forward keeps the first `tiles` 256-column tiles of the logits
// of every row in a buffer whose size is a caller-given constant (so the path's peak memory stays independent
// of V: a larger vocabulary simply has a smaller cached fraction, the rest is recomputed by GradEpi):
//   z16 [R][tiles * 256]  fp16 of (z - ref),   ref [tiles * 8][R] int16: one integer reference per row and
//   32-column piece, chosen by the forward as ceil(running row max of the thread that owns the piece).
// Because ref <= row max + 1, |z - ref| <= (row max - z) + 1 and the fp16 rounding error of the decoded logit is
// <= 2^-11 (d + 1) for an entry d below the row max, i.e. an absolute error of e^-d (d + 1) 2^-11 <= 5e-4 of the
// LARGEST probability of the row for the entry's probability - an order below the fp16 rounding of G itself
// (plain fp16 logits would carry 2^-12 |z| regardless of the entry's weight: 2e-3 at |z| = 8).
// With the cache the backward's gradient chunk is an HBM-bound elementwise kernel (read z16 2 B + teacher 2 B,
// write G 2 B per element) that runs beside the dW / dH GEMMs of the previous chunk instead of a fourth GEMM:
// executed FLOPs drop from 8 R H V to the algorithmic 6 R H V.
struct LogitCache {
  int tiles;            // cached 256-column tiles (0 = no cache)
  __half* z16;          // [R][tiles * 256]
  int16_t* ref;         // [tiles * 8][R]
  int64_t z_stride;     // tiles * 256
};

struct GradCachedParams {
  LogitCache zc;
  const void* y;
  int64_t y_stride;
  int y_vec_ok;
  const int32_t* row_target;
  const float* row_stats;
  const int32_t* n_norm;
  const float* coef;
  const int32_t* n_rows;  // live rows after compaction, or null
  float tau;
  int use_kl, g_fp16;
  int R, V, v0, cols_pad, label_off;  // chunk = vocabulary columns [v0, v0 + cols_pad), cols_pad % 256 == 0
  void* G;                            // [R][g_stride] 16-bit
  int64_t g_stride;
  SparseView sp;
};

// One warp per row strip, two 256-column pieces in flight per warp.  The kernel mostly runs BESIDE a dW / dH GEMM
// CTA (640 threads x 64 registers, 192 KB of shared memory), which leaves an SM 24 K registers: 512 threads x 48
// registers fit once, i.e. 16 warps x 2 KB in flight per SM - with one 256-thread x 64-register CTA (the first
// version) the kernel was latency-bound at 360 us per chunk and set the pace of the whole backward
// (profiles/r02a_bwd_trace.log); alone on the GPU two CTAs per SM run.
constexpr int kGcThreads = 512;
constexpr int kGcUnroll = 2;
constexpr int kGcMaxRegs = 48;

struct GcRow {  // per-row constants of the gradient formula (see GradEpi)
  float c1, c2, c_tau, off1, offt, offy, half_off1, k_tau;
  int target;
};

// gradient of 8 consecutive columns starting at vocabulary column `col` (local to this call); nrem = V - col
template <typename TY, bool DENSE, bool TAU2, bool SPARSE>
__device__ __forceinline__ uint4 gc_piece8(const GradCachedParams& p, const GcRow& r, int row, int col, int nrem,
                                           uint4 zv, float ref, const float (&fy)[8]) {
  float fz[8], gq[8];
  Vec8<__half> v;
  v.a = zv;
  v.unpack(fz);
  const float ref_tau = fmaf(ref, r.c_tau, -r.half_off1);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float gi;
    if (TAU2) {
      const float e = ex2(fmaf(fz[i], r.c_tau, ref_tau));
      gi = e * fmaf(e, r.c1, r.k_tau);
    } else {
      const float z = fz[i] + ref;
      gi = r.c1 * ex2(fmaf(z, kLog2e, -r.off1)) + r.c2 * ex2(fmaf(z, r.c_tau, -r.offt));
    }
    if (DENSE) gi = fmaf(-r.c2, ex2(fmaf(fy[i], r.c_tau, -r.offy)), gi);
    gq[i] = i < nrem ? gi : 0.f;
  }
  if (SPARSE) {  // - c2 P with P = scatter(i_k, p_k); duplicate indices accumulate
    const uint16_t* o = p.sp.off + (size_t)row * p.sp.off_stride + col / BN;  // uniform over the warp's piece
    const int e_beg = o[0], e_end = o[1];
    for (int e = e_beg; e < e_end; ++e) {
      const unsigned ds = (unsigned)(p.sp.idx[(size_t)row * p.sp.K + e] - col);
      if (ds < 8u) {
        const float pk = r.c2 * p.sp.p[(size_t)row * p.sp.K + e];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i == (int)ds) gq[i] -= pk;
      }
    }
  }
  const unsigned d = (unsigned)(r.target - col);
  if (d < 8u) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i == (int)d) gq[i] -= r.c1;
  }
  if (p.g_fp16) {
    Vec8<__half> o;
    o.pack(gq);
    return o.a;
  }
  Vec8<__nv_bfloat16> o;
  o.pack(gq);
  return o.a;
}

// the ragged pieces of a row (vocabulary edge, unaligned teacher rows): rare, kept out of the streaming loop
template <typename TY, bool DENSE, bool TAU2, bool SPARSE>
__device__ __noinline__ void gc_piece_slow(const GradCachedParams& p, const GcRow r, int row, int j) {
  const int col = p.v0 + j;
  const int nrem = p.V - col;
  uint16_t* grow = reinterpret_cast<uint16_t*>(p.G) + (size_t)row * p.g_stride;
  if (nrem <= 0) {
    *reinterpret_cast<uint4*>(grow + j) = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4 zv = ldg_stream(p.zc.z16 + (size_t)row * p.zc.z_stride + col);
  const float ref = (float)p.zc.ref[(size_t)(col >> 5) * p.R + row];
  float fy[8];
  if (DENSE) {
    const TY* yrow = reinterpret_cast<const TY*>(p.y) + (int64_t)row * p.y_stride + col;
#pragma unroll
    for (int i = 0; i < 8; ++i) fy[i] = i < nrem ? fmaxf(Elem<TY>::to_f(yrow[i]), kTeacherFloor) : -CUDART_INF_F;
  }
  *reinterpret_cast<uint4*>(grow + j) = gc_piece8<TY, DENSE, TAU2, SPARSE>(p, r, row, col, nrem, zv, ref, fy);
}

template <typename TY, bool DENSE, bool TAU2, bool SPARSE>
__global__ void __maxnreg__(kGcMaxRegs) kd_grad_cached_kernel(const __grid_constant__ GradCachedParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  GcRow r;
  {
    const int nn = *p.n_norm;
    float inv_n = nn > 0 ? 1.0f / (float)nn : 0.f;
    if (p.g_fp16) inv_n *= g_operand_scale(p.n_norm, p.coef, p.tau);
    r.c1 = p.coef[0] * inv_n;
    r.c2 = p.use_kl ? p.coef[1] * p.tau * inv_n : 0.f;
    r.c_tau = kLog2e / p.tau;
  }
  int r_end = p.R;
  if (p.n_rows != nullptr) {  // compacted rows: everything behind the last live 256-row tile is never read
    const int live = (*p.n_rows + 2 * BM - 1) / (2 * BM) * (2 * BM);
    r_end = live < r_end ? live : r_end;
  }
  const uint64_t pol_stream = l2_policy_evict_first();
  // columns [0, fast_end) of the chunk take the streaming path: whole 16-byte vectors of z16, teacher and G
  int fast_end = p.V - p.v0 < p.cols_pad ? p.V - p.v0 : p.cols_pad;
  fast_end = (DENSE && !p.y_vec_ok) ? 0 : (fast_end & ~7);
  for (int row = blockIdx.x * (kGcThreads / 32) + warp; row < r_end; row += gridDim.x * (kGcThreads / 32)) {
    bool valid;
    r.target = local_target(p.row_target[row], p.label_off, p.V, valid);
    uint16_t* grow = reinterpret_cast<uint16_t*>(p.G) + (size_t)row * p.g_stride;
    if (!valid) {  // rows that are not scored contribute nothing: zeros, no reads
      for (int j = lane * 8; j < p.cols_pad; j += 256) *reinterpret_cast<uint4*>(grow + j) = make_uint4(0, 0, 0, 0);
      continue;
    }
    {
      const float4 rs = *reinterpret_cast<const float4*>(p.row_stats + (size_t)row * 4);
      r.off1 = rs.x * kLog2e;
      r.offt = rs.y * kLog2e;
      r.offy = rs.z * kLog2e;
      r.half_off1 = 0.5f * r.off1;
      r.k_tau = r.c2 * ex2(r.half_off1 - r.offt);
    }
    // per-lane cursors, advanced by constants (lane l owns columns 8 l .. 8 l + 7 of every 256-column piece)
    const __half* zp = p.zc.z16 + (size_t)row * p.zc.z_stride + p.v0 + lane * 8;
    const int16_t* refp = p.zc.ref + (size_t)((p.v0 >> 5) + (lane >> 2)) * p.R + row;
    const TY* yp = DENSE ? reinterpret_cast<const TY*>(p.y) + (int64_t)row * p.y_stride + p.v0 + lane * 8 : nullptr;
    uint16_t* gp = grow + lane * 8;
    const size_t ref_step = (size_t)8 * p.R;  // one 256-column piece = 8 reference pieces
    for (int j0 = lane * 8; j0 < p.cols_pad; j0 += 256 * kGcUnroll, zp += 256 * kGcUnroll, yp += DENSE ? 256 * kGcUnroll : 0,
             gp += 256 * kGcUnroll, refp += kGcUnroll * ref_step) {
      if (j0 + 256 * (kGcUnroll - 1) + 8 <= fast_end) {
        uint4 zv[kGcUnroll];
        Vec8<TY> yv[kGcUnroll];
        float ref[kGcUnroll];
        // all loads first (memory-level parallelism), then the arithmetic; the streams are dead after this read:
        // evict-first keeps them from pushing the gradient chunk, which the GEMMs re-read twice, out of L2
#pragma unroll
        for (int u = 0; u < kGcUnroll; ++u) {
#if KD_OPT_GC_CURSOR
          zv[u] = ldg_hint(zp + u * 256, pol_stream);
          ref[u] = (float)refp[u * ref_step];
          if (DENSE) yv[u].load_global_hint(yp + u * 256, pol_stream);
#else
          zv[u] = ldg_stream(zp + u * 256);
          ref[u] = (float)refp[u * ref_step];
          if (DENSE) yv[u].load_global(yp + u * 256);
#endif
        }
#pragma unroll
        for (int u = 0; u < kGcUnroll; ++u) {
          const int j = j0 + u * 256;
          float fy[8];
          if (DENSE) {
            if (std::is_same<TY, __nv_bfloat16>::value) {  // -inf -> most negative finite value, two elements per op
              yv[u].a.x = clamp_neg_inf_bf16x2(yv[u].a.x);
              yv[u].a.y = clamp_neg_inf_bf16x2(yv[u].a.y);
              yv[u].a.z = clamp_neg_inf_bf16x2(yv[u].a.z);
              yv[u].a.w = clamp_neg_inf_bf16x2(yv[u].a.w);
              yv[u].unpack(fy);
            } else {
              yv[u].unpack(fy);
#pragma unroll
              for (int i = 0; i < 8; ++i) fy[i] = fmaxf(fy[i], kTeacherFloor);
            }
          }
          *reinterpret_cast<uint4*>(gp + u * 256) =
              gc_piece8<TY, DENSE, TAU2, SPARSE>(p, r, row, p.v0 + j, 8, zv[u], ref[u], fy);
        }
      } else {
#pragma unroll 1
        for (int u = 0; u < kGcUnroll; ++u) {
          const int j = j0 + u * 256;
          if (j < p.cols_pad) gc_piece_slow<TY, DENSE, TAU2, SPARSE>(p, r, row, j);
        }
      }
    }
  }
}

// ---- plain stores: dW rows (bf16), dH accumulation (fp32 -> bf16), test hook (fp32) ------------
enum StoreMode { kStoreF32 = 0, kAccumF32 = 1, kFinalBf16 = 2, kStoreBf16 = 3 };
struct StoreParams {
  int mode;
  int m_total, n_total;    // valid extent of the logical C (rows, cols)
  int m_begin;             // rows below this are not written (stage1: old vocabulary)
  float* c32;              // fp32 C / accumulator, row stride ld32
  int64_t ld32;
  __nv_bfloat16* c16;      // bf16 output, row stride ld16
  int64_t ld16;
  int64_t row0_32, row0_16;  // row offsets of tile row 0 inside c32 / c16
  // fp16 gradient operand: the accumulator holds S x the result; non-null pointers = multiply by 1 / S
  const int32_t* scale_n_norm;
  const float* scale_coef;
  float scale_tau;
};

struct StoreEpi {
  using Params = StoreParams;
  static constexpr int kYSlots = 0;
  static constexpr int kGSlots = 0;
  // The dW / dH GEMMs share their SMs with the gradient kernel of the next vocabulary chunk (kd_grad_cached_kernel,
  // HBM-bound): six operand stages (192 KB) and <= 64 registers per thread (launch bound of 1024 threads) leave it
  // 24 K registers and ~30 KB of shared memory per SM.
  static constexpr int kMaxStages = 6;
  static constexpr int kBoundThreads = 1024;
  const Params& p;
  EpiThread t;
  int row;

  float out_scale;
  __device__ StoreEpi(const Params& p_, const EpiThread& t_) : p(p_), t(t_) {
    out_scale = p.scale_n_norm != nullptr ? 1.0f / g_operand_scale(p.scale_n_norm, p.scale_coef, p.scale_tau) : 1.0f;
  }
  __device__ void begin_unit(const Geom&, int m0, int) { row = m0 + t.row_in_tile; }

  __device__ __forceinline__ void store16(const uint32_t (&raw)[16], int col0, int ncols) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) * out_scale;
    if (p.mode == kAccumF32 || p.mode == kFinalBf16) {
      const float* acc = p.c32 + (p.row0_32 + row) * p.ld32 + col0;
      if (ncols >= 16 && (p.ld32 & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 o = *reinterpret_cast<const float4*>(acc + 4 * q);
          v[4 * q] += o.x; v[4 * q + 1] += o.y; v[4 * q + 2] += o.z; v[4 * q + 3] += o.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ncols) v[j] += acc[j];
      }
    }
    if (p.mode == kStoreF32 || p.mode == kAccumF32) {
      float* dst = p.c32 + (p.row0_32 + row) * p.ld32 + col0;
      if (ncols >= 16 && (p.ld32 & 3) == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ncols) dst[j] = v[j];
      }
    } else {
      __nv_bfloat16* dst = p.c16 + (p.row0_16 + row) * p.ld16 + col0;
      if (ncols >= 16 && (p.ld16 & 7) == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float t8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t8[j] = v[8 * q + j];
          Vec8<__nv_bfloat16> pk;
          pk.pack(t8);
          *reinterpret_cast<uint4*>(dst + 8 * q) = pk.a;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
      }
    }
  }

  template <int CG>
  __device__ void tile(const Geom&, int n_blk, uint32_t tmem_acc, uint32_t tempty_bar) {
    const int col_base = n_blk * BN + t.cgrp * 32;
    const bool row_ok = row < p.m_total && row >= p.m_begin;
#pragma unroll 1
    for (int c = 0; c < kSteps; ++c) {
      uint32_t raw0[16], raw1[16];
      __syncwarp();
      const uint32_t taddr = tmem_acc + t.tmem_lane_off + (uint32_t)(c * kStepCols + t.cgrp * 32);
      tmem_ld16(taddr, raw0);
      tmem_ld16(taddr + 16u, raw1);
      tmem_ld_wait();
      if (c == kSteps - 1) {
        fence_before_sync();
        __syncwarp();
        if (t.lane == 0) release_tmem<CG>(tempty_bar);
      }
      const int col0 = col_base + c * kStepCols;
      const int nrem = p.n_total - col0;
      if (!row_ok || nrem <= 0) continue;
      store16(raw0, col0, nrem < 16 ? nrem : 16);
      if (nrem > 16) store16(raw1, col0 + 16, nrem - 16 < 16 ? nrem - 16 : 16);
    }
  }
  __device__ void end_unit(const Geom&) {}
  __device__ void finish() {}
};

// ---------------------------------------------------------------------------------------------
// Work-unit queue.  Static mode: unit u0, u0 + step, ... as a persistent kernel usually does.  Dynamic mode: one
// thread of the (leader) CTA draws unit numbers from a global atomic counter and publishes each to both CTAs of
// the pair through a kSchedSlots-deep shared-memory queue (st.async + mbarrier complete_tx, the TMA publish
// mechanism); every consumer role reads the same sequence and releases the slot on the leader's "empty" barrier.
// The next number is drawn when the operand producer has issued the last load of the current unit: early enough
// to hide the atomic's latency behind the operand ring, late enough that nobody hoards units.
// A CTA (pair) that starts late or shares its SM budget with another kernel (the three backward chains, NCCL)
// simply draws fewer units, so launches overlap without tail effects.  The pair that draws the last number resets
// the counter, which makes the counter reusable by a later launch without a memset.
// ---------------------------------------------------------------------------------------------
// a unit is dead when row compaction left no live row in its M tile (rows_dim 1) or no live k-block (rows_dim 2)
template <int CG>
__device__ __forceinline__ bool unit_live(const Geom& g, int u, int n_rows_live) {
  if (g.rows_dim == 1) return (g.n_fast ? u / num_ranges_of(g) : u % g.num_m_blk) * (CG * BM) < n_rows_live;
  if (g.rows_dim == 2) return n_rows_live > 0;
  return true;
}

struct UnitQueue {
  static_assert(kSchedSlots == 2, "slot / phase are decoded from the low two bits of `it`");
  uint32_t full0;   // first local "full" barrier; the unit slots sit 16 * kSchedSlots + 16 bytes behind it
  uint32_t empty0;  // first "empty" barrier of the leader CTA (shared::cluster address for a pair)
  int it;           // dynamic: number of entries consumed so far; static: the next unit

  // returns the next unit or -1; called by exactly one thread per consumer agent
  template <int CG>
  __device__ __forceinline__ int next(const Geom& g, int n_rows_live) {
    if (!g.dynamic) {
      for (;;) {  // static stride; dead units (row compaction) are stepped over by every consumer alike
        const int u = it;
        it += CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
        if (u >= g.num_units) return -1;
        if (unit_live<CG>(g, u, n_rows_live)) return u;
      }
    }
    const uint32_t slot = (uint32_t)it & 1u, phase = ((uint32_t)it >> 1) & 1u;
    ++it;
    mbar_wait(full0 + 8u * slot, phase);
    const int u = lds_s32(full0 + 16u * kSchedSlots + 16u + 4u * slot);
    if (CG == 2) mbar_arrive_cluster(empty0 + 8u * slot);
    else mbar_arrive(empty0 + 8u * slot);
    return u;
  }
};

// ---------------------------------------------------------------------------------------------
// The persistent warp-specialised GEMM.  CG = 1: one CTA per 128 x 256 tile.  CG = 2: a CTA pair
// (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile - each CTA stages its own 128 rows of A and
// HALF of the B tile, the leader CTA issues one MMA for both tensor cores, and each CTA's epilogue
// drains its own 128 accumulator rows.  The pair halves the L2 -> smem traffic of B and the smem
// bytes per stage (32 KB instead of 48 KB), which buys the deeper operand ring the epilogue rings need.
//   A: K-major  -> global [M rows][K] (K contiguous), one 64(k) x 128(m) box per stage
//      MN-major -> global [K rows][M] (M contiguous), two 64(m) x 64(k) boxes per stage
//   B: K-major  -> global [N rows][K], one 64 x (256/CG) box;  MN-major -> [K rows][N], 64 x 64 boxes
//   tma_y : teacher logits [rows][V] (16-bit), 64-column x 128-row boxes into the y ring (warp 3)
//   tma_g : gradient scratch [rows][v_chunk] bf16, 64 x 128 boxes stored from the g ring (epilogue)
// ---------------------------------------------------------------------------------------------
template <class Epi, bool A_MN, bool B_MN, int CG>
__global__ void __cluster_dims__(CG, 1, 1) __launch_bounds__(Epi::kBoundThreads, 1)
kd_umma_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_y, const __grid_constant__ CUtensorMap tma_g, const Geom g,
               const typename Epi::Params ep, int* __restrict__ sched_counter) {
  using Plan = SmemPlan<CG, Epi::kYSlots, Epi::kGSlots, Epi::kMaxStages>;
  constexpr int kStages = Plan::kStages;
  constexpr int kYSlots = Epi::kYSlots;
  constexpr uint32_t kBBytesL = Plan::kBBytesL;      // this CTA's share of the B tile
  constexpr uint32_t kStageBytesL = kABytes + kBBytesL;
  constexpr int BNL = BN / CG;                       // B rows (n) staged by this CTA
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + kStages * kABytes;
  const uint32_t sY = base + Plan::kYOff;
  const uint32_t sG = base + Plan::kGOff;
  const uint32_t sBar = base + Plan::kBarOff;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return sBar + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return sBar + 8u * (2 * kStages + 2 + s); };
  const uint32_t yfull0 = sBar + 8u * (2 * kStages + 4);
  const uint32_t yempty0 = yfull0 + 8u * (kYSlots > 0 ? kYSlots : 1);
  const uint32_t sched_req = sBar + 8u * Plan::kReqBar;
  const uint32_t sched_full0 = sBar + 8u * Plan::kSchedBar0;
  const uint32_t sched_empty0 = sched_full0 + 8u * kSchedSlots;
  const uint32_t tmem_slot = sBar + 8u * Plan::kNumBars;
  const uint32_t sched_unit0 = tmem_slot + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw_addr));

  const int warp = threadIdx.x >> 5;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int n_rows_live = g.n_rows != nullptr ? *g.n_rows : 0x7fffffff;
  // k-blocks actually run: all of them, or (dW with compacted rows) those that hold live rows
  const int num_k_live = g.rows_dim == 2 && (n_rows_live + BK - 1) / BK < g.num_k_blk ? (n_rows_live + BK - 1) / BK
                                                                                        : g.num_k_blk;
  UnitQueue uq;
  uq.full0 = sched_full0;
  uq.empty0 = CG == 2 ? mapa(sched_empty0, 0) : sched_empty0;
  uq.it = g.dynamic ? 0 : (CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    if (kYSlots > 0) prefetch_tmap(&tma_y);
    if (Epi::kGSlots > 0) prefetch_tmap(&tma_g);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), CG);   // one arrive.expect_tx per CTA of the pair (leader's copy is the live one)
      mbar_init(empty_bar(s), 1);   // one (multicast) MMA commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps * CG);  // epilogue warps of both CTAs release the leader's MMA
    }
    for (int s = 0; s < kYSlots; ++s) {
      mbar_init(yfull0 + 8u * s, 1);
      mbar_init(yempty0 + 8u * s, kEpiWarps);
    }
    mbar_init(sched_req, 1);  // the leader's operand producer, once per unit
    for (int s = 0; s < kSchedSlots; ++s) {
      mbar_init(sched_full0 + 8u * s, 1);  // the scheduler's arrive.expect_tx; the unit number is the 4 tx bytes
      // consumers of both CTAs: operand producer, epilogue warps, teacher-tile producer, + the MMA issuer
      mbar_init(sched_empty0 + 8u * s, CG * (1 + kEpiWarps + (kYSlots > 0 ? 1 : 0)) + 1);
    }
    fence_mbar_init();
  }
  if (CG == 2) cluster_sync_all();  // peer barriers exist before any remote arrive / TMA signal
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_slot, kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================= TMA producer (A / B operand ring) =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = uq.template next<CG>(g, n_rows_live); u >= 0; u = uq.template next<CG>(g, n_rows_live)) {
        int m_blk, range, n_begin, n_end;
        decode_unit(g, u, m_blk, range, n_begin, n_end);
        const int m_row = g.a_m0 + (m_blk * CG + (int)cta_rank) * BM;
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          const int n_row = g.b_n0 + n_blk * BN + (int)cta_rank * BNL;
          for (int kb = 0; kb < num_k_live; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t a_dst = sA + stage * kABytes, b_dst = sB + stage * kBBytesL;
            if (CG == 1) {
              const uint32_t fb = full_bar(stage);
              mbar_expect_tx(fb, kStageBytesL);
              if (!A_MN) {
                tma_load_2d(a_dst, &tma_a, g.a_k0 + kb * BK, m_row, fb);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j)
                  tma_load_2d(a_dst + j * kBoxMnBytes, &tma_a, m_row + 64 * j, g.a_k0 + kb * BK, fb);
              }
              if (!B_MN) {
                tma_load_2d(b_dst, &tma_b, g.b_k0 + kb * BK, n_row, fb);
              } else {
#pragma unroll
                for (int j = 0; j < BNL / 64; ++j)
                  tma_load_2d(b_dst + j * kBoxMnBytes, &tma_b, n_row + 64 * j, g.b_k0 + kb * BK, fb);
              }
            } else {
              const uint32_t fb = mapa(full_bar(stage), 0);  // the leader's full barrier counts both CTAs' bytes
              mbar_expect_tx_cluster(fb, kStageBytesL);
              if (!A_MN) {
                tma_load_2d_cg2(a_dst, &tma_a, g.a_k0 + kb * BK, m_row, fb);
              } else {
#pragma unroll
                for (int j = 0; j < BM / 64; ++j)
                  tma_load_2d_cg2(a_dst + j * kBoxMnBytes, &tma_a, m_row + 64 * j, g.a_k0 + kb * BK, fb);
              }
              if (!B_MN) {
                tma_load_2d_cg2(b_dst, &tma_b, g.b_k0 + kb * BK, n_row, fb);
              } else {
#pragma unroll
                for (int j = 0; j < BNL / 64; ++j)
                  tma_load_2d_cg2(b_dst + j * kBoxMnBytes, &tma_b, n_row + 64 * j, g.b_k0 + kb * BK, fb);
              }
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        // all loads of this unit are in flight: now (not earlier) the scheduler may draw the next unit, so a
        // pair never holds a unit it will not start for a long time while other pairs go idle
        if (g.dynamic && leader) mbar_arrive(sched_req);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of the pair only) =================
    if (lane == 0 && leader) {
      // A / B format field: 1 = bf16, 0 = fp16 (bits 7..9 and 10..12 of the kind::f16 instruction descriptor)
      const uint32_t idesc = instr_desc_bf16(BM * CG, BN, A_MN, B_MN) & (g.ab_fp16 ? ~((1u << 7) | (1u << 10)) : ~0u);
      constexpr uint64_t a_base = A_MN ? smem_desc_base(kBoxMnBytes, 1024) : smem_desc_base(16, 1024);
      constexpr uint64_t b_base = B_MN ? smem_desc_base(kBoxMnBytes, 1024) : smem_desc_base(16, 1024);
      constexpr uint32_t a_kstep = A_MN ? 2048u : 32u;  // bytes per UMMA_K = 16 elements
      constexpr uint32_t b_kstep = B_MN ? 2048u : 32u;
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int u = uq.template next<CG>(g, n_rows_live); u >= 0; u = uq.template next<CG>(g, n_rows_live)) {
        int m_blk, range, n_begin, n_end;
        decode_unit(g, u, m_blk, range, n_begin, n_end);
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          fence_after_sync();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < num_k_live; ++kb) {
            mbar_wait(full_bar(stage), phase);
            fence_after_sync();
            const uint32_t a_src = sA + stage * kABytes, b_src = sB + stage * kBBytesL;
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t da = smem_desc(a_base, a_src + k * a_kstep), db = smem_desc(b_base, b_src + k * b_kstep);
              if (CG == 2) mma_bf16_cg2(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              else mma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs) when these MMAs retire
            if (CG == 2) mma_commit_cg2(empty_bar(stage), 3);
            else mma_commit(empty_bar(stage));
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          // accumulator complete -> epilogue (of both CTAs)
          if (CG == 2) mma_commit_cg2(tfull_bar(acc), 3);
          else mma_commit(tfull_bar(acc));
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================= work-unit scheduler (leader CTA; dynamic mode only) =================
    if (g.dynamic && lane == 0 && leader) {
      const int npairs = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
      int slot = 0;
      uint32_t phase = 0, req_phase = 0;
      for (bool first = true;; first = false) {
        if (!first) {  // the producer has issued the current unit's last load
          mbar_wait(sched_req, req_phase);
          req_phase ^= 1u;
        }
        mbar_wait(sched_empty0 + 8u * slot, phase ^ 1u);  // every consumer of both CTAs has read the old entry
        int u;
        do {  // dead units (row compaction) are drawn and dropped right here
          const int drawn = atomicAdd(sched_counter, 1);
          u = drawn < g.num_units ? drawn : -1;
          if (drawn == g.num_units + npairs - 1) atomicExch(sched_counter, 0);  // last draw of the launch: re-arm
        } while (u >= 0 && !unit_live<CG>(g, u, n_rows_live));
#pragma unroll
        for (int r = 0; r < CG; ++r) {
          const uint32_t fb = mapa(sched_full0 + 8u * slot, (uint32_t)r);
          mbar_expect_tx_cluster(fb, 4);
          st_async_b32(mapa(sched_unit0 + 4u * slot, (uint32_t)r), (uint32_t)u, fb);
        }
        if (u < 0) break;
        if (++slot == kSchedSlots) {
          slot = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 3) {
    // ================= teacher-tile producer (y ring, local to each CTA) =================
    if (kYSlots > 0 && lane == 0) {
      // (an L2 evict-first policy on these loads was tried: same time, but 0.32 GB MORE DRAM reads per forward
      //  launch - 1.98 vs 1.66 GB in ncu - so the teacher tiles use the default policy)
      int slot = 0;
      uint32_t phase = 0;
      for (int u = uq.template next<CG>(g, n_rows_live); u >= 0; u = uq.template next<CG>(g, n_rows_live)) {
        int m_blk, range, n_begin, n_end;
        decode_unit(g, u, m_blk, range, n_begin, n_end);
        const int m_row = (m_blk * CG + (int)cta_rank) * BM;
        for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
          for (int c = 0; c < kSteps; ++c) {
            mbar_wait(yempty0 + 8u * slot, phase ^ 1u);
            const uint32_t fb = yfull0 + 8u * slot;
            mbar_expect_tx(fb, kStepBytes);
#pragma unroll
            for (int b = 0; b < kStepBoxes; ++b)
              tma_load_2d(sY + slot * kStepBytes + b * kBoxBytes, &tma_y,
                          g.b_n0 + n_blk * BN + c * kStepCols + 64 * b, m_row, fb);
            if (++slot == kYSlots) {
              slot = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue (each CTA drains its own 128 accumulator rows) =================
    EpiThread et;
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    et.row_in_tile = q * 32 + lane;
    et.cgrp = (warp - 4) >> 2;
    et.lane = lane;
    et.epi_tid = threadIdx.x - 128;
    et.tmem_lane_off = (uint32_t)(q * 32) << 16;
    et.y_base = sY;
    et.g_base = sG;
    et.yfull0 = yfull0;
    et.yempty0 = yempty0;
    et.tma_g = &tma_g;
    Epi epi(ep, et);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (;;) {
      int u = 0;
      if (lane == 0) u = uq.template next<CG>(g, n_rows_live);  // one agent per warp; static mode is warp-uniform anyway
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u < 0) break;
      int m_blk, range, n_begin, n_end;
      decode_unit(g, u, m_blk, range, n_begin, n_end);
      epi.begin_unit(g, (m_blk * CG + (int)cta_rank) * BM, range);
      for (int n_blk = n_begin; n_blk < n_end; ++n_blk) {
        // one warp polls the accumulator-full barrier; the other 15 wait in a hardware named barrier instead of
        // each spinning on try_wait (the polls were 17 % of all issued instructions of the forward kernel)
        if (warp == 4) mbar_wait(tfull_bar(acc), acc_phase);
        named_bar_sync(3, kEpiThreads);
        fence_after_sync();
        // the policy releases the TMEM buffer itself (arrive on the leader's tempty barrier)
        const uint32_t rel = CG == 2 ? mapa(tempty_bar(acc), 0) : tempty_bar(acc);
        epi.template tile<CG>(g, n_blk, tmem_base + (uint32_t)(acc * BN), rel);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
      epi.end_unit(g);
    }
    epi.finish();
  }

  __syncwarp();
  fence_before_sync();
  if (CG == 2) cluster_sync_all();  // the peer may still read this CTA's smem / signal its barriers
  else __syncthreads();
  if (warp == 2) {
    fence_after_sync();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// Row merge: partial records of every (range, half) -> row_stats + loss sums
// ---------------------------------------------------------------------------------------------
struct MergeParams {
  const float* partials;  // [nrec][R][8]
  int nrec, R, V;
  const int32_t* row_target;
  const void* y;
  int y_dtype;
  int64_t y_stride;
  int dense;
  int sparse;
  const float4* sp_rowc;  // sparse: [R] = (sum_k p_k log p_k, sum of v_k over hits, number of hits, 0)
  float inv_tau;
  float* row_stats;     // [R][4]
  float* block_sums;    // [gridDim.x][8]
  // vocab-parallel: when rank_rec != nullptr the kernel stops after merging this rank's column partials and
  // writes one kRankRecFloats record per row (nothing is finalised; kd_fused_merge_ranks does that)
  float* rank_rec;      // [R][kRankRecFloats]
  int label_off;
};

__global__ void __launch_bounds__(256) kd_fused_merge_kernel(const MergeParams p) {
  __shared__ float sm[8][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ce = 0.f, kl = 0.f, tce = 0.f, nv = 0.f, hits = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < p.R; row += gridDim.x * 8) {
    const int target = p.row_target[row];
    if (target < 0) {
      if (lane == 0 && p.rank_rec == nullptr)
        *reinterpret_cast<float4*>(p.row_stats + (size_t)row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float m = -CUDART_INF_F, s1 = 0.f, st = 0.f, mt = -CUDART_INF_F, t1 = 0.f, tt = 0.f, a = 0.f, zl = 0.f;
    for (int j = lane; j < p.nrec; j += 32) {
      const float* rec = p.partials + ((size_t)j * p.R + row) * kRecFloats;
      const float4 x = *reinterpret_cast<const float4*>(rec);
      const float4 w = *reinterpret_cast<const float4*>(rec + 4);
      merge_student(m, s1, st, x.x, x.y, x.z, p.inv_tau);
      if (p.dense) merge_teacher(mt, t1, tt, a, x.w, w.x, w.y, w.z, p.inv_tau);
      else a += w.z;  // sparse: partial sums of p_k z[i_k]
      zl += w.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), a2 = __shfl_xor_sync(0xffffffffu, s1, o),
                  b2 = __shfl_xor_sync(0xffffffffu, st, o);
      merge_student(m, s1, st, m2, a2, b2, p.inv_tau);
      if (p.dense) {
        const float mt2 = __shfl_xor_sync(0xffffffffu, mt, o), c2 = __shfl_xor_sync(0xffffffffu, t1, o),
                    d2 = __shfl_xor_sync(0xffffffffu, tt, o), e2 = __shfl_xor_sync(0xffffffffu, a, o);
        merge_teacher(mt, t1, tt, a, mt2, c2, d2, e2, p.inv_tau);
      } else {
        a += __shfl_xor_sync(0xffffffffu, a, o);
      }
      zl += __shfl_xor_sync(0xffffffffu, zl, o);
    }
    if (lane == 0 && p.rank_rec != nullptr) {
      const int tl = target - p.label_off;
      float yl = 0.f;
      if (p.dense && tl >= 0 && tl < p.V) {
        const int64_t off = (int64_t)row * p.y_stride + tl;
        if (p.y_dtype == KD_DTYPE_F32) yl = reinterpret_cast<const float*>(p.y)[off];
        else if (p.y_dtype == KD_DTYPE_BF16) yl = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.y)[off]);
        else yl = __half2float(reinterpret_cast<const __half*>(p.y)[off]);
      }
      const float4 rc = p.sparse ? p.sp_rowc[row] : make_float4(0.f, 0.f, 0.f, 0.f);
      float* rec = p.rank_rec + (size_t)row * kRankRecFloats;
      *reinterpret_cast<float4*>(rec) = make_float4(m, s1, st, mt);
      *reinterpret_cast<float4*>(rec + 4) = make_float4(t1, tt, a, zl);
      *reinterpret_cast<float4*>(rec + 8) = make_float4(yl, rc.x, rc.y, rc.z);
    } else if (lane == 0) {
      if (target - p.label_off >= p.V) zl = CUDART_NAN_F;  // label outside the vocabulary: loud, not silently wrong
      const float lse1 = m + ln_acc(s1);
      const float lset = m * p.inv_tau + ln_acc(st);
      float lsett = 0.f;
      ce += lse1 - zl;
      if (p.dense) {
        lsett = mt * p.inv_tau + ln_acc(tt);
        kl += a * p.inv_tau / tt - lsett + lset;
        float yl;
        const int64_t off = (int64_t)row * p.y_stride + target;
        if (p.y_dtype == KD_DTYPE_F32) yl = reinterpret_cast<const float*>(p.y)[off];
        else if (p.y_dtype == KD_DTYPE_BF16) yl = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.y)[off]);
        else yl = __half2float(reinterpret_cast<const __half*>(p.y)[off]);
        tce += (mt + ln_acc(t1)) - yl;
      } else if (p.sparse) {
        // KL_r = sum_k p_k (log p_k - log q_tau[i_k]) with log q_tau[i] = z_i / tau - LSE_tau and sum_k p_k = 1
        const float4 rc = p.sp_rowc[row];
        kl += rc.x - a * p.inv_tau + lset;
        tce += rc.y;   // distillation_loss.py:110-116: teacher log-probs of the hits
        hits += rc.z;
      }
      nv += 1.f;
      *reinterpret_cast<float4*>(p.row_stats + (size_t)row * 4) = make_float4(lse1, lset, lsett, 1.f);
    }
  }
  if (p.rank_rec != nullptr) return;  // block-uniform
  if (lane == 0) {
    sm[warp][0] = ce; sm[warp][1] = kl; sm[warp][2] = tce; sm[warp][3] = nv; sm[warp][4] = hits;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = 0.f;
    if (threadIdx.x < 5)
      for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
    p.block_sums[(size_t)blockIdx.x * kNumPartialSlots + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Vocab-parallel: merge the per-slice row records of all ranks (SURVEY.md 8e "optional mode" = the split-V
// merge rule of appendix C across GPUs) -> row_stats + loss sums.  One thread per row, ranks in fixed order.
// ---------------------------------------------------------------------------------------------
struct RankMergeParams {
  const float* rec;  // [G][R][kRankRecFloats]
  int G, R;
  const int32_t* row_target;
  int dense, sparse;
  float inv_tau;
  float* row_stats;
  float* block_sums;
};

__global__ void __launch_bounds__(256) kd_fused_rank_merge_kernel(const RankMergeParams p) {
  __shared__ float sm[8][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ce = 0.f, kl = 0.f, tce = 0.f, nv = 0.f, hits = 0.f;
  for (int row = blockIdx.x * 256 + threadIdx.x; row < p.R; row += gridDim.x * 256) {
    if (p.row_target[row] < 0) {
      *reinterpret_cast<float4*>(p.row_stats + (size_t)row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    float m = -CUDART_INF_F, s1 = 0.f, st = 0.f, mt = -CUDART_INF_F, t1 = 0.f, tt = 0.f, a = 0.f, zl = 0.f, yl = 0.f;
    float4 w8 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < p.G; ++g) {
      const float* rec = p.rec + ((size_t)g * p.R + row) * kRankRecFloats;
      const float4 x = *reinterpret_cast<const float4*>(rec);
      const float4 w = *reinterpret_cast<const float4*>(rec + 4);
      w8 = *reinterpret_cast<const float4*>(rec + 8);  // sparse row constants are identical on every rank
      merge_student(m, s1, st, x.x, x.y, x.z, p.inv_tau);
      if (p.dense) merge_teacher(mt, t1, tt, a, x.w, w.x, w.y, w.z, p.inv_tau);
      else a += w.z;
      zl += w.w;
      yl += w8.x;
    }
    const float lse1 = m + ln_acc(s1);
    const float lset = m * p.inv_tau + ln_acc(st);
    float lsett = 0.f;
    ce += lse1 - zl;
    if (p.dense) {
      lsett = mt * p.inv_tau + ln_acc(tt);
      kl += a * p.inv_tau / tt - lsett + lset;
      tce += (mt + ln_acc(t1)) - yl;
    } else if (p.sparse) {
      kl += w8.y - a * p.inv_tau + lset;
      tce += w8.z;
      hits += w8.w;
    }
    nv += 1.f;
    *reinterpret_cast<float4*>(p.row_stats + (size_t)row * 4) = make_float4(lse1, lset, lsett, 1.f);
  }
  ce = warp_sum(ce); kl = warp_sum(kl); tce = warp_sum(tce); nv = warp_sum(nv); hits = warp_sum(hits);
  if (lane == 0) {
    sm[warp][0] = ce; sm[warp][1] = kl; sm[warp][2] = tce; sm[warp][3] = nv; sm[warp][4] = hits;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = 0.f;
    if (threadIdx.x < 5)
      for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
    p.block_sums[(size_t)blockIdx.x * kNumPartialSlots + threadIdx.x] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Sparse teacher preparation (distillation_loss.py:77-95, 108-116): one CTA per valid row.
//   p_k = softmax(v / tau) over the K entries, sum_k p_k log p_k, label hits; entries sorted by vocabulary
//   index (bitonic sort in shared memory) and the per-256-column-tile offsets the epilogues index with.
// ---------------------------------------------------------------------------------------------
constexpr int kPrepThreads = 128;
constexpr int kMaxFusedTopK = 1024;

struct SparsePrepParams {
  const float* v;      // [R][K] teacher log-probs at tau = 1
  const int32_t* idx;  // [R][K]
  const int32_t* row_target;
  int R, K, Kp2, V, n_off;  // n_off = number of tile offsets per row (tiles + 1)
  int idx_off;              // vocab-parallel: entries are kept when idx - idx_off falls in [0, V)
  float inv_tau;
  int32_t* s_idx;
  float* s_p;
  float4* rowc;
  uint16_t* off;  // [R][off_stride]
  int off_stride;
};

__global__ void __launch_bounds__(kPrepThreads) kd_sparse_prepare_kernel(const SparsePrepParams p) {
  extern __shared__ uint8_t prep_smem[];
  int32_t* keys = reinterpret_cast<int32_t*>(prep_smem);
  float* vals = reinterpret_cast<float*>(prep_smem + (size_t)p.Kp2 * 4);
  __shared__ float red[kPrepThreads / 32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int row = blockIdx.x; row < p.R; row += gridDim.x) {
    const int target = p.row_target[row];
    if (target < 0) continue;  // block-uniform
    const float* vrow = p.v + (size_t)row * p.K;
    const int32_t* irow = p.idx + (size_t)row * p.K;
    // max and log-sum-exp of v / tau
    float vm = -CUDART_INF_F;
    for (int k = tid; k < p.K; k += kPrepThreads) vm = fmaxf(vm, vrow[k]);
    vm = warp_max(vm);
    if (lane == 0) red[warp][0] = vm;
    __syncthreads();
    vm = fmaxf(fmaxf(red[0][0], red[1][0]), fmaxf(red[2][0], red[3][0]));
    __syncthreads();
    float sum = 0.f;
    for (int k = tid; k < p.K; k += kPrepThreads) sum += __expf((vrow[k] - vm) * p.inv_tau);
    sum = warp_sum(sum);
    if (lane == 0) red[warp][0] = sum;
    __syncthreads();
    sum = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
    __syncthreads();
    const float lk = vm * p.inv_tau + ln_acc(sum);
    float plogp = 0.f, hit_sum = 0.f, hits = 0.f;
    for (int k = tid; k < p.Kp2; k += kPrepThreads) {
      int key = 0x7fffffff;
      float pk = 0.f;
      if (k < p.K) {
        const float lp = vrow[k] * p.inv_tau - lk;  // log p_k
        pk = __expf(lp);
        if (pk > 0.f) plogp = fmaf(pk, lp, plogp);
        const int i = irow[k];
        if (i == target) {
          hits += 1.f;
          hit_sum += vrow[k];
        }
        const int il = i - p.idx_off;
        if (i >= 0 && il >= 0 && il < p.V) key = il;  // out-of-range entries sort to the end, never visited
      }
      keys[k] = key;
      vals[k] = pk;
    }
    plogp = warp_sum(plogp);
    hit_sum = warp_sum(hit_sum);
    hits = warp_sum(hits);
    if (lane == 0) {
      red[warp][0] = plogp;
      red[warp][1] = hit_sum;
      red[warp][2] = hits;
    }
    __syncthreads();
    if (tid == 0) {
      p.rowc[row] = make_float4((red[0][0] + red[1][0]) + (red[2][0] + red[3][0]),
                                (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]),
                                (red[0][2] + red[1][2]) + (red[2][2] + red[3][2]), 0.f);
    }
    // bitonic sort by key, ascending
    for (int k2 = 2; k2 <= p.Kp2; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < p.Kp2; i += kPrepThreads) {
          const int l = i ^ j;
          if (l > i) {
            const bool up = (i & k2) == 0;
            const int32_t ki = keys[i], kl = keys[l];
            if ((ki > kl) == up) {
              keys[i] = kl;
              keys[l] = ki;
              const float t = vals[i];
              vals[i] = vals[l];
              vals[l] = t;
            }
          }
        }
        __syncthreads();
      }
    }
    for (int k = tid; k < p.K; k += kPrepThreads) {
      p.s_idx[(size_t)row * p.K + k] = keys[k];
      p.s_p[(size_t)row * p.K + k] = vals[k];
    }
    // off[t] = number of entries with index < 256 t (lower bound in the sorted keys)
    for (int t = tid; t < p.n_off; t += kPrepThreads) {
      const int bound = t * BN;
      int lo = 0, hi = p.K;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < bound) lo = mid + 1;
        else hi = mid;
      }
      p.off[(size_t)row * p.off_stride + t] = (uint16_t)lo;
    }
    __syncthreads();  // keys / vals are rewritten by the next row
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || sym == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// bf16 matrix [outer rows][inner cols] (inner contiguous), box = box_inner x box_outer; 128-byte swizzle for the
// 64-element (128-byte) boxes of the operand / tile rings, 64-byte swizzle for the 32-element boxes of the per-warp
// logit-cache stores
static int make_tmap(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                     uint32_t box_outer, const char* what, uint32_t box_inner = 64) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * 2) % 16 != 0) {
    set_error("%s: TMA needs a 16-byte aligned base and a row stride that is a multiple of 8 elements", what);
    return 1;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu)", what, (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_elems);
    return 1;
  }
  return 0;
}

// SMs the persistent GEMM kernels may occupy in this call (0 = all); a data-parallel caller that overlaps the
// dW all-reduce leaves NCCL's CTAs their own SMs, otherwise the last CTAs of every launch queue behind them
static thread_local int tl_sm_limit = 0;

static int sm_count() {  // of the current device (the entry points switch to the tensors' device first)
  static int n[kMaxDevices] = {};
  const int dev = current_device_slot();
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

// CTA-pair mode (tcgen05 cta_group::2) is the default; KD_UMMA_CTA_GROUP=1 selects the single-CTA kernels
static int cta_group() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("KD_UMMA_CTA_GROUP");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

// dynamic work-unit scheduling (default) needs one zero-initialised counter per in-flight launch: a per-device
// pool handed out round-robin; each launch re-arms its counter when it draws its last number (see UnitQueue)
static bool sched_dynamic() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KD_SCHED");
    v = (e && (e[0] == 's' || e[0] == '0')) ? 0 : 1;  // KD_SCHED=static keeps the fixed stride
  }
  return v != 0;
}

static int* next_sched_counter() {
  constexpr int kMaxDev = 64, kPool = 1024;
  static int* pools[kMaxDev] = {};
  static unsigned seq[kMaxDev] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    int* p = nullptr;
    if (cudaMalloc(&p, kPool * sizeof(int)) != cudaSuccess) return nullptr;
    if (cudaMemset(p, 0, kPool * sizeof(int)) != cudaSuccess) return nullptr;
    pools[dev] = p;
  }
  return pools[dev] + (seq[dev]++ % kPool);
}

template <class Epi, bool A_MN, bool B_MN, int CG>
static int launch_umma_cg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty, const CUtensorMap& tg,
                          const Geom& g, const typename Epi::Params& ep, cudaStream_t stream) {
  auto kern = kd_umma_kernel<Epi, A_MN, B_MN, CG>;
  constexpr uint32_t smem = SmemPlan<CG, Epi::kYSlots, Epi::kGSlots, Epi::kMaxStages>::kBytes;
  static bool attr_set[kMaxDevices] = {};  // per instantiation and device (function attributes are per device)
  const int dev_slot = current_device_slot();
  if (!attr_set[dev_slot]) {
    if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "kd_umma smem attribute"))
      return 1;
    attr_set[dev_slot] = true;
  }
  if (g.num_units <= 0 || g.num_k_blk <= 0) {
    set_error("kd_umma: empty problem (units=%d, k blocks=%d)", g.num_units, g.num_k_blk);
    return 1;
  }
  int sms = sm_count();
  if (tl_sm_limit > 0 && tl_sm_limit < sms) sms = tl_sm_limit;
  const int slots = sms / CG >= 1 ? sms / CG : 1;  // persistent: one CTA (pair) per SM (pair)
  const int grid = (g.num_units < slots ? g.num_units : slots) * CG;
  Geom gg = g;
  int* counter = nullptr;
  gg.dynamic = 0;
  if (sched_dynamic()) {
    counter = next_sched_counter();
    if (!counter) {
      set_error("kd_umma: could not allocate the work-unit counters");
      return 1;
    }
    gg.dynamic = 1;
  }
  kern<<<grid, kThreads, smem, stream>>>(ta, tb, ty, tg, gg, ep, counter);
  return check_launch("kd_umma launch");
}
template <class Epi, bool A_MN, bool B_MN>
static int launch_umma(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& ty, const CUtensorMap& tg,
                       const Geom& g, const typename Epi::Params& ep, cudaStream_t stream) {
  if (cta_group() == 2) return launch_umma_cg<Epi, A_MN, B_MN, 2>(ta, tb, ty, tg, g, ep, stream);
  return launch_umma_cg<Epi, A_MN, B_MN, 1>(ta, tb, ty, tg, g, ep, stream);
}
template <class Epi, bool A_MN, bool B_MN>
static int launch_umma(const CUtensorMap& ta, const CUtensorMap& tb, const Geom& g, const typename Epi::Params& ep,
                       cudaStream_t stream) {
  return launch_umma<Epi, A_MN, B_MN>(ta, tb, ta, ta, g, ep, stream);  // no y / g rings: maps unused
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int tile_m() { return BM * cta_group(); }     // rows of C per work tile
static inline int b_box_rows() { return BN / cta_group(); }  // rows of a K-major B box (per CTA)

// ---- fp16 gradient operand: switch and operand copies -----------------------------------------------
// KD_G_FP16=0 keeps the bf16 gradient operand (and the bf16 h / W operands) for dW and dH
static bool g_fp16_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KD_G_FP16");
    v = e ? (e[0] != '0' ? 1 : 0) : KD_G_FP16_DEFAULT;
  }
  return v != 0;
}

// dst[r, :] (fp16, contiguous rows of `cols`) = src[r, :] (bf16, row stride src_stride); cols % 8 == 0.
// bf16 -> fp16 is exact inside fp16's normal range.  Magnitudes below 6e-8 flush to zero (harmless: far below the
// rounding of the products they enter); magnitudes above 65504 - not seen in LM-head activations or weights, but
// nothing guarantees it - become +-inf on purpose: the gradients then come out non-finite, which a training loop
// notices, instead of being silently wrong as a saturating cast would leave them (KD_G_FP16=0 is the way out).
__global__ void __launch_bounds__(256) kd_cast_bf16_f16_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_stride,
                                                              __half* __restrict__ dst, int rows, int cols) {
  const int vec_per_row = cols >> 3;
  const int64_t total = (int64_t)rows * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec_per_row), c = (int)(i - (int64_t)r * vec_per_row) * 8;
    Vec8<__nv_bfloat16> v;
    v.load_global(src + (int64_t)r * src_stride + c);
    float f[8];
    v.unpack(f);
    Vec8<__half> o;
    o.pack(f);
    *reinterpret_cast<uint4*>(dst + (int64_t)r * cols + c) = o.a;
  }
}

static int cast_bf16_f16(const void* src, int64_t src_stride, void* dst, int rows, int cols, cudaStream_t s) {
  if (rows <= 0) return 0;
  const int64_t total = (int64_t)rows * (cols >> 3);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  kd_cast_bf16_f16_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(src), src_stride,
                                                 reinterpret_cast<__half*>(dst), rows, cols);
  return check_launch("kd_cast_bf16_f16 launch");
}

// ---- workspace layout -----------------------------------------------------------------------------
constexpr int kFwdTilesPerRange = 8;  // A/B in bench.py: 8 is 1 % faster than 4 (fewer partial records), 2 is 2.5 % slower
// 74 x 256 columns: per chunk the gradient GEMM has 16 x 74 tiles (16 per CTA pair) and dW 74 x 4 (4 per pair) -
// whole waves of the 74 pairs - and the backward is 27 launches instead of 51.  Round-robin A/B at configs[1] size
// (tools/vchunk_ab.py): 4.52 ms/step vs 5.10 with 9472 in the power-capped regime, 4.65 vs 4.83 in bench.py.
constexpr int kDefaultVChunk = 18944;
constexpr int kMergeBlocksMax = 1024;

static int norm_v_chunk(int v_chunk, int V) {
  if (v_chunk <= 0) {
    static int env_chunk = -1;  // KD_V_CHUNK overrides the library default (experiments)
    if (env_chunk < 0) {
      const char* e = getenv("KD_V_CHUNK");
      env_chunk = e ? atoi(e) : 0;
    }
    v_chunk = env_chunk > 0 ? env_chunk : kDefaultVChunk;
  }
  v_chunk = cdiv(v_chunk, BN) * BN;
  const int vmax = cdiv(V, BN) * BN;
  return v_chunk < vmax ? v_chunk : vmax;
}

struct Workspace {
  size_t partials_off, partials_bytes;  // forward records
  size_t bsums_off, bsums_bytes;        // merge block sums (+1 reduced record)
  size_t g_off, g_bytes;                // backward gradient chunks, 2 x bf16 [R][v_chunk] (double buffered)
  size_t g_buf_bytes;                   // one of the two
  size_t dh_off, dh_bytes;              // backward dH accumulator, fp32 [R][H]
  size_t h16_off, h16_bytes;            // fp16 gradient operand: fp16 copy of h [R][H] ...
  size_t w16_off, w16_buf_bytes;        // ... and of the current W chunk [v_chunk][H], double buffered
  // sparse teacher view, rebuilt by each call (after the forward region resp. the backward region)
  size_t sp_bytes, sp_idx_rel, sp_p_rel, sp_rowc_rel, sp_off_rel;  // offsets relative to the region start
  int sp_off_stride, sp_n_off;
  size_t fwd_bytes, bwd_bytes;          // end of the forward / backward regions = start of the sparse region
  size_t total;
};

static Workspace plan_workspace(int R, int H, int V, int v_chunk, int K) {
  Workspace w;
  const int num_ranges = cdiv(cdiv(V, BN), kFwdTilesPerRange);
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  w.partials_off = 0;
  w.partials_bytes = up((size_t)num_ranges * kColGroups * R * kRecFloats * sizeof(float));
  w.bsums_off = w.partials_off + w.partials_bytes;
  w.bsums_bytes = up((size_t)(kMergeBlocksMax + 1) * kNumPartialSlots * sizeof(float));
  // the backward reuses the same region from offset 0
  const int vc = norm_v_chunk(v_chunk, V);
  w.g_off = 0;
  w.g_buf_bytes = up((size_t)R * vc * 2);
  w.g_bytes = 2 * w.g_buf_bytes;
  w.dh_off = w.g_off + w.g_bytes;
  w.dh_bytes = up((size_t)R * H * sizeof(float));
  w.h16_off = w.dh_off + w.dh_bytes;
  w.h16_bytes = g_fp16_enabled() ? up((size_t)R * H * 2) : 0;
  w.w16_off = w.h16_off + w.h16_bytes;
  w.w16_buf_bytes = g_fp16_enabled() ? up((size_t)vc * H * 2) : 0;
  w.fwd_bytes = w.bsums_off + w.bsums_bytes;
  w.bwd_bytes = w.w16_off + 2 * w.w16_buf_bytes;
  w.sp_bytes = 0;
  w.sp_idx_rel = w.sp_p_rel = w.sp_rowc_rel = w.sp_off_rel = 0;
  w.sp_n_off = cdiv(V, BN) + 1;
  w.sp_off_stride = (w.sp_n_off + 7) & ~7;
  if (K > 0) {
    w.sp_idx_rel = 0;
    w.sp_p_rel = w.sp_idx_rel + up((size_t)R * K * 4);
    w.sp_rowc_rel = w.sp_p_rel + up((size_t)R * K * 4);
    w.sp_off_rel = w.sp_rowc_rel + up((size_t)R * 16);
    w.sp_bytes = w.sp_off_rel + up((size_t)R * w.sp_off_stride * 2);
  }
  const size_t fwd = w.fwd_bytes + w.sp_bytes, bwd = w.bwd_bytes + w.sp_bytes;
  w.total = fwd > bwd ? fwd : bwd;
  return w;
}

static int check_common(const void* h, int64_t h_stride, const void* W, int64_t w_stride, int R, int H, int V,
                        float tau, const char* who) {
  if (!h || !W || R <= 0 || H <= 0 || V <= 0) {
    set_error("%s: null pointer or empty shape (R=%d H=%d V=%d)", who, R, H, V);
    return 1;
  }
  if (H % 8 != 0 || h_stride % 8 != 0 || w_stride % 8 != 0) {
    set_error("%s: hidden size and row strides must be multiples of 8 (TMA 16-byte rule); H=%d", who, H);
    return 1;
  }
  if (!(tau > 0.f)) {
    set_error("%s: temperature must be > 0", who);
    return 1;
  }
  return 0;
}

template <bool DENSE, typename TY>
static int launch_fwd(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* ty, const CUtensorMap& tz,
                      const Geom& g, const FwdParams& fp, bool tau2, cudaStream_t s) {
  // tz: the logit cache as a TMA store target (any valid map when the cache is off)
  if constexpr (DENSE && sizeof(TY) == 2) {
    if (ty != nullptr) {  // teacher tile staged by TMA
      if (tau2) return launch_umma<FwdEpi<TY, true, true, true>, false, false>(ta, tb, *ty, tz, g, fp, s);
      return launch_umma<FwdEpi<TY, true, false, true>, false, false>(ta, tb, *ty, tz, g, fp, s);
    }
  }
  if (tau2) return launch_umma<FwdEpi<TY, DENSE, true, false>, false, false>(ta, tb, ta, tz, g, fp, s);
  return launch_umma<FwdEpi<TY, DENSE, false, false>, false, false>(ta, tb, ta, tz, g, fp, s);
}
template <bool DENSE, typename TY>
static int launch_grad(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* ty, const CUtensorMap& tg,
                       const Geom& g, const GradParams& gp, bool tau2, cudaStream_t s) {
  if constexpr (DENSE && sizeof(TY) == 2) {
    if (ty != nullptr) {
      if (tau2) return launch_umma<GradEpi<TY, true, true, true>, false, false>(ta, tb, *ty, tg, g, gp, s);
      return launch_umma<GradEpi<TY, true, false, true>, false, false>(ta, tb, *ty, tg, g, gp, s);
    }
  }
  if (tau2) return launch_umma<GradEpi<TY, DENSE, true, false>, false, false>(ta, tb, ta, tg, g, gp, s);
  return launch_umma<GradEpi<TY, DENSE, false, false>, false, false>(ta, tb, ta, tg, g, gp, s);
}

// ---- backward trace (measurement hook): per-kernel start / end times on the three backward streams ----------
// kd_fused_bwd_trace_begin() arms it; every launch of the following backward calls is bracketed by two timed events
// on the stream it is launched on; kd_fused_bwd_trace_read() synchronises and returns (class, chunk, start ms,
// end ms) records relative to the first event.  bench.py uses it for the per-kernel rooflines of the step - the
// live, concurrent timeline, which a serialising profiler cannot show.  Off (the default) it costs one branch.
enum TraceClass { kTraceCast = 0, kTraceGrad = 1, kTraceDw = 2, kTraceDh = 3, kTraceGradRecompute = 4 };
struct TraceRec {
  int cls, chunk;
  cudaEvent_t e0, e1;
};
struct BwdTrace {
  std::mutex mu;
  bool on = false;
  std::vector<TraceRec> recs;
  std::vector<cudaEvent_t> pool;  // recycled timed events
  cudaEvent_t take() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
static BwdTrace g_trace;

struct TraceScope {  // records e0 now and e1 at scope exit on `s`; inert when the trace is off
  cudaStream_t s;
  cudaEvent_t e1 = nullptr;
  TraceScope(int cls, int chunk, cudaStream_t s_) : s(s_) {
    if (!g_trace.on) return;
    std::lock_guard<std::mutex> lock(g_trace.mu);
    TraceRec r = {cls, chunk, g_trace.take(), g_trace.take()};
    cudaEventRecord(r.e0, s);
    e1 = r.e1;
    g_trace.recs.push_back(r);
  }
  ~TraceScope() {
    if (e1) cudaEventRecord(e1, s);
  }
};

// ---- backward pipeline: three serial chains on three internal streams (forked from / joined into the caller's) -------
//   stream G        : grad(0) grad(1) grad(2) ...          (G chunk c -> buffer c & 1; + the fp16 operand copies)
//   stream W        :         dW(0)   dW(1) ...            (waits grad(c))
//   stream H        :         dH(0)   dH(1) ...            (waits grad(c); serial: fp32 accumulation order is fixed)
// and grad(c + 2) waits for dW(c) and dH(c) before it overwrites their buffer.  Every kernel is persistent with
// one CTA per SM, so a kernel of the next chain fills the SMs the previous one leaves idle in its tail (and the
// 10 CTA pairs a 64-tile dH launch never uses); results are bit-identical to the serial order.
struct BwdPipe {
  std::mutex enqueue;  // the streams and events are per device: one backward is enqueued at a time
  cudaStream_t sg = nullptr, sw = nullptr, sh = nullptr;
  cudaEvent_t eg[2] = {nullptr, nullptr}, ew[2] = {nullptr, nullptr}, eh[2] = {nullptr, nullptr};
  cudaEvent_t efork = nullptr, ejoin = nullptr;
  // state that survives from one vocabulary-range call of a backward to the next (reset at KD_RANGE_FIRST): which
  // G buffers still have readers in flight, and the last dW / dH events (the dH chain is joined only at the end)
  bool rec_w[2] = {false, false}, rec_h[2] = {false, false};
  bool any_w = false, any_h = false;
  int last_w = 0, last_h = 0;
  const void* owner = nullptr;  // workspace of the backward whose ranges the state above describes
  bool ready = false;
};

static bool bwd_pipe_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KD_BWD_STREAMS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static BwdPipe* get_bwd_pipe() {
  constexpr int kMaxDev = 64;
  static BwdPipe pipes[kMaxDev];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  BwdPipe& p = pipes[dev];
  if (p.ready) return &p;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  // Stream priorities decide which pending CTAs the block scheduler places first.  Measured at configs[1] size in a
  // settled loop (tools/k1_ab.py, profiles/r02c_bwd_prio.log): all three chains at the same (default) priority
  // 4.20 ms/step; gradient chain high + GEMM chains low 4.34 (the GEMMs starve); GEMM chains high (round 1) 4.26 (the
  // gradient kernel's small CTAs, which fit BESIDE a GEMM CTA, are not placed while a GEMM kernel still has CTAs
  // pending and the kernel runs after the GEMMs instead of beside them).  KD_BWD_PRIO=g / w select the other two.
  int prio_g = lo, prio_mm = lo;
  {
    const char* e = getenv("KD_BWD_PRIO");
    if (e && e[0] == 'g') prio_g = hi;
    if (e && e[0] == 'w') prio_mm = hi;
  }
  if (cudaStreamCreateWithPriority(&p.sg, cudaStreamNonBlocking, prio_g) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&p.sw, cudaStreamNonBlocking, prio_mm) != cudaSuccess) return nullptr;
  if (cudaStreamCreateWithPriority(&p.sh, cudaStreamNonBlocking, prio_mm) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&p.efork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&p.ejoin, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreateWithFlags(&p.eg[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&p.ew[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&p.eh[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  p.ready = true;
  return &p;
}

// sparse teacher: validate, run the preparation kernel into `region`, return the view for the epilogues
static int prepare_sparse(const float* topk_v, const int32_t* topk_i, int K, const int32_t* row_target, int R, int V,
                          int v_offset, float tau, const Workspace& ws, uint8_t* region, SparseView* view, const float4** rowc,
                          cudaStream_t s, const char* who) {
  if (!topk_v || !topk_i || K <= 0 || K > kMaxFusedTopK) {
    set_error("%s: sparse teacher needs topk_v, topk_i and 1 <= K <= %d (K=%d)", who, kMaxFusedTopK, K);
    return 1;
  }
  int Kp2 = 1;
  while (Kp2 < K) Kp2 <<= 1;
  SparsePrepParams pp = {};
  pp.v = topk_v;
  pp.idx = topk_i;
  pp.row_target = row_target;
  pp.R = R;
  pp.K = K;
  pp.Kp2 = Kp2;
  pp.V = V;
  pp.idx_off = v_offset;
  pp.n_off = ws.sp_n_off;
  pp.inv_tau = 1.0f / tau;
  pp.s_idx = reinterpret_cast<int32_t*>(region + ws.sp_idx_rel);
  pp.s_p = reinterpret_cast<float*>(region + ws.sp_p_rel);
  pp.rowc = reinterpret_cast<float4*>(region + ws.sp_rowc_rel);
  pp.off = reinterpret_cast<uint16_t*>(region + ws.sp_off_rel);
  pp.off_stride = ws.sp_off_stride;
  const int blocks = R < 8 * sm_count() ? R : 8 * sm_count();
  kd_sparse_prepare_kernel<<<blocks, kPrepThreads, (size_t)Kp2 * 8, s>>>(pp);
  if (check_launch("kd_sparse_prepare launch")) return 1;
  view->idx = pp.s_idx;
  view->p = pp.s_p;
  view->off = pp.off;
  view->K = K;
  view->off_stride = ws.sp_off_stride;
  if (rowc) *rowc = pp.rowc;
  return 0;
}

// teacher logits as a TMA source: 16-bit, 16-byte aligned base and row stride; else the direct-load path
static bool make_teacher_tmap(CUtensorMap* m, const void* y, int y_dtype, int64_t y_stride, int R, int V) {
  if (!y || y_dtype != KD_DTYPE_BF16) return false;
  if ((reinterpret_cast<uintptr_t>(y) & 15) != 0 || (y_stride * 2) % 16 != 0) return false;
  return make_tmap(m, y, (uint64_t)V, (uint64_t)R, (uint64_t)y_stride, BM, "teacher logits") == 0;
}


// ---- logit cache: layout and the cached gradient launch -------------------------------------------------
static inline size_t zc_tile_bytes(int R) { return (size_t)R * BN * 2 + (size_t)(BN / 32) * R * 2; }  // z16 + ref

// the cache a caller-provided buffer of `bytes` holds for R rows of a V-column head (tiles = 0: none)
static LogitCache plan_logit_cache(void* buf, size_t bytes, int R, int V) {
  LogitCache zc = {};
  if (!buf || bytes == 0 || (reinterpret_cast<uintptr_t>(buf) & 255) != 0) return zc;
  size_t tiles = bytes / zc_tile_bytes(R);
  const size_t all = (size_t)cdiv(V, BN);
  if (tiles > all) tiles = all;
  zc.tiles = (int)tiles;
  zc.z16 = reinterpret_cast<__half*>(buf);
  zc.z_stride = (int64_t)tiles * BN;
  zc.ref = reinterpret_cast<int16_t*>(reinterpret_cast<uint8_t*>(buf) + (size_t)R * tiles * BN * 2);
  return zc;
}

static bool dw_n_fast() {  // KD_DW_ORDER=m restores round 1's unit order of the dW GEMM (A/B experiments)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KD_DW_ORDER");
    v = (e && e[0] == 'm') ? 0 : 1;
  }
  return v != 0;
}

static bool grad_cached_overlap() {  // KD_GRAD_OVERLAP=0: the cached gradient kernel stays on the caller's stream order
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KD_GRAD_OVERLAP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <typename TY, bool DENSE, bool SPARSE>
static int launch_grad_cached_t(const GradCachedParams& gp, bool tau2, int blocks, cudaStream_t s) {
  if (tau2) kd_grad_cached_kernel<TY, DENSE, true, SPARSE><<<blocks, kGcThreads, 0, s>>>(gp);
  else kd_grad_cached_kernel<TY, DENSE, false, SPARSE><<<blocks, kGcThreads, 0, s>>>(gp);
  return check_launch("kd_grad_cached launch");
}

static int launch_grad_cached(const GradCachedParams& gp, int teacher_kind, int y_dtype, bool tau2, cudaStream_t s) {
  // one warp per row strip; two CTAs per SM when the kernel has the SMs to itself, one beside a GEMM CTA
  int blocks = cdiv(gp.R, kGcThreads / 32);
  const int cap = 2 * sm_count();
  if (blocks > cap) blocks = cap;
  if (teacher_kind == KD_TEACHER_DENSE) {
    if (y_dtype == KD_DTYPE_BF16) return launch_grad_cached_t<__nv_bfloat16, true, false>(gp, tau2, blocks, s);
    return launch_grad_cached_t<float, true, false>(gp, tau2, blocks, s);
  }
  if (teacher_kind == KD_TEACHER_SPARSE) return launch_grad_cached_t<__nv_bfloat16, false, true>(gp, tau2, blocks, s);
  return launch_grad_cached_t<__nv_bfloat16, false, false>(gp, tau2, blocks, s);
}

}  // namespace fused
}  // namespace kd

using namespace kd;
using namespace kd::fused;

extern "C" size_t kd_fused_workspace_bytes(int R, int H, int V, int v_chunk, int K) {
  if (R <= 0 || H <= 0 || V <= 0 || K < 0) return 0;
  return plan_workspace(R, H, V, v_chunk, K).total;
}

extern "C" size_t kd_fused_logit_cache_bytes(int R, int V, int v_chunk, size_t budget_bytes) {
  if (R <= 0 || V <= 0) return 0;
  // whole backward chunks only (a chunk is either read from the cache or recomputed), the last one may be ragged
  const size_t per_tile = zc_tile_bytes(R);
  const size_t all = (size_t)cdiv(V, BN), chunk_tiles = (size_t)norm_v_chunk(v_chunk, V) / BN;
  size_t tiles = budget_bytes / per_tile;
  if (tiles >= all) tiles = all;
  else tiles = tiles / chunk_tiles * chunk_tiles;
  return tiles * per_tile;
}

static int fused_fwd_impl(const void* h, int64_t h_stride, const void* W, int64_t w_stride, int teacher_kind,
                          const void* y, int y_dtype, int64_t y_stride, const float* topk_v, const int32_t* topk_i,
                          int K, const int32_t* row_target, const int32_t* n_rows, int R, int H, int V, int v_offset,
                          float tau, float* sums, float* row_stats, float* rank_rec, void* logit_cache,
                          size_t logit_cache_bytes, void* workspace, size_t workspace_bytes, void* stream);

extern "C" int kd_fused_linear_fwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride, int teacher_kind,
                                   const void* y, int y_dtype, int64_t y_stride, const float* topk_v,
                                   const int32_t* topk_i, int K, const int32_t* row_target, const int32_t* n_rows,
                                   int R, int H, int V, float tau, float alpha, float* sums, float* row_stats,
                                   void* logit_cache, size_t logit_cache_bytes, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  (void)alpha;
  if (!sums || !row_stats) {
    set_error("kd_fused_linear_fwd: null pointer argument");
    return 1;
  }
  return fused_fwd_impl(h, h_stride, W, w_stride, teacher_kind, y, y_dtype, y_stride, topk_v, topk_i, K, row_target,
                        n_rows, R, H, V, 0, tau, sums, row_stats, nullptr, logit_cache, logit_cache_bytes, workspace,
                        workspace_bytes, stream);
}

extern "C" int kd_fused_linear_fwd_partial(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                                           int teacher_kind, const void* y, int y_dtype, int64_t y_stride,
                                           const float* topk_v, const int32_t* topk_i, int K,
                                           const int32_t* row_target, const int32_t* n_rows, int R, int H, int V,
                                           int v_offset, float tau, float* rank_rec, void* logit_cache,
                                           size_t logit_cache_bytes, void* workspace, size_t workspace_bytes,
                                           void* stream) {
  if (!rank_rec || v_offset < 0) {
    set_error("kd_fused_linear_fwd_partial: null record buffer or negative vocabulary offset");
    return 1;
  }
  return fused_fwd_impl(h, h_stride, W, w_stride, teacher_kind, y, y_dtype, y_stride, topk_v, topk_i, K, row_target,
                        n_rows, R, H, V, v_offset, tau, nullptr, nullptr, rank_rec, logit_cache, logit_cache_bytes,
                        workspace, workspace_bytes, stream);
}

extern "C" size_t kd_fused_merge_workspace_bytes(void) {
  return (size_t)(kMergeBlocksMax + 1) * kNumPartialSlots * sizeof(float);
}

extern "C" int kd_fused_merge_ranks(const float* rank_recs, int G, const int32_t* row_target, int R, int teacher_kind,
                                    float tau, float* sums, float* row_stats, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  DeviceGuard device_guard(rank_recs);
  if (!rank_recs || !row_target || !sums || !row_stats || !workspace || G <= 0 || R <= 0 || !(tau > 0.f)) {
    set_error("kd_fused_merge_ranks: bad arguments");
    return 1;
  }
  if (workspace_bytes < kd_fused_merge_workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0) {
    set_error("kd_fused_merge_ranks: workspace too small or misaligned (need %zu bytes)",
              kd_fused_merge_workspace_bytes());
    return 1;
  }
  RankMergeParams mp = {};
  mp.rec = rank_recs;
  mp.G = G;
  mp.R = R;
  mp.row_target = row_target;
  mp.dense = teacher_kind == KD_TEACHER_DENSE ? 1 : 0;
  mp.sparse = teacher_kind == KD_TEACHER_SPARSE ? 1 : 0;
  mp.inv_tau = 1.0f / tau;
  mp.row_stats = row_stats;
  mp.block_sums = reinterpret_cast<float*>(workspace);
  int blocks = cdiv(R, 256);
  if (blocks > kMergeBlocksMax) blocks = kMergeBlocksMax;
  cudaStream_t s = (cudaStream_t)stream;
  kd_fused_rank_merge_kernel<<<blocks, 256, 0, s>>>(mp);
  if (check_launch("kd_fused_rank_merge launch")) return 1;
  return reduce_partials(mp.block_sums, blocks, sums, s);
}

static int fused_fwd_impl(const void* h, int64_t h_stride, const void* W, int64_t w_stride, int teacher_kind,
                          const void* y, int y_dtype, int64_t y_stride, const float* topk_v, const int32_t* topk_i,
                          int K, const int32_t* row_target, const int32_t* n_rows, int R, int H, int V, int v_offset,
                          float tau, float* sums, float* row_stats, float* rank_rec, void* logit_cache,
                          size_t logit_cache_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  DeviceGuard device_guard(h);
  if (check_common(h, h_stride, W, w_stride, R, H, V, tau, "kd_fused_linear_fwd")) return 1;
  if (!row_target || !workspace) {
    set_error("kd_fused_linear_fwd: null pointer argument");
    return 1;
  }
  if (teacher_kind == KD_TEACHER_DENSE && (!y || (y_dtype != KD_DTYPE_BF16 && y_dtype != KD_DTYPE_F32))) {
    set_error("kd_fused_linear_fwd: dense teacher must be bf16 or fp32");
    return 1;
  }
  const bool sparse = teacher_kind == KD_TEACHER_SPARSE;
  const Workspace ws = plan_workspace(R, H, V, 0, sparse ? K : 0);
  if (workspace_bytes < ws.fwd_bytes + ws.sp_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("kd_fused_linear_fwd: workspace too small or not 256-byte aligned (need %zu)", ws.total);
    return 1;
  }
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* wsp = reinterpret_cast<uint8_t*>(workspace);
  float* partials = reinterpret_cast<float*>(wsp + ws.partials_off);
  float* bsums = reinterpret_cast<float*>(wsp + ws.bsums_off);

  CUtensorMap ta, tb;
  if (make_tmap(&ta, h, (uint64_t)H, (uint64_t)R, (uint64_t)h_stride, BM, "hidden")) return 1;
  if (make_tmap(&tb, W, (uint64_t)H, (uint64_t)V, (uint64_t)w_stride, b_box_rows(), "lm_head weight")) return 1;
  Geom g = {};
  g.num_m_blk = cdiv(R, tile_m());
  g.num_n_blk = cdiv(V, BN);
  g.num_k_blk = cdiv(H, BK);
  g.n_per_unit = kFwdTilesPerRange;
  const int num_ranges = cdiv(g.num_n_blk, g.n_per_unit);
  g.num_units = g.num_m_blk * num_ranges;
  g.n_rows = n_rows;
  g.rows_dim = n_rows ? 1 : 0;

  FwdParams fp = {};
  fp.row_target = row_target;
  fp.y = y;
  fp.y_stride = y_stride;
  const size_t ys = y_dtype == KD_DTYPE_F32 ? 4 : 2;
  fp.y_vec_ok = (y && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (y_stride * ys) % 16 == 0) ? 1 : 0;
  fp.R = R;
  fp.V = V;
  fp.inv_tau = 1.0f / tau;
  fp.partials = partials;
  fp.label_off = v_offset;
  {
    const char* e = getenv("KD_DEBUG_SKIP_MATH");
    fp.debug_skip_math = (e && e[0] == '1') ? 1 : 0;
  }
  const bool tau2 = tau == 2.0f;
  int rc;
  CUtensorMap ty, tz = ta;
  const LogitCache zc = plan_logit_cache(logit_cache, logit_cache_bytes, R, V);
  if (logit_cache != nullptr && logit_cache_bytes > 0 && zc.tiles == 0) {
    set_error("kd_fused_linear_fwd: logit cache must be 256-byte aligned and hold at least one 256-column tile "
              "(%zu bytes for R=%d)", zc_tile_bytes(R), R);
    return 1;
  }
  if (zc.tiles > 0) {
    fp.zc_tiles = zc.tiles;
    fp.zc_ref = zc.ref;
    if (make_tmap(&tz, zc.z16, (uint64_t)zc.z_stride, (uint64_t)R, (uint64_t)zc.z_stride, 32, "logit cache", 32))
      return 1;
  }
  const float4* sp_rowc = nullptr;
  const bool y_tma = teacher_kind == KD_TEACHER_DENSE && make_teacher_tmap(&ty, y, y_dtype, y_stride, R, V);
  if (teacher_kind == KD_TEACHER_DENSE) {
    rc = y_dtype == KD_DTYPE_BF16 ? launch_fwd<true, __nv_bfloat16>(ta, tb, y_tma ? &ty : nullptr, tz, g, fp, tau2, s)
                                  : launch_fwd<true, float>(ta, tb, nullptr, tz, g, fp, tau2, s);
  } else if (sparse) {
    if (prepare_sparse(topk_v, topk_i, K, row_target, R, V, v_offset, tau, ws, wsp + ws.fwd_bytes, &fp.sp, &sp_rowc, s,
                       "kd_fused_linear_fwd"))
      return 1;
    rc = tau2 ? launch_umma<FwdEpi<__nv_bfloat16, false, true, false, true>, false, false>(ta, tb, ta, tz, g, fp, s)
              : launch_umma<FwdEpi<__nv_bfloat16, false, false, false, true>, false, false>(ta, tb, ta, tz, g, fp, s);
  } else {
    rc = launch_fwd<false, __nv_bfloat16>(ta, tb, nullptr, tz, g, fp, tau2, s);
  }
  if (rc) return rc;

  MergeParams mp = {};
  mp.partials = partials;
  mp.nrec = num_ranges * kColGroups;
  mp.R = R;
  mp.V = V;
  mp.row_target = row_target;
  mp.y = y;
  mp.y_dtype = y_dtype;
  mp.y_stride = y_stride;
  mp.dense = teacher_kind == KD_TEACHER_DENSE ? 1 : 0;
  mp.sparse = sparse ? 1 : 0;
  mp.sp_rowc = sp_rowc;
  mp.inv_tau = 1.0f / tau;
  mp.row_stats = row_stats;
  mp.block_sums = bsums;
  mp.rank_rec = rank_rec;
  mp.label_off = v_offset;
  int blocks = cdiv(R, 8);
  if (blocks > kMergeBlocksMax) blocks = kMergeBlocksMax;
  kd_fused_merge_kernel<<<blocks, 256, 0, s>>>(mp);
  if (check_launch("kd_fused_merge launch")) return 1;
  if (rank_rec != nullptr) return 0;  // vocab-parallel: the cross-rank merge finalises
  return reduce_partials(bsums, blocks, sums, s);
}

extern "C" int kd_fused_linear_bwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride, int teacher_kind,
                                   const void* y, int y_dtype, int64_t y_stride, const float* topk_v,
                                   const int32_t* topk_i, int K, const int32_t* row_target, const int32_t* n_rows,
                                   const float* row_stats, int R, int H, int V, float tau, const int32_t* n_norm,
                                   const float* grad_coef, int grad_dtype, void* dH, int64_t dh_stride, void* dW,
                                   int64_t dw_stride, int64_t dw_row_begin, int v_chunk, const void* logit_cache,
                                   size_t logit_cache_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  return kd_fused_linear_bwd_range(h, h_stride, W, w_stride, teacher_kind, y, y_dtype, y_stride, topk_v, topk_i, K,
                                   row_target, n_rows, row_stats, R, H, V, tau, n_norm, grad_coef, grad_dtype, dH, dh_stride,
                                   dW, dw_stride, dw_row_begin, v_chunk, 0, V, KD_RANGE_FIRST | KD_RANGE_LAST, 0, 0,
                                   logit_cache, logit_cache_bytes, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int kd_fused_bwd_trace_begin(void) {
  std::lock_guard<std::mutex> lock(g_trace.mu);
  for (const TraceRec& r : g_trace.recs) {
    g_trace.pool.push_back(r.e0);
    g_trace.pool.push_back(r.e1);
  }
  g_trace.recs.clear();
  g_trace.on = true;
  return 0;
}

extern "C" int kd_fused_bwd_trace_read(float* host_out, int max_records) {
  // host_out: HOST float[max_records][4] = (class, chunk, start ms, end ms); returns the number of records (or -1)
  std::lock_guard<std::mutex> lock(g_trace.mu);
  g_trace.on = false;
  if (g_trace.recs.empty()) return 0;
  if (check_cuda(cudaDeviceSynchronize(), "trace sync")) return -1;
  const cudaEvent_t t0 = g_trace.recs[0].e0;
  int n = 0;
  for (const TraceRec& r : g_trace.recs) {
    if (n >= max_records) break;
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, t0, r.e0) != cudaSuccess || cudaEventElapsedTime(&b, t0, r.e1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    host_out[4 * n] = (float)r.cls;
    host_out[4 * n + 1] = (float)r.chunk;
    host_out[4 * n + 2] = a;
    host_out[4 * n + 3] = b;
    ++n;
  }
  return n;
}

struct SmLimitScope {
  explicit SmLimitScope(int n) { tl_sm_limit = n; }
  ~SmLimitScope() { tl_sm_limit = 0; }
};

extern "C" int kd_fused_linear_bwd_range(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                                         int teacher_kind, const void* y, int y_dtype, int64_t y_stride,
                                         const float* topk_v, const int32_t* topk_i, int K, const int32_t* row_target,
                                         const int32_t* n_rows, const float* row_stats, int R, int H, int V, float tau,
                                         const int32_t* n_norm,
                                         const float* grad_coef, int grad_dtype, void* dH, int64_t dh_stride, void* dW,
                                         int64_t dw_stride, int64_t dw_row_begin, int v_chunk, int v_begin, int v_end,
                                         int range_flags, int sm_limit, int v_offset, const void* logit_cache,
                                         size_t logit_cache_bytes, void* workspace, size_t workspace_bytes,
                                         void* dw_ready_stream, void* stream) {
  DeviceGuard device_guard(h);
  if (check_common(h, h_stride, W, w_stride, R, H, V, tau, "kd_fused_linear_bwd")) return 1;
  if (v_begin < 0 || v_end > V || v_begin >= v_end || v_begin % BN != 0) {
    set_error("kd_fused_linear_bwd_range: bad vocabulary range [%d, %d) (V=%d; begin must be a multiple of %d)",
              v_begin, v_end, V, BN);
    return 1;
  }
  SmLimitScope sm_scope(sm_limit);
  if (!row_target || !row_stats || !n_norm || !grad_coef || !workspace || (!dH && !dW)) {
    set_error("kd_fused_linear_bwd: null pointer argument");
    return 1;
  }
  if (teacher_kind == KD_TEACHER_DENSE && (!y || (y_dtype != KD_DTYPE_BF16 && y_dtype != KD_DTYPE_F32))) {
    set_error("kd_fused_linear_bwd: dense teacher must be bf16 or fp32");
    return 1;
  }
  const int vc = norm_v_chunk(v_chunk, V);
  const bool sparse = teacher_kind == KD_TEACHER_SPARSE;
  const Workspace ws = plan_workspace(R, H, V, vc, sparse ? K : 0);
  if (workspace_bytes < ws.bwd_bytes + ws.sp_bytes || (reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("kd_fused_linear_bwd: workspace too small or not 256-byte aligned (need %zu)", ws.total);
    return 1;
  }
  const int dw_dtype = grad_dtype & 0xff;
  if (dw_dtype != KD_DTYPE_BF16 && dw_dtype != KD_DTYPE_F32) {
    set_error("kd_fused_linear_bwd: grad_dtype must be KD_DTYPE_BF16 or KD_DTYPE_F32 (optionally | KD_GRAD_DH_F32)");
    return 1;
  }
  const LogitCache zc = plan_logit_cache(const_cast<void*>(logit_cache), logit_cache_bytes, R, V);
  const bool out32 = dw_dtype == KD_DTYPE_F32;                            // dW
  const bool dh_out32 = out32 || (grad_dtype & KD_GRAD_DH_F32) != 0;      // dH
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* wsp = reinterpret_cast<uint8_t*>(workspace);
  float* dh32 = reinterpret_cast<float*>(wsp + ws.dh_off);
  const bool tau2 = tau == 2.0f;
  const size_t ys = y_dtype == KD_DTYPE_F32 ? 4 : 2;
  const int n_chunks = cdiv(v_end - v_begin, vc);
  const bool range_first = (range_flags & KD_RANGE_FIRST) != 0, range_last = (range_flags & KD_RANGE_LAST) != 0;

  CUtensorMap t_h_k, t_w_k, t_g_k[2], t_g_mn[2], t_h_mn, t_w_mn, t_y;
  const bool y_tma = teacher_kind == KD_TEACHER_DENSE && make_teacher_tmap(&t_y, y, y_dtype, y_stride, R, V);
  if (make_tmap(&t_h_k, h, (uint64_t)H, (uint64_t)R, (uint64_t)h_stride, BM, "hidden")) return 1;
  if (make_tmap(&t_w_k, W, (uint64_t)H, (uint64_t)V, (uint64_t)w_stride, b_box_rows(), "lm_head weight")) return 1;
  for (int b = 0; b < 2; ++b) {
    const __nv_bfloat16* Gb = reinterpret_cast<const __nv_bfloat16*>(wsp + ws.g_off + b * ws.g_buf_bytes);
    if (make_tmap(&t_g_k[b], Gb, (uint64_t)vc, (uint64_t)R, (uint64_t)vc, BM, "G (K-major)")) return 1;
    if (make_tmap(&t_g_mn[b], Gb, (uint64_t)vc, (uint64_t)R, (uint64_t)vc, 64, "G (MN-major)")) return 1;
  }
  if (make_tmap(&t_h_mn, h, (uint64_t)H, (uint64_t)R, (uint64_t)h_stride, 64, "hidden (MN-major)")) return 1;
  if (make_tmap(&t_w_mn, W, (uint64_t)H, (uint64_t)V, (uint64_t)w_stride, 64, "lm_head weight (MN-major)")) return 1;
  // fp16 gradient operand: dW and dH multiply the fp16 G with fp16 copies of h (once per call) and of the chunk's W rows
  const bool g16 = g_fp16_enabled();
  uint8_t* h16 = wsp + ws.h16_off;
  if (g16 && make_tmap(&t_h_mn, h16, (uint64_t)H, (uint64_t)R, (uint64_t)H, 64, "hidden fp16 (MN-major)")) return 1;

  SparseView sp_view = {};
  if (sparse && prepare_sparse(topk_v, topk_i, K, row_target, R, V, v_offset, tau, ws, wsp + ws.bwd_bytes, &sp_view, nullptr, s,
                               "kd_fused_linear_bwd"))
    return 1;

  // three chains (grad on the caller's stream, dW, dH) when the pipeline is on; one serial chain otherwise
  // a backward split into vocabulary ranges uses the pipeline in every range (its dH chain runs across them)
  const bool multi_range = !(range_first && range_last);
  BwdPipe* pipe = (bwd_pipe_enabled() && (n_chunks > 1 || multi_range)) ? get_bwd_pipe() : nullptr;
  std::unique_lock<std::mutex> pipe_lock;
  if (pipe) pipe_lock = std::unique_lock<std::mutex>(pipe->enqueue);  // host threads sharing a device take turns
  cudaStream_t s_g = pipe ? pipe->sg : s, s_w = pipe ? pipe->sw : s, s_h = pipe ? pipe->sh : s;
  BwdPipe local_state;  // serial mode: the flags are unused
  BwdPipe& st = pipe ? *pipe : local_state;
  if (pipe && !range_first && pipe->owner != workspace) {
    // another backward used the pipeline between two ranges of this one (a second host thread): its state is not
    // ours.  Drain both side chains into the caller's stream and start from a clean slate - rare, and always safe.
    if (check_cuda(cudaEventRecord(pipe->ew[0], s_w), "drain dW")) return 1;
    if (check_cuda(cudaEventRecord(pipe->eh[0], s_h), "drain dH")) return 1;
    if (check_cuda(cudaEventRecord(pipe->ejoin, s_g), "drain grad")) return 1;
    if (check_cuda(cudaStreamWaitEvent(s, pipe->ew[0], 0), "drain dW")) return 1;
    if (check_cuda(cudaStreamWaitEvent(s, pipe->eh[0], 0), "drain dH")) return 1;
    if (check_cuda(cudaStreamWaitEvent(s, pipe->ejoin, 0), "drain grad")) return 1;
    st.rec_w[0] = st.rec_w[1] = st.rec_h[0] = st.rec_h[1] = false;
    st.any_w = st.any_h = false;
  }
  if (range_first) {
    st.rec_w[0] = st.rec_w[1] = st.rec_h[0] = st.rec_h[1] = false;
    st.any_w = st.any_h = false;
  }
  if (pipe) pipe->owner = workspace;
  if (pipe) {  // fork: the gradient chain starts behind everything the caller has enqueued so far
    if (check_cuda(cudaEventRecord(pipe->efork, s), "fork")) return 1;
    if (check_cuda(cudaStreamWaitEvent(s_g, pipe->efork, 0), "fork")) return 1;
  }
  if (g16 && range_first) {
    // fp16 copy of h, once per backward: later ranges reuse it in the same workspace while dW kernels may still read it
    TraceScope ts(kTraceCast, -1, s_g);
    if (cast_bf16_f16(h, h_stride, h16, R, H, s_g)) return 1;
  }
  bool (&rec_w)[2] = st.rec_w;
  bool (&rec_h)[2] = st.rec_h;
  bool& any_w = st.any_w;
  bool& any_h = st.any_h;
  int& last_w = st.last_w;
  int& last_h = st.last_h;

  for (int c = 0; c < n_chunks; ++c) {
    const int b = c & 1;
    const int v0 = v_begin + c * vc;
    const int cols = v_end - v0 < vc ? v_end - v0 : vc;
    const int n_blks = cdiv(cols, BN);  // 256-wide column blocks of this chunk (G is zero-padded to the block)
    const bool need_dw = dW != nullptr && (int64_t)(v0 + cols) > dw_row_begin;
    // ---- 1. gradient chunk G: from the logit cache, or by recomputing the logits tile ----
    {
      if (pipe) {  // the buffer's previous readers (chunk c - 2) must be done
        if (rec_w[b] && check_cuda(cudaStreamWaitEvent(s_g, pipe->ew[b], 0), "wait dW")) return 1;
        if (rec_h[b] && check_cuda(cudaStreamWaitEvent(s_g, pipe->eh[b], 0), "wait dH")) return 1;
        rec_w[b] = rec_h[b] = false;
        if (!grad_cached_overlap()) {  // experiment: no gradient kernel beside the previous chunk's GEMMs
          if (rec_w[b ^ 1] && check_cuda(cudaStreamWaitEvent(s_g, pipe->ew[b ^ 1], 0), "wait dW")) return 1;
          if (rec_h[b ^ 1] && check_cuda(cudaStreamWaitEvent(s_g, pipe->eh[b ^ 1], 0), "wait dH")) return 1;
        }
      }
      if (g16 && dH) {  // this chunk's W rows as fp16 (read by dH(c); the buffer's previous reader was dH(c - 2))
        const uint8_t* wrow = reinterpret_cast<const uint8_t*>(W) + (size_t)v0 * (size_t)w_stride * 2;
        TraceScope ts(kTraceCast, c, s_g);
        if (cast_bf16_f16(wrow, w_stride, wsp + ws.w16_off + b * ws.w16_buf_bytes, cols, H, s_g)) return 1;
      }
      int rc;
      if (v0 / BN + n_blks <= zc.tiles) {
        // the forward kept this chunk's logits: elementwise gradient kernel (HBM-bound) instead of the recompute GEMM
        GradCachedParams cp = {};
        cp.zc = zc;
        cp.y = y;
        cp.y_stride = y_stride;
        cp.y_vec_ok = (y && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (y_stride * ys) % 16 == 0) ? 1 : 0;
        cp.row_target = row_target;
        cp.row_stats = row_stats;
        cp.n_norm = n_norm;
        cp.coef = grad_coef;
        cp.n_rows = n_rows;
        cp.tau = tau;
        cp.use_kl = teacher_kind == KD_TEACHER_NONE ? 0 : 1;
        cp.g_fp16 = g16 ? 1 : 0;
        cp.R = R;
        cp.V = V;
        cp.v0 = v0;
        cp.cols_pad = n_blks * BN;
        cp.label_off = v_offset;
        cp.G = wsp + ws.g_off + b * ws.g_buf_bytes;
        cp.g_stride = vc;
        cp.sp = sp_view;
        TraceScope ts(kTraceGrad, c, s_g);
        rc = launch_grad_cached(cp, teacher_kind, y_dtype, tau2, s_g);
      } else {
      TraceScope ts(kTraceGradRecompute, c, s_g);
      Geom g = {};
      g.num_m_blk = cdiv(R, tile_m());
      g.num_n_blk = n_blks;
      g.num_k_blk = cdiv(H, BK);
      g.b_n0 = v0;
      g.n_per_unit = 1;
      g.num_units = g.num_m_blk * g.num_n_blk;
      g.n_rows = n_rows;
      g.rows_dim = n_rows ? 1 : 0;
      GradParams gp = {};
      gp.row_target = row_target;
      gp.row_stats = row_stats;
      gp.y = y;
      gp.y_stride = y_stride;
      gp.y_vec_ok = (y && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (y_stride * ys) % 16 == 0) ? 1 : 0;
      gp.R = R;
      gp.V = V;
      gp.tau = tau;
      gp.use_kl = teacher_kind == KD_TEACHER_NONE ? 0 : 1;
      gp.n_norm = n_norm;
      gp.coef = grad_coef;
      gp.v0 = v0;
      gp.label_off = v_offset;
      gp.g_fp16 = g16 ? 1 : 0;
      gp.sp = sp_view;
      if (teacher_kind == KD_TEACHER_DENSE) {
        rc = y_dtype == KD_DTYPE_BF16
                 ? launch_grad<true, __nv_bfloat16>(t_h_k, t_w_k, y_tma ? &t_y : nullptr, t_g_k[b], g, gp, tau2, s_g)
                 : launch_grad<true, float>(t_h_k, t_w_k, nullptr, t_g_k[b], g, gp, tau2, s_g);
      } else if (sparse) {
        rc = tau2 ? launch_umma<GradEpi<__nv_bfloat16, false, true, false, true>, false, false>(t_h_k, t_w_k, t_h_k,
                                                                                                  t_g_k[b], g, gp, s_g)
                  : launch_umma<GradEpi<__nv_bfloat16, false, false, false, true>, false, false>(t_h_k, t_w_k, t_h_k,
                                                                                                   t_g_k[b], g, gp, s_g);
      } else {
        rc = launch_grad<false, __nv_bfloat16>(t_h_k, t_w_k, nullptr, t_g_k[b], g, gp, tau2, s_g);
      }
      }
      if (rc) return rc;
      if (pipe && check_cuda(cudaEventRecord(pipe->eg[b], s_g), "record grad")) return 1;
    }
    // ---- 2. dW[v0 : v0+cols, :] = G^T h   (rows are final: every token is in this GEMM's K) ----
    if (need_dw) {
      if (pipe && check_cuda(cudaStreamWaitEvent(s_w, pipe->eg[b], 0), "wait grad")) return 1;
      Geom g = {};
      g.num_m_blk = cdiv(n_blks * BN, tile_m());
      g.num_n_blk = cdiv(H, BN);
      g.num_k_blk = cdiv(R, BK);
      g.n_per_unit = 1;
      g.num_units = g.num_m_blk * g.num_n_blk;
      g.n_rows = n_rows;
      g.rows_dim = n_rows ? 2 : 0;  // K runs over rows: stop at the last live k-block
      g.ab_fp16 = g16 ? 1 : 0;
      g.n_fast = dw_n_fast() ? 1 : 0;  // the h tiles of one vocabulary tile run side by side and share G^T through L2
      StoreParams sp = {};
      if (g16) {
        sp.scale_n_norm = n_norm;
        sp.scale_coef = grad_coef;
        sp.scale_tau = tau;
      }
      sp.mode = out32 ? kStoreF32 : kStoreBf16;
      sp.m_total = cols;
      sp.n_total = H;
      const int64_t mb = dw_row_begin - v0;
      sp.m_begin = mb > 0 ? (int)mb : 0;
      sp.c16 = reinterpret_cast<__nv_bfloat16*>(dW);
      sp.ld16 = dw_stride;
      sp.row0_16 = v0;
      sp.c32 = reinterpret_cast<float*>(dW);
      sp.ld32 = dw_stride;
      sp.row0_32 = v0;
      {
        TraceScope ts(kTraceDw, c, s_w);
        if (launch_umma<StoreEpi, true, true>(t_g_mn[b], t_h_mn, g, sp, s_w)) return 1;
      }
      if (pipe) {
        if (check_cuda(cudaEventRecord(pipe->ew[b], s_w), "record dW")) return 1;
        rec_w[b] = any_w = true;
        last_w = b;
      }
    }
    // ---- 3. dH (+)= G W[v0 : v0+cols, :] ----
    if (dH) {
      if (pipe && check_cuda(cudaStreamWaitEvent(s_h, pipe->eg[b], 0), "wait grad")) return 1;
      Geom g = {};
      g.num_m_blk = cdiv(R, tile_m());
      g.num_n_blk = cdiv(H, BN);
      g.num_k_blk = n_blks * (BN / BK);
      g.b_k0 = v0;
      g.n_per_unit = 1;
      g.num_units = g.num_m_blk * g.num_n_blk;
      g.n_rows = n_rows;
      g.rows_dim = n_rows ? 1 : 0;
      StoreParams sp = {};
      CUtensorMap t_w16_mn;
      if (g16) {
        // the chunk's fp16 rows start at 0 in their own buffer; rows beyond `cols` are out of bounds = zero
        g.b_k0 = 0;
        g.ab_fp16 = 1;
        if (make_tmap(&t_w16_mn, wsp + ws.w16_off + b * ws.w16_buf_bytes, (uint64_t)H, (uint64_t)cols, (uint64_t)H, 64,
                      "lm_head weight chunk fp16 (MN-major)"))
          return 1;
        sp.scale_n_norm = n_norm;
        sp.scale_coef = grad_coef;
        sp.scale_tau = tau;
      }
      const bool first = range_first && c == 0, last = range_last && c == n_chunks - 1;
      sp.m_total = R;
      sp.n_total = H;
      sp.m_begin = 0;
      if (dh_out32) {  // accumulate straight into the caller's fp32 dH
        sp.mode = first ? kStoreF32 : kAccumF32;
        sp.c32 = reinterpret_cast<float*>(dH);
        sp.ld32 = dh_stride;
      } else {
        sp.mode = first && last ? kStoreBf16 : (first ? kStoreF32 : (last ? kFinalBf16 : kAccumF32));
        sp.c32 = dh32;
        sp.ld32 = H;
        sp.c16 = reinterpret_cast<__nv_bfloat16*>(dH);
        sp.ld16 = dh_stride;
      }
      {
        TraceScope ts(kTraceDh, c, s_h);
        if (launch_umma<StoreEpi, false, true>(t_g_k[b], g16 ? t_w16_mn : t_w_mn, g, sp, s_h)) return 1;
      }
      if (pipe) {
        if (check_cuda(cudaEventRecord(pipe->eh[b], s_h), "record dH")) return 1;
        rec_h[b] = any_h = true;
        last_h = b;
      }
    }
  }
  if (!pipe && dw_ready_stream != nullptr && !range_last) {
    // serial mode (KD_BWD_STREAMS=0, single chunk): everything ran on the caller's stream
    BwdPipe* p2 = get_bwd_pipe();
    if (p2) {
      if (check_cuda(cudaEventRecord(p2->ejoin, s), "ready")) return 1;
      if (check_cuda(cudaStreamWaitEvent((cudaStream_t)dw_ready_stream, p2->ejoin, 0), "ready")) return 1;
    }
  }
  if (pipe) {
    // join: both side chains are serial, so their last events cover everything.  dW is joined after every range
    // (its rows are handed to the all-reduce), dH only after the last one: between ranges the dH chain keeps
    // running and the next range's gradient kernels wait per buffer, as inside a range.
    if (check_cuda(cudaEventRecord(pipe->ejoin, s_g), "join grad")) return 1;
    if (check_cuda(cudaStreamWaitEvent(s, pipe->ejoin, 0), "join grad")) return 1;
    // Between ranges the caller's stream must NOT wait for the dW chain when the consumer of the finished rows (the
    // gradient all-reduce) runs on a stream of its own: the next range's gradient chain forks from the caller's
    // stream, and a join here drains the three-chain pipeline at every range boundary (measured: ~0.1 ms per
    // boundary, the larger part of the 0.6 ms the overlapped all-reduce cost at 8 GPUs).  With dw_ready_stream the
    // rows' completion is handed to that stream instead; the last range joins everything into the caller's stream.
    cudaStream_t s_ready = (dw_ready_stream != nullptr && !range_last) ? (cudaStream_t)dw_ready_stream : s;
    if (any_w && check_cuda(cudaStreamWaitEvent(s_ready, pipe->ew[last_w], 0), "join dW")) return 1;
    if (range_last && any_h && check_cuda(cudaStreamWaitEvent(s, pipe->eh[last_h], 0), "join dH")) return 1;
  }
  return 0;
}

// Stage-1 entry points named in SURVEY.md 8b: causal-LM cross-entropy through the LM head (no teacher) with the
// frozen-vocabulary mask folded into the dW GEMM (dw_row_begin = V_old).  Thin forwards of the general calls.
extern "C" int kd_ce_fused_linear_fwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                                      const int32_t* row_target, const int32_t* n_rows, int R, int H, int V,
                                      float* sums, float* row_stats, void* logit_cache, size_t logit_cache_bytes,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  return kd_fused_linear_fwd(h, h_stride, W, w_stride, KD_TEACHER_NONE, nullptr, 0, 0, nullptr, nullptr, 0, row_target,
                             n_rows, R, H, V, 1.0f, 1.0f, sums, row_stats, logit_cache, logit_cache_bytes, workspace,
                             workspace_bytes, stream);
}
extern "C" int kd_ce_fused_linear_bwd(const void* h, int64_t h_stride, const void* W, int64_t w_stride,
                                      const int32_t* row_target, const int32_t* n_rows, const float* row_stats, int R,
                                      int H, int V, const int32_t* n_norm, const float* grad_coef, int grad_dtype,
                                      void* dH, int64_t dh_stride, void* dW, int64_t dw_stride, int64_t dw_row_begin,
                                      int v_chunk, const void* logit_cache, size_t logit_cache_bytes, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  return kd_fused_linear_bwd(h, h_stride, W, w_stride, KD_TEACHER_NONE, nullptr, 0, 0, nullptr, nullptr, 0, row_target,
                             n_rows, row_stats, R, H, V, 1.0f, n_norm, grad_coef, grad_dtype, dH, dh_stride, dW,
                             dw_stride, dw_row_begin, v_chunk, logit_cache, logit_cache_bytes, workspace,
                             workspace_bytes, stream);
}

// out[R, V] (bf16, row stride out_stride) = h[R, H] * W[V, H]^T : the LM head alone, on the K1 pipeline
// (CTA pairs, fp32 accumulation in TMEM, tile -> swizzled smem -> TMA store).  Used for the teacher head in front
// of the top-k compaction (train.py:60-94, extract_teacher_logits.py:109-129), block of rows by block of rows.
extern "C" int kd_linear_bf16(const void* h, int64_t h_stride, const void* W, int64_t w_stride, void* out,
                              int64_t out_stride, int R, int H, int V, void* stream) {
  DeviceGuard device_guard(h);
  if (check_common(h, h_stride, W, w_stride, R, H, V, 1.0f, "kd_linear_bf16")) return 1;
  if (!out) {
    set_error("kd_linear_bf16: null output");
    return 1;
  }
  CUtensorMap ta, tb, tg;
  if (make_tmap(&ta, h, (uint64_t)H, (uint64_t)R, (uint64_t)h_stride, BM, "hidden")) return 1;
  if (make_tmap(&tb, W, (uint64_t)H, (uint64_t)V, (uint64_t)w_stride, b_box_rows(), "lm_head weight")) return 1;
  if (make_tmap(&tg, out, (uint64_t)V, (uint64_t)R, (uint64_t)out_stride, BM, "logits out")) return 1;
  Geom g = {};
  g.num_m_blk = cdiv(R, tile_m());
  g.num_n_blk = cdiv(V, BN);
  g.num_k_blk = cdiv(H, BK);
  g.n_per_unit = 1;
  g.num_units = g.num_m_blk * g.num_n_blk;
  GradParams gp = {};
  gp.R = R;
  gp.V = V;
  gp.tau = 1.0f;
  return launch_umma<GradEpi<__nv_bfloat16, false, false, false, false, true>, false, false>(ta, tb, ta, tg, g, gp,
                                                                                             (cudaStream_t)stream);
}

// ---- teacher head -> top-k without the teacher's [R,V] logits: GEMM + selection statistics --------------------
// Work units of the head GEMM: n_per_unit consecutive 256-column tiles of one row block (the online (m, s) of a
// thread lives in registers across them).  A row block is only a few m blocks tall, so the unit count is chosen to
// fill whole waves of the persistent CTA pairs: the largest n_per_unit within 2 % of the best wave efficiency.
static int head_n_per_unit(int R, int V) {
  const int m_blks = cdiv(R, tile_m()), n_blks = cdiv(V, BN);
  const int slots = sm_count() / cta_group() >= 1 ? sm_count() / cta_group() : 1;
  double best = 0.0;
  double eff[33] = {};
  for (int npu = 1; npu <= 32; ++npu) {
    const int units = m_blks * cdiv(n_blks, npu);
    const int waves = cdiv(units, slots);
    eff[npu] = (double)m_blks * n_blks / ((double)waves * slots * npu);
    if (eff[npu] > best) best = eff[npu];
  }
  for (int npu = 32; npu >= 1; --npu)
    if (eff[npu] >= best - 0.02) return npu;
  return 1;
}

extern "C" int kd_head_topk_layout(int V, int* pmax_stride, int* part_stride) {
  if (V <= 0 || !pmax_stride || !part_stride) {
    set_error("kd_head_topk_layout: bad argument");
    return 1;
  }
  *pmax_stride = 8 * cdiv(V, BN);           // 32-column pieces of whole 256-column tiles
  *part_stride = kColGroups * cdiv(V, BN);  // upper bound (one tile per range)
  return 0;
}

extern "C" int kd_head_logits_stats(const void* h, int64_t h_stride, const void* W, int64_t w_stride, void* out,
                                    int64_t out_stride, void* pmax, int pmax_stride, void* part, int part_stride,
                                    int* n_part, int R, int H, int V, void* stream) {
  DeviceGuard device_guard(h);
  if (check_common(h, h_stride, W, w_stride, R, H, V, 1.0f, "kd_head_logits_stats")) return 1;
  if (!out || !pmax || !part || !n_part) {
    set_error("kd_head_logits_stats: null output");
    return 1;
  }
  const int npu = head_n_per_unit(R, V);
  const int ranges = cdiv(cdiv(V, BN), npu);
  if (pmax_stride < 8 * cdiv(V, BN) || part_stride < kColGroups * ranges) {
    set_error("kd_head_logits_stats: pmax_stride / part_stride smaller than kd_head_topk_layout reports");
    return 1;
  }
  *n_part = kColGroups * ranges;
  CUtensorMap ta, tb, tg;
  if (make_tmap(&ta, h, (uint64_t)H, (uint64_t)R, (uint64_t)h_stride, BM, "hidden")) return 1;
  if (make_tmap(&tb, W, (uint64_t)H, (uint64_t)V, (uint64_t)w_stride, b_box_rows(), "lm_head weight")) return 1;
  if (make_tmap(&tg, out, (uint64_t)V, (uint64_t)R, (uint64_t)out_stride, 32, "logits out", 32)) return 1;
  Geom g = {};
  g.num_m_blk = cdiv(R, tile_m());
  g.num_n_blk = cdiv(V, BN);
  g.num_k_blk = cdiv(H, BK);
  g.n_per_unit = npu;
  g.num_units = g.num_m_blk * ranges;
  HeadParams hp = {};
  hp.R = R;
  hp.V = V;
  hp.pmax = reinterpret_cast<__nv_bfloat16*>(pmax);
  hp.pmax_stride = pmax_stride;
  hp.part = reinterpret_cast<float2*>(part);
  hp.part_stride = part_stride;
  return launch_umma<HeadEpi, false, false>(ta, tb, ta, tg, g, hp, (cudaStream_t)stream);
}

extern "C" int kd_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                            float* C, int64_t ldc, int M, int N, int K, void* stream) {
  DeviceGuard device_guard(A);
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) {
    set_error("kd_gemm_bf16: bad arguments");
    return 1;
  }
  CUtensorMap ta, tb;
  if (a_mn_major) {
    if (make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, "A (MN-major)")) return 1;
  } else {
    if (make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BM, "A (K-major)")) return 1;
  }
  if (b_mn_major) {
    if (make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, "B (MN-major)")) return 1;
  } else {
    if (make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, b_box_rows(), "B (K-major)")) return 1;
  }
  Geom g = {};
  g.num_m_blk = cdiv(M, tile_m());
  g.num_n_blk = cdiv(N, BN);
  g.num_k_blk = cdiv(K, BK);
  g.n_per_unit = 1;
  g.num_units = g.num_m_blk * g.num_n_blk;
  StoreParams sp = {};
  sp.mode = kStoreF32;
  sp.m_total = M;
  sp.n_total = N;
  sp.c32 = C;
  sp.ld32 = ldc;
  cudaStream_t s = (cudaStream_t)stream;
  if (a_mn_major && b_mn_major) return launch_umma<StoreEpi, true, true>(ta, tb, g, sp, s);
  if (!a_mn_major && b_mn_major) return launch_umma<StoreEpi, false, true>(ta, tb, g, sp, s);
  if (!a_mn_major && !b_mn_major) return launch_umma<StoreEpi, false, false>(ta, tb, g, sp, s);
  set_error("kd_gemm_bf16: the (A MN-major, B K-major) combination is not instantiated");
  return 1;
}
