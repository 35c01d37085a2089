// K2: streaming KD loss + gradient on materialised logits (HBM-bound).
//
// Replaces distillation_loss.py:31-128 and its autograd.  One CTA owns one row (b,t) at a time (rows drawn from an
// atomic counter, ~148 in flight): sweep 1 streams z and y from HBM into the online soft-max statistics, one block
// reduction gives the row's log-sum-exps, sweep 2 re-reads the row from L2 and streams the gradient out.  HBM traffic
// = read z + read y + write dz (6 B / element for bf16), the algorithmic minimum of SURVEY.md 8(d).
// (Round 1 also carried a cluster / shared-memory-stash form and a TMA-ring form; both measured slower - 2210 and
// 1153 us against 1095 us at the configs[1] shape, profiles/r01e_hbm_bench.log - and were removed.)
//
// Per valid row (appendix C of SURVEY.md), all fp32:
//   student: m, S1 = sum e^{z-m}, St = sum e^{(z-m)/tau}
//   teacher: mt, T1 = sum e^{y-mt}, Tt = sum e^{(y-mt)/tau}, A = sum e^{(y-mt)/tau} (y - z)
//   CE = LSE1 - z_l ; KL = A/(tau Tt) - LSEt_tau + LSE_tau ; teacherCE = LSEt1 - y_l
//   G  = c1 (e^{z-LSE1} - [v=l]) + c2 (e^{z/tau-LSE_tau} - P),  c1 = alpha g/N, c2 = (1-alpha) tau g/N
#include <cstdlib>
#include <type_traits>

#include "kd_common.cuh"
#include "kd_umma.cuh"

namespace kd {

constexpr int kMaxTopK = 1024;

struct StreamParams {
  const void* z;
  const void* y;
  int64_t z_sb, z_st, y_sb, y_st;  // strides in elements
  const float* topk_v;
  const int32_t* topk_i;
  int K;
  const int32_t* row_target;
  int B, T, V;
  float tau, alpha, grad_scale;
  const int32_t* n_norm;
  void* dlogits;
  float* partials;  // [CTAs][kNumPartialSlots]
  int vec_ok;       // 16-byte vector path legal for every row pointer
  int n_stash;      // dense + gradient: leading register sets of a row kept in shared memory for sweep 2
  int l2_ahead;     // sweep 1: thread 0 asks the L2 for the row's bytes this many sets ahead of the loads (0 = off)
};

struct Stats7 {
  float m, s1, st;       // student
  float mt, t1, tt, a;   // teacher (dense only)
};

template <bool DENSE>
__device__ __forceinline__ Stats7 warp_merge(Stats7 s, float inv_tau) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, s.m, o);
    const float a2 = __shfl_xor_sync(0xffffffffu, s.s1, o);
    const float b2 = __shfl_xor_sync(0xffffffffu, s.st, o);
    merge_student(s.m, s.s1, s.st, m2, a2, b2, inv_tau);
    if (DENSE) {
      const float mt2 = __shfl_xor_sync(0xffffffffu, s.mt, o);
      const float t12 = __shfl_xor_sync(0xffffffffu, s.t1, o);
      const float tt2 = __shfl_xor_sync(0xffffffffu, s.tt, o);
      const float aa2 = __shfl_xor_sync(0xffffffffu, s.a, o);
      merge_teacher(s.mt, s.t1, s.tt, s.a, mt2, t12, tt2, aa2, inv_tau);
    }
  }
  return s;
}

// ------------------------------------------------------------------------------------------
// K2, row-per-CTA form (default).  One CTA owns a whole row: sweep 1 streams z and y from HBM (16-byte loads,
// the next 16 elements' loads in flight while 16 are being reduced) into the online statistics, one block
// reduction gives the row's log-sum-exps, sweep 2 re-reads the row - 612 KB for bf16 z + y, still resident in
// the 126 MB L2 because only ~148 rows are in flight - and streams the gradient out with evict-first stores.
// HBM traffic stays at the algorithmic 6 B / element; compared with round 1's cluster form there is no shared-
// memory stash, no DSMEM exchange and 8x fewer reductions per row (the cluster form spent ~2/3 of its issued
// instructions on them and is latency-bound at 18 us per row).
// ------------------------------------------------------------------------------------------
constexpr int kRowThreadsMax = 1024;

// one instruction brings a contiguous range of global memory into L2 (no registers, no shared memory, no completion to
// wait for); `bytes` is a multiple of 16
__device__ __forceinline__ void l2_prefetch_bulk(const void* src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy)
               : "memory");
}

struct RowShared {
  double acc_ce, acc_kl, acc_t;  // the CTA's sums over its rows: thread 0 alone reads and writes them, once per row
  int acc_n, acc_hits;           // (in registers they were 8 live values per thread through both sweeps)
  Stats7 warp_stats[kRowThreadsMax / 32];
  Stats7 row_stats;
  float sp_lk;
  int next_row;
};

template <typename TZ, typename TY, bool DENSE, bool TAU2, bool GRAD, int kRowThreads>
__global__ void __launch_bounds__(kRowThreads, kRowThreads <= 256 ? 2 : 1) kd_stream_row_kernel(const StreamParams p, int* __restrict__ row_counter) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ RowShared sh;
  float* sp_p = reinterpret_cast<float*>(dyn_smem);  // sparse teacher only: p_k and i_k of the current row
  int32_t* sp_i = reinterpret_cast<int32_t*>(sp_p + p.K);
  // dense teacher with gradient: the first n_stash register sets of a row (set = the 4 pieces one thread holds: 2 of z,
  // 2 of y) are kept in shared memory by the thread that loaded them, so that sweep 2 reads them back from there and
  // only the rest of the row has to stay in L2: 148 rows x 612 KB = 90 MB of evict_last lines do not fit one 63 MB L2
  // partition (ncu: 24 % of sweep 2's sectors missed, 0.6 GB of extra DRAM reads at the configs[1] shape)
  constexpr bool kStash = DENSE && GRAD;
  constexpr int kSetBytes = kRowThreads * 2 * (int)(sizeof(Vec8<TZ>) + sizeof(Vec8<TY>));
  const int n_stash = kStash ? p.n_stash : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* const stash_z = dyn_smem + (size_t)tid * sizeof(Vec8<TZ>);
  uint8_t* const stash_y = dyn_smem + (size_t)kRowThreads * 2 * sizeof(Vec8<TZ>) + (size_t)tid * sizeof(Vec8<TY>);
  const float inv_tau = 1.0f / p.tau;
  const int V = p.V;
  const bool vec_ok = p.vec_ok != 0;
  const int nvec = vec_ok ? V / 8 : 0;  // 16-byte pieces of a row; [8 nvec, V) is the scalar tail
  const int n_rows = p.B * p.T;
  float norm = 0.f;
  if (GRAD) {
    const int nn = *p.n_norm;
    norm = nn > 0 ? p.grad_scale / (float)nn : 0.f;
  }
  const float c1 = p.alpha * norm;
  const float c2 = (1.f - p.alpha) * p.tau * norm;
  if (tid == 0) {
    sh.acc_ce = sh.acc_kl = sh.acc_t = 0.0;
    sh.acc_n = sh.acc_hits = 0;
  }
  // sweep 1 keeps the row's lines in L2 (evict_last) only when a second sweep will read them
  const uint64_t pol_keep = GRAD ? l2_policy_evict_last() : l2_policy_evict_first();
  const uint64_t pol_drop = l2_policy_evict_first();

  for (;;) {
    // rows are drawn from a counter: rows that are not scored cost a fifth of a scored one
    if (tid == 0) sh.next_row = atomicAdd(row_counter, 1);
    __syncthreads();
    const int row = sh.next_row;
    if (row >= n_rows) break;
    const int b = row / p.T, t = row - b * p.T;
    const int target = p.row_target[row];
    TZ* out_row = GRAD ? reinterpret_cast<TZ*>(p.dlogits) + (size_t)row * V : nullptr;
    if (target < 0) {  // zero gradient, nothing to read
      if (GRAD) {
        constexpr int kPer16 = 16 / sizeof(TZ);
        if (vec_ok) {
          const uint4 zero4 = make_uint4(0, 0, 0, 0);
          const int n16 = V / kPer16;
          for (int i = tid; i < n16; i += kRowThreads) stg_stream(out_row + (size_t)i * kPer16, zero4);
          for (int i = n16 * kPer16 + tid; i < V; i += kRowThreads) out_row[i] = Elem<TZ>::from_f(0.f);
        } else {
          for (int i = tid; i < V; i += kRowThreads) out_row[i] = Elem<TZ>::from_f(0.f);
        }
      }
      __syncthreads();  // sh.next_row is rewritten by the next draw
      continue;
    }
    const TZ* zrow = reinterpret_cast<const TZ*>(p.z) + (int64_t)b * p.z_sb + (int64_t)t * p.z_st;
    const TY* yrow = DENSE ? reinterpret_cast<const TY*>(p.y) + (int64_t)b * p.y_sb + (int64_t)t * p.y_st : nullptr;

    // ---------------- sweep 1: HBM -> registers, online statistics, 16 elements per update ----------------
    Stats7 s;
    s.m = s.mt = -CUDART_INF_F;
    s.s1 = s.st = s.t1 = s.tt = s.a = 0.f;
    {
      // thread `tid` takes pieces tid, tid + T, ... in pairs (q, q + kRowThreads); in the fast loop the next pair is
      // loaded before the current one is reduced
      Vec8<TZ> z0, z1;
      Vec8<TY> y0, y1;
      int q = tid;
      auto load_pair = [&](int qq, Vec8<TZ>& a, Vec8<TZ>& bb, Vec8<TY>& c, Vec8<TY>& d) {
        if (qq < nvec) {
          a.load_global_hint(zrow + (size_t)qq * 8, pol_keep);
          if (DENSE) c.load_global_hint(yrow + (size_t)qq * 8, pol_keep);
        }
        if (qq + kRowThreads < nvec) {
          bb.load_global_hint(zrow + (size_t)(qq + kRowThreads) * 8, pol_keep);
          if (DENSE) d.load_global_hint(yrow + (size_t)(qq + kRowThreads) * 8, pol_keep);
        }
      };
      // Fast loop (ncu: the kernel issues 38 instructions per element at 61 % issue and 56 % XU utilisation with DRAM
      // at 48 % - it is bound by instruction issue, 2.8 of them register copies of the prefetch and ~6 predicates and
      // addressing): whole pairs only, two register sets that alternate instead of being copied, no predicates, the
      // teacher's -inf clamped on the packed words so that the cross term needs no per-element guard.
      {
        struct Set {
          Vec8<TZ> z0, z1;
          Vec8<TY> y0, y1;
        };
        // sets below n_stash go to shared memory as well (thread-private slots: no barrier) and are not kept in L2
        auto load_full = [&](Set& p, int qq, int i) {
          const uint64_t pol = (kStash && i < n_stash) ? pol_drop : pol_keep;
          p.z0.load_global_hint(zrow + (size_t)qq * 8, pol);
          p.z1.load_global_hint(zrow + (size_t)(qq + kRowThreads) * 8, pol);
          if (DENSE) {
            p.y0.load_global_hint(yrow + (size_t)qq * 8, pol);
            p.y1.load_global_hint(yrow + (size_t)(qq + kRowThreads) * 8, pol);
          }
        };
        auto reduce16 = [&](Set& p, int i) {
          if (kStash && i < n_stash) {
            uint8_t* sz = stash_z + (size_t)i * kSetBytes;
            uint8_t* sy = stash_y + (size_t)i * kSetBytes;
            p.z0.store_shared(reinterpret_cast<TZ*>(sz));
            p.z1.store_shared(reinterpret_cast<TZ*>(sz + kRowThreads * sizeof(Vec8<TZ>)));
            p.y0.store_shared(reinterpret_cast<TY*>(sy));
            p.y1.store_shared(reinterpret_cast<TY*>(sy + kRowThreads * sizeof(Vec8<TY>)));
          }
          float fz[16], fy[16], t8[8];
          p.z0.unpack(t8);
#pragma unroll
          for (int j = 0; j < 8; ++j) fz[j] = t8[j];
          p.z1.unpack(t8);
#pragma unroll
          for (int j = 0; j < 8; ++j) fz[8 + j] = t8[j];
          student_update<TAU2, 16, DENSE>(fz, 16, inv_tau, s.m, s.s1, s.st);
          if (DENSE) {
            constexpr bool kPacked = std::is_same<TY, __nv_bfloat16>::value;  // -inf -> most negative finite, 2 per op
            if (kPacked) {
              p.y0.a.x = clamp_neg_inf_bf16x2(p.y0.a.x); p.y0.a.y = clamp_neg_inf_bf16x2(p.y0.a.y);
              p.y0.a.z = clamp_neg_inf_bf16x2(p.y0.a.z); p.y0.a.w = clamp_neg_inf_bf16x2(p.y0.a.w);
              p.y1.a.x = clamp_neg_inf_bf16x2(p.y1.a.x); p.y1.a.y = clamp_neg_inf_bf16x2(p.y1.a.y);
              p.y1.a.z = clamp_neg_inf_bf16x2(p.y1.a.z); p.y1.a.w = clamp_neg_inf_bf16x2(p.y1.a.w);
            }
            p.y0.unpack(t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) fy[j] = t8[j];
            p.y1.unpack(t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) fy[8 + j] = t8[j];
            // The unguarded cross term (one subtract and one FMA per element) is exact for finite student logits once
            // the teacher is clamped.  A -inf student logit gives what the reference gives: inf where the teacher has
            // mass, NaN where it has none (0 * inf inside kl_div: logits masked with -inf on both sides are NaN in the
            // reference too); the guarded form used outside the hot loop (GUARD = 2) does the same.  (Detecting such
            // groups with packed maxima and sending them through a form that returns 0 for the second case cost 6 % of
            // the kernel: 957 -> 1017 us.)
            teacher_update<TAU2, 16, (kPacked ? 0 : 2), true>(fy, fz, 16, inv_tau, s.mt, s.t1, s.tt, s.a);
          }
        };
        // L2 prefetch: the loads of a set leave when the previous set is being reduced, i.e. 64 bytes per thread = 32 KB
        // per SM are on their way to registers at any time, and ncu puts 14 % of all warp samples on the first use
        // of a sweep-1 load.  Thread 0 therefore asks the L2 for sets l2_ahead .. ahead of the loads (16 KB of z and
        // of y per instruction): the register loads then find their lines in L2, whatever DRAM's latency is.
        // (Gradient launches only: the forward-only form - two 256-thread CTAs per SM - measured 3 % slower with it;
        // asking for the head of a row that will be drawn a few microseconds later changed nothing either way.)
        constexpr bool kL2Fetch = GRAD;
        const int l2_ahead = kL2Fetch ? p.l2_ahead : 0;
        auto l2_fetch_set = [&](int i) {
          const int first = i * 2 * kRowThreads;  // first 16-byte piece of set i
          if (first >= nvec) return;
          const int n = nvec - first < 2 * kRowThreads ? nvec - first : 2 * kRowThreads;
          const uint64_t pol = (kStash && i < n_stash) ? pol_drop : pol_keep;
          l2_prefetch_bulk(zrow + (size_t)first * 8, (uint32_t)n * (uint32_t)sizeof(Vec8<TZ>), pol);
          if (DENSE) l2_prefetch_bulk(yrow + (size_t)first * 8, (uint32_t)n * (uint32_t)sizeof(Vec8<TY>), pol);
        };
        if (kL2Fetch && tid == 0 && l2_ahead > 0) {
          for (int i = 1; i <= l2_ahead; ++i) l2_fetch_set(i);  // set 0 is being loaded right now
        }
        if (q + kRowThreads < nvec) {
          Set a, b;
          int si = 0;  // set index of `a`: set i holds pieces q = tid + 2 i S and q + S
          load_full(a, q, 0);
          while (q + 5 * kRowThreads < nvec) {  // pairs q, q + 2S and q + 4S lie inside the row
            if (kL2Fetch && tid == 0 && l2_ahead > 0) {
              l2_fetch_set(si + l2_ahead + 1);
              l2_fetch_set(si + l2_ahead + 2);
            }
            load_full(b, q + 2 * kRowThreads, si + 1);
            reduce16(a, si);
            load_full(a, q + 4 * kRowThreads, si + 2);
            reduce16(b, si + 1);
            q += 4 * kRowThreads;
            si += 2;
          }
          reduce16(a, si);
          q += 2 * kRowThreads;
        }
      }
      for (; q < nvec; q += 2 * kRowThreads) {  // what the fast loop left: at most two pairs, possibly partial
        load_pair(q, z0, z1, y0, y1);
        float fz[16], fy[16];
        const bool two = q + kRowThreads < nvec;
        {
          float t8[8];
          z0.unpack(t8);
#pragma unroll
          for (int j = 0; j < 8; ++j) fz[j] = t8[j];
          if (two) {
            z1.unpack(t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) fz[8 + j] = t8[j];
          }
          if (DENSE) {
            y0.unpack(t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) fy[j] = t8[j];
            if (two) {
              y1.unpack(t8);
#pragma unroll
              for (int j = 0; j < 8; ++j) fy[8 + j] = t8[j];
            }
          }
        }
        if (two) {
          student_update<TAU2, 16>(fz, 16, inv_tau, s.m, s.s1, s.st);
          if (DENSE) teacher_update<TAU2, 16, 2>(fy, fz, 16, inv_tau, s.mt, s.t1, s.tt, s.a);
        } else {
          student_update<TAU2, 16>(fz, 8, inv_tau, s.m, s.s1, s.st);
          if (DENSE) teacher_update<TAU2, 16, 2>(fy, fz, 8, inv_tau, s.mt, s.t1, s.tt, s.a);
        }
      }
      for (int i = nvec * 8 + tid; i < V; i += kRowThreads) {  // scalar tail / unaligned rows
        float fz[8], fy[8];
        fz[0] = Elem<TZ>::to_f(zrow[i]);
        student_update<TAU2, 8>(fz, 1, inv_tau, s.m, s.s1, s.st);
        if (DENSE) {
          fy[0] = Elem<TY>::to_f(yrow[i]);
          teacher_update<TAU2, 8, 2>(fy, fz, 1, inv_tau, s.mt, s.t1, s.tt, s.a);
        }
      }
    }
    // ---------------- block reduction ----------------
    s = warp_merge<DENSE>(s, inv_tau);
    if (lane == 0) sh.warp_stats[warp] = s;
    if (!DENSE && warp == 1) {  // sparse teacher: p_k = softmax(v / tau) over the K entries (:94-95)
      const float* vrow = p.topk_v + (size_t)row * p.K;
      const int32_t* irow = p.topk_i + (size_t)row * p.K;
      float vm = -CUDART_INF_F;
      for (int k = lane; k < p.K; k += 32) vm = fmaxf(vm, vrow[k]);
      vm = warp_max(vm);
      float sum = 0.f;
      for (int k = lane; k < p.K; k += 32) sum += ex2((vrow[k] - vm) * kLog2e * inv_tau);
      sum = warp_sum(sum);
      const float lk = vm * inv_tau + ln_acc(sum);
      for (int k = lane; k < p.K; k += 32) {
        sp_p[k] = __expf(vrow[k] * inv_tau - lk);
        sp_i[k] = irow[k];
      }
      if (lane == 0) sh.sp_lk = lk;
    }
    __syncthreads();
    if (warp == 0) {
      Stats7 w;
      if (lane < kRowThreads / 32) {
        w = sh.warp_stats[lane];
      } else {
        w.m = w.mt = -CUDART_INF_F;
        w.s1 = w.st = w.t1 = w.tt = w.a = 0.f;
      }
      w = warp_merge<DENSE>(w, inv_tau);
      if (lane == 0) sh.row_stats = w;
    }
    __syncthreads();
    const Stats7 f = sh.row_stats;
    const float lse1 = f.m + ln_acc(f.s1);
    const float lset = f.m * inv_tau + ln_acc(f.st);
    const float lsett = DENSE ? f.mt * inv_tau + ln_acc(f.tt) : 0.f;
    const float sp_lk = DENSE ? 0.f : sh.sp_lk;

    // ---------------- per-row scalars ----------------
    if (DENSE) {
      if (tid == 0) {
        // a label outside [0, V) (the reference's F.cross_entropy asserts on it) poisons the loss instead of reading
        // out of bounds: NaN is loud and needs no host sync
        const float zl = target < p.V ? Elem<TZ>::to_f(zrow[target]) : CUDART_NAN_F;
        const float yl = target < p.V ? Elem<TY>::to_f(yrow[target]) : CUDART_NAN_F;
        sh.acc_ce += (double)(lse1 - zl);
        sh.acc_kl += (double)(f.a * inv_tau / f.tt - lsett + lset);
        sh.acc_t += (double)((f.mt + ln_acc(f.t1)) - yl);
        sh.acc_n += 1;
      }
    } else if (warp == 0) {
      // KL_r = sum_k p_k (log p_k - z_{i_k}/tau) + LSE_tau ; monitor hits (distillation_loss.py:104-116)
      const float* vrow = p.topk_v + (size_t)row * p.K;
      float part = 0.f, hsum = 0.f;
      int hits = 0;
      for (int k = lane; k < p.K; k += 32) {
        const int idx = sp_i[k];
        const float pk = sp_p[k];
        const float vk = vrow[k];
        if (idx >= 0 && idx < V) {
          const float zk = Elem<TZ>::to_f(zrow[idx]);
          part += pk * ((vk * inv_tau - sp_lk) - zk * inv_tau);
        }
        if (idx == target) {
          hits += 1;
          hsum += vk;
        }
      }
      part = warp_sum(part);
      hsum = warp_sum(hsum);
      hits = __reduce_add_sync(0xffffffffu, hits);
      if (lane == 0) {
        const float zl = target < p.V ? Elem<TZ>::to_f(zrow[target]) : CUDART_NAN_F;
        sh.acc_ce += (double)(lse1 - zl);
        sh.acc_kl += (double)(part + lset);
        sh.acc_t += (double)hsum;
        sh.acc_hits += hits;
        sh.acc_n += 1;
      }
    }

    // ---------------- sweep 2: gradient; z and y come back from L2 ----------------
    if (GRAD) {
      const float c_tau = kLog2e * inv_tau;
      const float off1 = lse1 * kLog2e, offt = lset * kLog2e, offy = lsett * kLog2e;
      const float half_off1 = off1 * 0.5f;
      const float k_tau = c2 * ex2(half_off1 - offt);  // tau = 2: e^{z/2 - LSE_tau} = E e^{LSE1/2 - LSE_tau}
      auto grad8 = [&](const float(&fz)[8], const float(&fy)[8], int base, int nvalid, float(&g)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < nvalid) {
            float gi;
            if (TAU2) {
              const float e = ex2(fmaf(fz[i], c_tau, -half_off1));
              gi = e * fmaf(e, c1, k_tau);
            } else {
              gi = c1 * ex2(fmaf(fz[i], kLog2e, -off1)) + c2 * ex2(fmaf(fz[i], c_tau, -offt));
            }
            if (DENSE) gi = fmaf(-c2, ex2(fmaf(fy[i], c_tau, -offy)), gi);
            g[i] = gi;
          }
        }
        const unsigned d = (unsigned)(target - base);
        if (d < (unsigned)nvalid) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if ((int)d == i) g[i] -= c1;
        }
      };
      {
        Vec8<TZ> z0, z1;
        Vec8<TY> y0, y1;
        int q = tid;
        auto load_pair = [&](int qq, Vec8<TZ>& a, Vec8<TZ>& bb, Vec8<TY>& c, Vec8<TY>& d) {
          if (qq < nvec) {
            a.load_global_hint(zrow + (size_t)qq * 8, pol_drop);
            if (DENSE) c.load_global_hint(yrow + (size_t)qq * 8, pol_drop);
          }
          if (qq + kRowThreads < nvec) {
            bb.load_global_hint(zrow + (size_t)(qq + kRowThreads) * 8, pol_drop);
            if (DENSE) d.load_global_hint(yrow + (size_t)(qq + kRowThreads) * 8, pol_drop);
          }
        };
        {  // fast loop: whole pairs, alternating register sets, no predicates (see sweep 1)
          struct Set {
            Vec8<TZ> z0, z1;
            Vec8<TY> y0, y1;
          };
          auto load_full = [&](Set& p, int qq, int i) {
            if (kStash && i < n_stash) {  // this thread's own copy from sweep 1 (CTA-uniform branch)
              const uint8_t* sz = stash_z + (size_t)i * kSetBytes;
              const uint8_t* sy = stash_y + (size_t)i * kSetBytes;
              p.z0.load_shared(reinterpret_cast<const TZ*>(sz));
              p.z1.load_shared(reinterpret_cast<const TZ*>(sz + kRowThreads * sizeof(Vec8<TZ>)));
              p.y0.load_shared(reinterpret_cast<const TY*>(sy));
              p.y1.load_shared(reinterpret_cast<const TY*>(sy + kRowThreads * sizeof(Vec8<TY>)));
              return;
            }
            p.z0.load_global_hint(zrow + (size_t)qq * 8, pol_drop);
            p.z1.load_global_hint(zrow + (size_t)(qq + kRowThreads) * 8, pol_drop);
            if (DENSE) {
              p.y0.load_global_hint(yrow + (size_t)qq * 8, pol_drop);
              p.y1.load_global_hint(yrow + (size_t)(qq + kRowThreads) * 8, pol_drop);
            }
          };
          auto emit = [&](const Set& p, int qq) {
            float fz[8], fy[8], g[8];
            Vec8<TZ> vo;
            p.z0.unpack(fz);
            if (DENSE) p.y0.unpack(fy);
            grad8(fz, fy, qq * 8, 8, g);
            vo.pack(g);
            vo.store_global(out_row + (size_t)qq * 8);
            p.z1.unpack(fz);
            if (DENSE) p.y1.unpack(fy);
            grad8(fz, fy, (qq + kRowThreads) * 8, 8, g);
            vo.pack(g);
            vo.store_global(out_row + (size_t)(qq + kRowThreads) * 8);
          };
          if (q + kRowThreads < nvec) {  // the same control flow as sweep 1: the stashed sets are the same ones
            Set a, b;
            int si = 0;
            load_full(a, q, 0);
            while (q + 5 * kRowThreads < nvec) {
              load_full(b, q + 2 * kRowThreads, si + 1);
              emit(a, q);
              load_full(a, q + 4 * kRowThreads, si + 2);
              emit(b, q + 2 * kRowThreads);
              q += 4 * kRowThreads;
              si += 2;
            }
            emit(a, q);
            q += 2 * kRowThreads;
          }
        }
        for (; q < nvec; q += 2 * kRowThreads) {  // the remainder: at most two pairs, possibly partial
          load_pair(q, z0, z1, y0, y1);
          float fz[8], fy[8], g[8];
          z0.unpack(fz);
          if (DENSE) y0.unpack(fy);
          grad8(fz, fy, q * 8, 8, g);
          Vec8<TZ> vo;
          vo.pack(g);
          vo.store_global(out_row + (size_t)q * 8);
          if (q + kRowThreads < nvec) {
            z1.unpack(fz);
            if (DENSE) y1.unpack(fy);
            grad8(fz, fy, (q + kRowThreads) * 8, 8, g);
            vo.pack(g);
            vo.store_global(out_row + (size_t)(q + kRowThreads) * 8);
          }
        }
        for (int i = nvec * 8 + tid; i < V; i += kRowThreads) {
          float fz[8], fy[8], g[8];
          fz[0] = Elem<TZ>::to_f(zrow[i]);
          if (DENSE) fy[0] = Elem<TY>::to_f(yrow[i]);
          grad8(fz, fy, i, 1, g);
          out_row[i] = Elem<TZ>::from_f(g[0]);
        }
      }
      if (!DENSE) {
        // scatter part of the sparse gradient: G[i_k] -= c2 * (sum of p_j with i_j == i_k), exact in fp32 and
        // written once per distinct index, after the sweep's own stores (CTA barrier)
        __syncthreads();
        for (int k = tid; k < p.K; k += kRowThreads) {
          const int idx = sp_i[k];
          if (idx < 0 || idx >= V) continue;
          float ptot = 0.f;
          bool first = true;
          for (int j = 0; j < p.K; ++j) {
            if (sp_i[j] == idx) {
              ptot += sp_p[j];
              if (j < k) first = false;
            }
          }
          if (!first) continue;
          const float zk = Elem<TZ>::to_f(zrow[idx]);
          float gi = c1 * ex2(fmaf(zk, kLog2e, -off1)) + c2 * (ex2(fmaf(zk, c_tau, -offt)) - ptot);
          if (idx == target) gi -= c1;
          out_row[idx] = Elem<TZ>::from_f(gi);
        }
      }
    }
    __syncthreads();  // sh.* and sp_* are rewritten by the next row
  }
  if (tid == 0) {
    float* out = p.partials + (size_t)blockIdx.x * kNumPartialSlots;
    out[0] = (float)sh.acc_ce;
    out[1] = (float)sh.acc_kl;
    out[2] = (float)sh.acc_t;
    out[3] = (float)sh.acc_n;
    out[4] = (float)sh.acc_hits;
    out[5] = out[6] = out[7] = 0.f;
  }
}


// deterministic fixed-order reduction of the per-cluster partial records -> sums[8]
__global__ void kd_reduce_partials_kernel(const float* __restrict__ partials, int n, float* __restrict__ sums) {
  __shared__ double sm[kNumPartialSlots][33];
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;  // 8 warps, one per slot
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) acc += (double)partials[(size_t)i * kNumPartialSlots + slot];
  sm[slot][lane] = acc;
  __syncwarp();
  if (lane == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sm[slot][i];
    sums[slot] = (float)t;
  }
}

int reduce_partials(const float* partials, int n, float* sums, cudaStream_t stream) {
  kd_reduce_partials_kernel<<<1, 32 * kNumPartialSlots, 0, stream>>>(partials, n, sums);
  return check_launch("kd_reduce_partials launch");
}

__global__ void kd_prepare_rows_kernel(const int64_t* __restrict__ labels, const uint8_t* __restrict__ mask, int B,
                                       int T, int64_t ignore_index, int32_t* __restrict__ row_target,
                                       int32_t* __restrict__ n_valid) {
  __shared__ int warp_counts[32];
  int cnt = 0;
  const int n = B * T;
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int b = r / T, t = r - b * T;
    int64_t l = 0;
    const bool ok = row_is_valid(labels, mask, T, b, t, ignore_index, &l);
    // a scored row whose label is not a vocabulary index (negative but not ignore_index, or beyond int32) keeps
    // counting as valid - as in the reference, whose cross_entropy then fails - and gets a target no vocabulary
    // holds: every kernel turns that into a NaN loss instead of a silently wrong one
    const int32_t tgt = (l < 0 || l > 0x7ffffffe) ? 0x7fffffff : (int32_t)l;
    if (row_target) row_target[r] = ok ? tgt : -1;
    cnt += ok ? 1 : 0;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += warp_counts[i];
    *n_valid = t;
  }
}

__global__ void kd_finalize_kernel(const float* __restrict__ sums, float tau, float alpha, int sparse,
                                   float* __restrict__ losses) {
  const float n = sums[3];
  if (n <= 0.f) {  // distillation_loss.py:47-53
    losses[0] = losses[1] = losses[2] = losses[3] = 0.f;
    return;
  }
  const float task = sums[0] / n;
  const float distill = tau * tau * sums[1] / n;
  float teacher;
  if (sparse) {
    teacher = sums[4] > 0.f ? -sums[2] / sums[4] : 0.f;  // :113-118
  } else {
    teacher = sums[2] / n;
  }
  losses[0] = alpha * task + (1.f - alpha) * distill;
  losses[1] = task;
  losses[2] = distill;
  losses[3] = teacher;
}

template <typename T>
__global__ void kd_scale_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ scale) {
  const float s = *scale;
  if (s == 1.0f) return;
  const int64_t n8 = n / 8;
  const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (aligned) {
    for (int64_t i = gid; i < n8; i += stride) {
      Vec8<T> v;
      v.load_global(x + i * 8);
      float f[8];
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= s;
      v.pack(f);
      v.store_global(x + i * 8);
    }
    for (int64_t i = n8 * 8 + gid; i < n; i += stride) x[i] = Elem<T>::from_f(Elem<T>::to_f(x[i]) * s);
  } else {
    for (int64_t i = gid; i < n; i += stride) x[i] = Elem<T>::from_f(Elem<T>::to_f(x[i]) * s);
  }
}

template <typename T>
__global__ void kd_zero_rows_kernel(T* __restrict__ x, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = Elem<T>::from_f(0.f);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
constexpr int kMaxClusters = 1024;

template <typename TZ, typename TY, bool DENSE, bool TAU2, bool GRAD, int kRowThreads>
static int launch_stream_rows_t(const StreamParams& p0, cudaStream_t stream, int ctas_per_sm = 1);

template <typename TZ, typename TY, bool DENSE, bool TAU2, bool GRAD>
static int launch_stream_rows(const StreamParams& p0, cudaStream_t stream) {
  // dense with gradient: 512 threads x 128 registers hold two sets of z and y per thread, one CTA per SM (the stash
  // takes the shared memory); the top-k form has no teacher stream and is faster with 1024 x 64 (measured: 726 vs
  // 862 us at configs[1] shape); forward only (no second sweep, nothing to keep in L2): two 256-thread CTAs per SM
  // overlap one row's block reduction with the other's streaming (measured 551 vs 604 us at the configs[1] shape).
  // (768 and 1024 threads for the dense form measured slower: fewer registers per thread, spills.)
  if constexpr (DENSE && !GRAD) return launch_stream_rows_t<TZ, TY, DENSE, TAU2, GRAD, 256>(p0, stream, 2);
  else if constexpr (DENSE) return launch_stream_rows_t<TZ, TY, DENSE, TAU2, GRAD, 512>(p0, stream);
  else return launch_stream_rows_t<TZ, TY, DENSE, TAU2, GRAD, 1024>(p0, stream);
}

template <typename TZ, typename TY, bool DENSE, bool TAU2, bool GRAD, int kRowThreads>
static int launch_stream_rows_t(const StreamParams& p0, cudaStream_t stream, int ctas_per_sm) {
  StreamParams p = p0;
  auto kern = kd_stream_row_kernel<TZ, TY, DENSE, TAU2, GRAD, kRowThreads>;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device_slot());
  const int n_rows = p.B * p.T;
  static int per_sm_env = -1;
  if (per_sm_env < 0) {
    const char* e = getenv("KD_STREAM_CTAS_PER_SM");
    per_sm_env = e ? atoi(e) : 0;
    if (per_sm_env < 0 || per_sm_env > 4) per_sm_env = 0;
  }
  const int per_sm = per_sm_env > 0 ? per_sm_env : ctas_per_sm;
  int grid = sms * per_sm < n_rows ? sms * per_sm : n_rows;  // ~148 rows in flight keep z + y of a row in L2
  if (grid > kMaxClusters) grid = kMaxClusters;
  // the row counter sits behind the reduced record in the workspace
  int* counter = reinterpret_cast<int*>(p.partials + (size_t)(kMaxClusters + 1) * kNumPartialSlots);
  if (check_cuda(cudaMemsetAsync(counter, 0, sizeof(int), stream), "row counter")) return 1;
  size_t dyn = DENSE ? 0 : (size_t)p.K * 8;
  p.n_stash = 0;
  {
    static int ahead_env = -2;  // KD_STREAM_L2_AHEAD = sets (0 = no L2 prefetch)
    if (ahead_env == -2) {
      const char* e = getenv("KD_STREAM_L2_AHEAD");
      ahead_env = e ? atoi(e) : 3;
      if (ahead_env < 0 || ahead_env > 16) ahead_env = 3;
    }
    p.l2_ahead = p.vec_ok ? ahead_env : 0;
  }
  if (DENSE && GRAD && p.vec_ok) {
    // shared-memory stash of the leading sets of every row (see the kernel): as many as fit beside the static part;
    // KD_STREAM_STASH = n caps it (0 = every sweep-2 read comes from L2, the round-2 form; -1 = as many as fit)
    static int stash_env = -2;
    if (stash_env == -2) {
      const char* e = getenv("KD_STREAM_STASH");
      stash_env = e ? atoi(e) : 4;  // measured at the configs[1] shape: 4 sets 1021 us, 7 sets (all that fit) 1042, none 1095
    }
    static int max_dyn[kMaxDevices] = {};  // per instantiation and device
    const int slot = current_device_slot();
    if (max_dyn[slot] == 0) {
      cudaFuncAttributes fa;
      int optin = 0;
      if (check_cuda(cudaFuncGetAttributes(&fa, kern), "kd_stream_row attributes")) return 1;
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, slot);
      int avail = (per_sm > 1 ? (optin + 1024) / per_sm - 1024 : optin) - (int)fa.sharedSizeBytes;
      if (avail < 0) avail = 0;
      if (avail > 48 * 1024 &&
          check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, avail), "kd_stream_row smem"))
        return 1;
      max_dyn[slot] = avail > 0 ? avail : -1;
    }
    const int set_bytes = kRowThreads * 2 * (int)(sizeof(Vec8<TZ>) + sizeof(Vec8<TY>));
    int n = max_dyn[slot] > 0 ? max_dyn[slot] / set_bytes : 0;
    const int whole_sets = (p.V / 8) / (2 * kRowThreads);
    if (n > whole_sets) n = whole_sets;
    if (stash_env >= 0 && n > stash_env) n = stash_env;
    p.n_stash = n;
    dyn = (size_t)n * set_bytes;
  }
  kern<<<grid, kRowThreads, dyn, stream>>>(p, counter);
  if (check_launch("kd_stream_row launch")) return 1;
  return reduce_partials(p.partials, grid, p0.partials + (size_t)kMaxClusters * kNumPartialSlots, stream);
}

template <typename TZ, typename TY, bool DENSE, bool TAU2, bool GRAD>
static int launch_stream(const StreamParams& p0, cudaStream_t stream) {
  return launch_stream_rows<TZ, TY, DENSE, TAU2, GRAD>(p0, stream);
}

template <typename TZ, typename TY, bool DENSE>
static int dispatch_flags(const StreamParams& p, cudaStream_t s) {
  const bool tau2 = p.tau == 2.0f;
  const bool grad = p.dlogits != nullptr;
  if (tau2) return grad ? launch_stream<TZ, TY, DENSE, true, true>(p, s) : launch_stream<TZ, TY, DENSE, true, false>(p, s);
  return grad ? launch_stream<TZ, TY, DENSE, false, true>(p, s) : launch_stream<TZ, TY, DENSE, false, false>(p, s);
}

template <typename TZ>
static int dispatch_y(const StreamParams& p, int y_dtype, cudaStream_t s) {
  switch (y_dtype) {
    case KD_DTYPE_F32: return dispatch_flags<TZ, float, true>(p, s);
    case KD_DTYPE_BF16: return dispatch_flags<TZ, __nv_bfloat16, true>(p, s);
    case KD_DTYPE_F16: return dispatch_flags<TZ, __half, true>(p, s);
  }
  set_error("kd_dense_fwd_bwd: unsupported teacher dtype code %d", y_dtype);
  return 1;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static size_t dtype_size(int code) { return code == KD_DTYPE_F32 ? 4 : 2; }

static int stream_common(StreamParams& p, int z_dtype, int y_dtype, bool dense, float* sums, void* workspace,
                         size_t workspace_bytes, cudaStream_t s) {
  if (p.B <= 0 || p.T <= 0 || p.V <= 0) {
    set_error("kd_stream: empty shape B=%d T=%d V=%d", p.B, p.T, p.V);
    return 1;
  }
  if (!(p.tau > 0.f)) {
    set_error("kd_stream: temperature must be > 0");
    return 1;
  }
  if (workspace == nullptr || workspace_bytes < kd_stream_workspace_bytes() || !aligned16(workspace)) {
    set_error("kd_stream: workspace must be 16-byte aligned and >= %zu bytes", kd_stream_workspace_bytes());
    return 1;
  }
  p.partials = reinterpret_cast<float*>(workspace);
  const size_t zs = dtype_size(z_dtype), ys = dtype_size(y_dtype);
  bool ok = aligned16(p.z) && (p.z_sb * zs) % 16 == 0 && (p.z_st * zs) % 16 == 0;
  if (dense) ok = ok && aligned16(p.y) && (p.y_sb * ys) % 16 == 0 && (p.y_st * ys) % 16 == 0;
  if (p.dlogits) ok = ok && aligned16(p.dlogits) && ((size_t)p.V * zs) % 16 == 0;
  p.vec_ok = ok ? 1 : 0;
  int rc;
  if (dense) {
    switch (z_dtype) {
      case KD_DTYPE_F32: rc = dispatch_y<float>(p, y_dtype, s); break;
      case KD_DTYPE_BF16: rc = dispatch_y<__nv_bfloat16>(p, y_dtype, s); break;
      case KD_DTYPE_F16: rc = dispatch_y<__half>(p, y_dtype, s); break;
      default: set_error("kd_stream: unsupported student dtype code %d", z_dtype); return 1;
    }
  } else {
    switch (z_dtype) {
      case KD_DTYPE_F32: rc = dispatch_flags<float, float, false>(p, s); break;
      case KD_DTYPE_BF16: rc = dispatch_flags<__nv_bfloat16, __nv_bfloat16, false>(p, s); break;
      case KD_DTYPE_F16: rc = dispatch_flags<__half, __half, false>(p, s); break;
      default: set_error("kd_stream: unsupported student dtype code %d", z_dtype); return 1;
    }
  }
  if (rc) return rc;
  // reduced record sits right after the per-cluster partials; copy it out
  return check_cuda(cudaMemcpyAsync(sums, p.partials + (size_t)kMaxClusters * kNumPartialSlots,
                                    kNumPartialSlots * sizeof(float), cudaMemcpyDeviceToDevice, s),
                    "sums copy");
}

}  // namespace kd

using namespace kd;

extern "C" size_t kd_stream_workspace_bytes(void) {
  return (size_t)(kMaxClusters + 1) * kNumPartialSlots * sizeof(float) + 16 /* row counter */;
}

extern "C" int kd_prepare_rows(const int64_t* labels, const uint8_t* mask, int B, int T, int64_t ignore_index,
                               int32_t* row_target, int32_t* n_valid, void* stream) {
  DeviceGuard device_guard(labels);
  if (labels == nullptr || n_valid == nullptr || B <= 0 || T <= 0) {
    set_error("kd_prepare_rows: bad arguments");
    return 1;
  }
  kd_prepare_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(labels, mask, B, T, ignore_index, row_target, n_valid);
  return check_launch("kd_prepare_rows launch");
}

extern "C" int kd_finalize_losses(const float* sums, float tau, float alpha, int sparse, float* losses, void* stream) {
  DeviceGuard device_guard(sums);
  if (!sums || !losses) {
    set_error("kd_finalize_losses: null pointer");
    return 1;
  }
  kd_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, tau, alpha, sparse, losses);
  return check_launch("kd_finalize launch");
}

extern "C" int kd_dense_fwd_bwd(const void* z, int z_dtype, int64_t z_stride_b, int64_t z_stride_t, const void* y,
                                int y_dtype, int64_t y_stride_b, int64_t y_stride_t, const int32_t* row_target, int B,
                                int T, int V, float tau, float alpha, const int32_t* n_norm, float grad_scale,
                                float* sums, void* dlogits, void* workspace, size_t workspace_bytes, void* stream) {
  DeviceGuard device_guard(z);
  if (!z || !y || !row_target || !sums || (dlogits && !n_norm)) {
    set_error("kd_dense_fwd_bwd: null pointer argument");
    return 1;
  }
  StreamParams p = {};
  p.z = z; p.y = y;
  p.z_sb = z_stride_b; p.z_st = z_stride_t; p.y_sb = y_stride_b; p.y_st = y_stride_t;
  p.row_target = row_target;
  p.B = B; p.T = T; p.V = V;
  p.tau = tau; p.alpha = alpha; p.grad_scale = grad_scale;
  p.n_norm = n_norm; p.dlogits = dlogits;
  return stream_common(p, z_dtype, y_dtype, true, sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int kd_sparse_fwd_bwd(const void* z, int z_dtype, int64_t z_stride_b, int64_t z_stride_t,
                                 const float* topk_v, const int32_t* topk_i, int K, const int32_t* row_target, int B,
                                 int T, int V, float tau, float alpha, const int32_t* n_norm, float grad_scale,
                                 float* sums, void* dlogits, void* workspace, size_t workspace_bytes, void* stream) {
  DeviceGuard device_guard(z);
  if (!z || !topk_v || !topk_i || !row_target || !sums || (dlogits && !n_norm)) {
    set_error("kd_sparse_fwd_bwd: null pointer argument");
    return 1;
  }
  if (K <= 0 || K > kMaxTopK) {
    set_error("kd_sparse_fwd_bwd: K=%d outside [1, %d]", K, kMaxTopK);
    return 1;
  }
  StreamParams p = {};
  p.z = z;
  p.z_sb = z_stride_b; p.z_st = z_stride_t;
  p.topk_v = topk_v; p.topk_i = topk_i; p.K = K;
  p.row_target = row_target;
  p.B = B; p.T = T; p.V = V;
  p.tau = tau; p.alpha = alpha; p.grad_scale = grad_scale;
  p.n_norm = n_norm; p.dlogits = dlogits;
  return stream_common(p, z_dtype, z_dtype, false, sums, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int kd_scale_inplace(void* x, int dtype, int64_t n, const float* scale, void* stream) {
  DeviceGuard device_guard(x);
  if (!x || !scale || n < 0) {
    set_error("kd_scale_inplace: bad arguments");
    return 1;
  }
  if (n == 0) return 0;
  const int threads = 256;
  int64_t want = (n / 8 + threads - 1) / threads;
  const int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case KD_DTYPE_F32: kd_scale_kernel<float><<<blocks, threads, 0, s>>>((float*)x, n, scale); break;
    case KD_DTYPE_BF16: kd_scale_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>((__nv_bfloat16*)x, n, scale); break;
    case KD_DTYPE_F16: kd_scale_kernel<__half><<<blocks, threads, 0, s>>>((__half*)x, n, scale); break;
    default: set_error("kd_scale_inplace: unsupported dtype code %d", dtype); return 1;
  }
  return check_launch("kd_scale launch");
}

extern "C" int kd_mask_rows(void* grad, int dtype, int64_t old_vocab, int64_t H, void* stream) {
  DeviceGuard device_guard(grad);
  if (!grad || old_vocab < 0 || H <= 0) {
    set_error("kd_mask_rows: bad arguments");
    return 1;
  }
  const int64_t n = old_vocab * H;
  if (n == 0) return 0;
  // a memset is the fastest zero-fill the device offers (all supported dtypes have all-zero-bits 0.0)
  return check_cuda(cudaMemsetAsync(grad, 0, (size_t)n * dtype_size(dtype), (cudaStream_t)stream), "kd_mask_rows memset");
}
