// K3: teacher top-k log-prob compaction (HBM-bound, one read of the logits).
//
// Replaces the three torch calls of extract_teacher_logits.py:114-129 / train.py:82-91:
//   log_softmax(logits) -> topk(k) -> values fp16, indices int32.
// One CTA per row:
//   pass 1 (HBM): online log-sum-exp + per-thread running maximum;
//   threshold   : the k-th largest of the 512 per-thread maxima is a lower bound of the k-th
//                 largest logit (at least k elements are >= it);
//   pass 2 (L2) : every logit >= threshold is appended to a shared-memory candidate list
//                 (typically 1-2 k entries for k = 64..128 at V = 153k);
//   select      : bitonic sort of the candidates by (logit desc, index asc), first k win;
//   values      : round_to_input_dtype((x - max) - log(sum)) -> fp16.
// Rounding is monotone, so this is a valid top-k of the rounded log-probs under a fixed
// tie-break and equals torch.topk(log_softmax(x)) index-for-index on tie-free rows
// (hard part 4).  Rows whose candidate list overflows (massive ties, constant
// rows) take an exact but slow bisection path.
#include <cstdlib>

#include "kd_common.cuh"

namespace kd {

constexpr int kTopkMaxThreads = 512;  // thread counts: 128 (k <= 64), 256 (k <= 128), 512 (k <= 512)
constexpr int kTopkCap = 2048;

__device__ __forceinline__ uint32_t order_key(float x) {
  x = x + 0.0f;  // -0.0 -> +0.0 so that equal values get equal keys
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

struct TopkShared {
  uint64_t cand[kTopkCap];
  uint64_t xch[2 * kTopkMaxThreads];  // cross-warp exchange of block_sort_desc (u32 keys use the same storage)
  uint32_t thr;
  float red_m[kTopkMaxThreads / 32];
  float red_s[kTopkMaxThreads / 32];
  int count;
  int cnt_a, cnt_b;
  float lse_m, lse_log;
};

// descending bitonic sort of n (power of two) 64-bit keys in shared memory
__device__ void bitonic_desc_u64(uint64_t* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool desc = (i & size) == 0;
        const uint64_t x = a[i], y = a[j];
        if ((x < y) == desc) {
          a[i] = y;
          a[j] = x;
        }
      }
    }
  }
  __syncthreads();
}
__device__ void bitonic_desc_u32(uint32_t* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
        const int i = 2 * t - (t & (stride - 1));
        const int j = i + stride;
        const bool desc = (i & size) == 0;
        const uint32_t x = a[i], y = a[j];
        if ((x < y) == desc) {
          a[i] = y;
          a[j] = x;
        }
      }
    }
  }
  __syncthreads();
}

constexpr int kTopkUnroll = 4;  // independent 16-byte loads in flight per thread (the loops are latency-bound)

// Descending bitonic sort of one key per thread over the whole CTA (NT keys): afterwards thread i holds
// the i-th largest.  Exchanges inside a warp are shuffles (no barrier); the 10 stages whose partner sits in
// another warp go through a double-buffered shared-memory array with one barrier each (the all-smem version
// needs one barrier for each of its 45 stages).
__device__ __forceinline__ uint32_t shfl_xor_key(uint32_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ uint64_t shfl_xor_key(uint64_t v, int m) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
  const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((uint64_t)hi << 32) | lo;
}
template <typename KeyT, int NT>
__device__ __forceinline__ KeyT block_sort_desc(KeyT key, KeyT* xch /* [2][NT] */) {
  const int tid = threadIdx.x;
  int buf = 0;
#pragma unroll 1
  for (int size = 2; size <= NT; size <<= 1) {
    const bool desc = (tid & size) == 0;
#pragma unroll 1
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      KeyT other;
      if (stride < 32) {
        other = shfl_xor_key(key, stride);
      } else {
        xch[buf * NT + tid] = key;
        __syncthreads();
        other = xch[buf * NT + (tid ^ stride)];
        buf ^= 1;  // the next cross-warp stage writes the other buffer: no second barrier needed
      }
      const bool lower = (tid & stride) == 0;            // this thread keeps the "first" element of the pair
      const bool take_max = lower == desc;
      const KeyT mx = key > other ? key : other, mn = key > other ? other : key;
      key = take_max ? mx : mn;
    }
  }
  return key;
}

template <typename T, int NT, typename F>
__device__ __forceinline__ void for_each_elem(const T* __restrict__ row, int V, bool vec_ok, F&& fn) {
  const int tid = threadIdx.x;
  const int vhi = vec_ok ? (V & ~7) : 0;
  const uint64_t pol_drop = l2_policy_evict_first();  // second (and later) reads: the lines are dead afterwards
  constexpr int kStep = NT * 8;
  int i = tid * 8;
  for (; i + (kTopkUnroll - 1) * kStep < vhi; i += kTopkUnroll * kStep) {
    Vec8<T> v[kTopkUnroll];
#pragma unroll
    for (int u = 0; u < kTopkUnroll; ++u) v[u].load_global_hint(row + i + u * kStep, pol_drop);
#pragma unroll
    for (int u = 0; u < kTopkUnroll; ++u) {
      float f[8];
      v[u].unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) fn(f[j], i + u * kStep + j);
    }
  }
  for (; i < vhi; i += kStep) {
    Vec8<T> v;
    v.load_global_hint(row + i, pol_drop);
    float f[8];
    v.unpack(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) fn(f[j], i + j);
  }
  for (int i2 = vhi + tid; i2 < V; i2 += NT) fn(Elem<T>::to_f(row[i2]), i2);
}

// any element of 8 packed 16-bit values >= thr (NaN counts as >=)?  Two elements per HSETP2; fp32 rows compare
// element-wise.  Only vectors that pass are unpacked.
template <typename T>
__device__ __forceinline__ bool any_ge(const Vec8<T>& v, float thr);
template <>
__device__ __forceinline__ bool any_ge<__nv_bfloat16>(const Vec8<__nv_bfloat16>& v, float thr) {
  const __nv_bfloat162 t2 = __float2bfloat162_rn(thr);
  const uint32_t w[4] = {v.a.x, v.a.y, v.a.z, v.a.w};
  bool any = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) any |= !__hblt2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]), t2);
  return any;
}
template <>
__device__ __forceinline__ bool any_ge<__half>(const Vec8<__half>& v, float thr) {
  const __half2 t2 = __float2half2_rn(thr);
  const uint32_t w[4] = {v.a.x, v.a.y, v.a.z, v.a.w};
  bool any = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) any |= !__hblt2(*reinterpret_cast<const __half2*>(&w[i]), t2);
  return any;
}
template <>
__device__ __forceinline__ bool any_ge<float>(const Vec8<float>& v, float thr) {
  float f[8];
  v.unpack(f);
  bool any = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) any |= !(f[j] < thr);
  return any;
}

// pass 2: append every element with key >= t0 to the shared-memory candidate list
template <typename T, int NT>
__device__ __forceinline__ void collect_candidates(const T* __restrict__ row, int V, bool vec_ok, uint32_t t0,
                                                   TopkShared& sh) {
  const int tid = threadIdx.x;
  const int vhi = vec_ok ? (V & ~7) : 0;
  const float t0f = key_to_float(t0);  // exactly representable in T: it is one of the row's elements
  const uint64_t pol_drop = l2_policy_evict_first();
  auto push = [&](float x, int idx) {
    const uint32_t key = order_key(x);
    if (key >= t0) {
      const int slot = atomicAdd(&sh.count, 1);
      if (slot < kTopkCap) sh.cand[slot] = ((uint64_t)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)idx);
    }
  };
  auto visit = [&](const Vec8<T>& v, int base) {
    if (any_ge<T>(v, t0f)) {  // rare: ~k of the V / 8 vectors of a row
      float f[8];
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) push(f[j], base + j);
    }
  };
  constexpr int kStep = NT * 8;
  int i = tid * 8;
  for (; i + (kTopkUnroll - 1) * kStep < vhi; i += kTopkUnroll * kStep) {
    Vec8<T> v[kTopkUnroll];
#pragma unroll
    for (int u = 0; u < kTopkUnroll; ++u) v[u].load_global_hint(row + i + u * kStep, pol_drop);
#pragma unroll
    for (int u = 0; u < kTopkUnroll; ++u) visit(v[u], i + u * kStep);
  }
  for (; i < vhi; i += kStep) {
    Vec8<T> v;
    v.load_global_hint(row + i, pol_drop);
    visit(v, i);
  }
  for (int i2 = vhi + tid; i2 < V; i2 += NT) push(Elem<T>::to_f(row[i2]), i2);
}

// block-wide count of elements satisfying pred (two alternating counters avoid a reset barrier)
template <typename T, int NT, typename P>
__device__ int block_count(const T* row, int V, bool vec_ok, TopkShared& sh, P&& pred) {
  int c = 0;
  for_each_elem<T, NT>(row, V, vec_ok, [&](float x, int idx) { c += pred(order_key(x), idx) ? 1 : 0; });
  c = __reduce_add_sync(0xffffffffu, c);
  __syncthreads();
  if (threadIdx.x == 0) sh.cnt_a = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh.cnt_a, c);
  __syncthreads();
  return sh.cnt_a;
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) kd_topk_kernel(const T* __restrict__ logits, int64_t R, int V,
                                                               int64_t row_stride, int k, __half* __restrict__ out_v,
                                                               int32_t* __restrict__ out_i, int vec_ok_i) {
  __shared__ TopkShared sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec_ok = vec_ok_i != 0;

  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    const T* row = logits + r * row_stride;
    // ---- pass 1: online LSE, thread maximum ------------------------------------------------
    float m = -CUDART_INF_F, s = 0.f;
    {
      const int vhi = vec_ok ? (V & ~7) : 0;
      const uint64_t pol_keep = l2_policy_evict_last();  // the row is read again from L2 in pass 2
      constexpr int kStep = NT * 8;
      auto update8 = [&](const Vec8<T>& v) {
        float f[8];
        v.unpack(f);
        float vm = f[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) vm = fmaxf(vm, f[j]);
        if (vm > m) {
          s *= exp_diff(m, vm, kLog2e);
          m = vm;
        }
        if (m != -CUDART_INF_F) {
          const float off = m * kLog2e;
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            p0 += ex2(fmaf(f[j], kLog2e, -off));
            p1 += ex2(fmaf(f[j + 1], kLog2e, -off));
          }
          s += p0 + p1;
        }
      };
      int i = tid * 8;
      for (; i + (kTopkUnroll - 1) * kStep < vhi; i += kTopkUnroll * kStep) {
        // 32 elements per online update: one maximum / rescale decision for all four vectors, then the
        // exponentials back to back (the 8-element form spent a third of its instructions on the branches)
        Vec8<T> v[kTopkUnroll];
#pragma unroll
        for (int u = 0; u < kTopkUnroll; ++u) v[u].load_global_hint(row + i + u * kStep, pol_keep);
        float f[kTopkUnroll][8];
        float vm = -CUDART_INF_F;
#pragma unroll
        for (int u = 0; u < kTopkUnroll; ++u) {
          v[u].unpack(f[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) vm = fmaxf(vm, f[u][j]);
        }
        if (vm > m) {
          s *= exp_diff(m, vm, kLog2e);
          m = vm;
        }
        if (m != -CUDART_INF_F) {
          const float off = m * kLog2e;
          float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
          for (int u = 0; u < kTopkUnroll; ++u) {
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              p0 += ex2(fmaf(f[u][j], kLog2e, -off));
              p1 += ex2(fmaf(f[u][j + 1], kLog2e, -off));
              p2 += ex2(fmaf(f[u][j + 2], kLog2e, -off));
              p3 += ex2(fmaf(f[u][j + 3], kLog2e, -off));
            }
          }
          s += (p0 + p1) + (p2 + p3);
        }
      }
      for (; i < vhi; i += kStep) {
        Vec8<T> v;
        v.load_global_hint(row + i, pol_keep);
        update8(v);
      }
      for (int i = vhi + tid; i < V; i += NT) {
        const float x = Elem<T>::to_f(row[i]);
        if (x > m) {
          s *= exp_diff(m, x, kLog2e);
          m = x;
        }
        if (m != -CUDART_INF_F) s += ex2((x - m) * kLog2e);
      }
    }
    // block LSE
    float wm = warp_max(m);
    float ws = warp_sum(s * exp_diff(m, wm, kLog2e));
    if (lane == 0) {
      sh.red_m[warp] = wm;
      sh.red_s[warp] = ws;
    }
    if (tid == 0) sh.count = 0;
    __syncthreads();
    if (warp == 0) {
      float a = lane < NT / 32 ? sh.red_m[lane] : -CUDART_INF_F;
      float b = lane < NT / 32 ? sh.red_s[lane] : 0.f;
      const float mm = warp_max(a);
      b = warp_sum(b * exp_diff(a, mm, kLog2e));
      if (lane == 0) {
        sh.lse_m = mm;
        sh.lse_log = ln_acc(b);
      }
    }
    // ---- threshold: k-th largest thread maximum ---------------------------------------------
    {
      const uint32_t sorted = block_sort_desc<uint32_t, NT>(order_key(m), reinterpret_cast<uint32_t*>(sh.xch));
      if (tid == k - 1) sh.thr = sorted;
    }
    __syncthreads();
    const uint32_t t0 = sh.thr;
    // ---- pass 2: collect candidates ----------------------------------------------------------
    collect_candidates<T, NT>(row, V, vec_ok, t0, sh);
    __syncthreads();
    int count = sh.count;
    if (count > kTopkCap) {
      // ---- exact slow path: bisection on the key, then on the index among ties ---------------
      uint32_t lo = t0, hi = 0xffffffffu;  // invariant: count(key >= lo) >= k
      while (lo < hi) {
        const uint32_t mid = lo + (uint32_t)(((uint64_t)hi - lo + 1) >> 1);
        const int c = block_count<T, NT>(row, V, vec_ok, sh, [&](uint32_t key, int) { return key >= mid; });
        if (c >= k) lo = mid; else hi = mid - 1;
      }
      const uint32_t kth = lo;
      const int above = block_count<T, NT>(row, V, vec_ok, sh, [&](uint32_t key, int) { return key > kth; });
      const int need = k - above;  // >= 1 ties at kth to take, smallest indices first
      int jl = 0, jh = V - 1;      // smallest J with count(key == kth && idx <= J) >= need
      while (jl < jh) {
        const int mid = jl + ((jh - jl) >> 1);
        const int c = block_count<T, NT>(row, V, vec_ok, sh, [&](uint32_t key, int idx) { return key == kth && idx <= mid; });
        if (c >= need) jh = mid; else jl = mid + 1;
      }
      const int jmax = jl;
      __syncthreads();
      if (tid == 0) sh.count = 0;
      __syncthreads();
      for_each_elem<T, NT>(row, V, vec_ok, [&](float x, int idx) {
        const uint32_t key = order_key(x);
        if (key > kth || (key == kth && idx <= jmax)) {
          const int slot = atomicAdd(&sh.count, 1);
          if (slot < kTopkCap) sh.cand[slot] = ((uint64_t)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)idx);
        }
      });
      __syncthreads();
      count = sh.count;  // == k
    }
    // ---- order the candidates, emit the first k ----------------------------------------------
    const float lm = sh.lse_m, ll = sh.lse_log;
    auto emit = [&](uint64_t c, int j) {
      const float x = key_to_float((uint32_t)(c >> 32));
      const int idx = (int)(0xffffffffu - (uint32_t)(c & 0xffffffffu));
      const float lp = (x - lm) - ll;
      const float lp_r = Elem<T>::to_f(Elem<T>::from_f(lp));  // log_softmax returns the logits' dtype
      out_v[r * k + j] = __float2half_rn(lp_r);
      out_i[r * k + j] = idx;
    };
    if (count <= NT) {  // typical: one candidate per thread, sorted in registers
      const uint64_t mine = tid < count ? sh.cand[tid] : 0ull;  // pads sort last
      const uint64_t sorted = block_sort_desc<uint64_t, NT>(mine, sh.xch);
      if (tid < k) emit(sorted, tid);
    } else {
      int n = 32;
      while (n < count) n <<= 1;
      for (int i = count + tid; i < n; i += NT) sh.cand[i] = 0ull;
      bitonic_desc_u64(sh.cand, n);
      for (int j = tid; j < k; j += NT) emit(sh.cand[j], j);
    }
    __syncthreads();
  }
}

}  // namespace kd

using namespace kd;

extern "C" int kd_topk_logprobs(const void* logits, int dtype, int64_t R, int V, int64_t row_stride, int k, void* out_v,
                                int32_t* out_i, void* stream) {
  kd::DeviceGuard device_guard(logits);
  if (!logits || !out_v || !out_i) {
    set_error("kd_topk_logprobs: null pointer argument");
    return 1;
  }
  if (R < 0 || V <= 0 || k <= 0 || k > V || k > kTopkMaxThreads) {
    set_error("kd_topk_logprobs: need 1 <= k <= min(V, %d); got k=%d V=%d R=%lld", kTopkMaxThreads, k, V, (long long)R);
    return 1;
  }
  if (R == 0) return 0;
  const size_t es = dtype == KD_DTYPE_F32 ? 4 : 2;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (row_stride * es) % 16 == 0) ? 1 : 0;
  // persistent: a few CTAs per SM loop over the rows, so that the rows in flight (306 KB each for bf16 at
  // V = 152,936) stay L2-resident between the two passes; KD_TOPK_CTAS_PER_SM overrides the default of 4
  // Threads per row: the fewest that still give 2k thread maxima to take the threshold from.  Fewer threads per
  // CTA = more CTAs (rows) per SM (64 registers per thread): while some rows sit in their sorts, enough others
  // stream - with 512-thread CTAs only two rows fit an SM and the HBM idles during their sort phases (0.41 of the
  // peak); KD_TOPK_THREADS / KD_TOPK_CTAS_PER_SM override.
  static int nt_env = -1, per_sm_env = -1;
  if (nt_env < 0) {
    const char* e = getenv("KD_TOPK_THREADS");
    nt_env = e ? atoi(e) : 0;
    const char* c = getenv("KD_TOPK_CTAS_PER_SM");
    per_sm_env = c ? atoi(c) : 0;
  }
  int nt = k <= 128 ? 256 : 512;  // measured at k = 64: 256 x 4 CTAs/SM 780 us, 128 x 8 958 us, 512 x 2 943 us
  if ((nt_env == 128 || nt_env == 256 || nt_env == 512) && nt_env >= nt) nt = nt_env;
  int per_sm = per_sm_env > 0 ? per_sm_env : (nt == 128 ? 8 : (nt == 256 ? 5 : 2));  // what the register file holds
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = (int)(R < (int64_t)sms * per_sm ? R : (int64_t)sms * per_sm);
  cudaStream_t s = (cudaStream_t)stream;
#define KD_TOPK_LAUNCH(TYPE, NT)                                                                                   \
  kd_topk_kernel<TYPE, NT><<<grid, NT, 0, s>>>((const TYPE*)logits, R, V, row_stride, k, (__half*)out_v, out_i, vec_ok)
#define KD_TOPK_DISPATCH(TYPE)          \
  do {                                  \
    if (nt == 128) {                    \
      KD_TOPK_LAUNCH(TYPE, 128);        \
    } else if (nt == 256) {             \
      KD_TOPK_LAUNCH(TYPE, 256);        \
    } else {                            \
      KD_TOPK_LAUNCH(TYPE, 512);        \
    }                                   \
  } while (0)
  switch (dtype) {
    case KD_DTYPE_F32: KD_TOPK_DISPATCH(float); break;
    case KD_DTYPE_BF16: KD_TOPK_DISPATCH(__nv_bfloat16); break;
    case KD_DTYPE_F16: KD_TOPK_DISPATCH(__half); break;
    default:
      set_error("kd_topk_logprobs: unsupported dtype code %d", dtype);
      return 1;
  }
#undef KD_TOPK_DISPATCH
#undef KD_TOPK_LAUNCH
  return check_launch("kd_topk launch");
}
