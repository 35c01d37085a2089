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
#include <cstdint>
#include <cstdlib>
#include <mutex>

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

// Everything after the candidate collection, shared by kd_topk_kernel and kd_head_select_kernel: the exact slow
// path for rows whose candidate list overflowed, the ordering of the candidates and the k output entries.
template <typename T, int NT>
__device__ __forceinline__ void select_and_emit(const T* __restrict__ row, int V, bool vec_ok, int k, uint32_t t0,
                                                TopkShared& sh, int64_t r, __half* __restrict__ out_v,
                                                int32_t* __restrict__ out_i) {
  const int tid = threadIdx.x;
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
    select_and_emit<T, NT>(row, V, vec_ok, k, t0, sh, r, out_v, out_i);
    __syncthreads();
  }
}

// =====================================================================================================================
// Warp-per-row form (k <= 128, 16-byte aligned rows): the default for kd_topk_logprobs and the selection behind the
// teacher head GEMM.  The CTA-per-row kernel above keeps a row's sorts, barriers and candidate list inside one CTA, so
// only ~5 rows per SM are in flight and HBM idles while they sit in their select phases (0.5 of the HBM peak).  Here
// every warp owns a row and nothing in a row's life needs a CTA barrier:
//   pass 1   : one sweep over the row (HBM): online log-sum-exp per lane and the maximum of every 32-element piece,
//              kept as a 16-bit order key in shared memory (9.6 KB per row at V = 152,936);
//   threshold: the exact k-th largest piece key (bisection with packed 16-bit compares, no atomics) - at least k pieces,
//              hence k elements, reach it;
//   collect  : only the pieces whose key reaches the threshold are read again (about k pieces of 64 bytes instead of
//              the row), their elements >= threshold go to the warp's candidate list;
//   emit     : rank of a candidate = number of larger candidates (keys are unique: value, then lower index first);
//              ranks < k are written with the same value rounding as above.
// Rows whose candidate list overflows (massive ties) are solved by the same warp with the bisection of the CTA kernel,
// restricted to the qualifying pieces.  16 row-warps per SM stream concurrently while others select.
constexpr int kWrCap = 256;     // candidates / qualifying pieces per row (k <= 128)
constexpr int kWrWarps = 8;     // rows per CTA
constexpr int kWrPiece = 32;    // elements per piece behind the head GEMM and in the two-kernel form (its epilogue produces them)
constexpr int kWarpPiece = 64;  // alternative piece size of the warp-per-row kernel (-DKD_TOPK_WARP_PIECE=kWarpPiece): half the
                                // keys per row, 24 row-warps per SM instead of 16 - measured slower, see kd_topk_warp_kernel

// Piece maxima live in shared memory as bf16 bit patterns rounded DOWN (exact for bf16 rows), so that counting
// "pieces >= t" is one packed hardware compare (HSET2.BF16) + one packed add per two pieces.
__device__ __forceinline__ uint16_t bf16_bits_rd(float x) {
  const __nv_bfloat16 b = __float2bfloat16_rd(x);
  return *reinterpret_cast<const uint16_t*>(&b);
}
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
// piece maximum -> stored key: bf16 rows hold bf16 values, so the key is the float's upper half (no conversion)
template <typename T>
__device__ __forceinline__ uint16_t piece_key(float x) { return bf16_bits_rd(x); }
template <>
__device__ __forceinline__ uint16_t piece_key<__nv_bfloat16>(float x) { return (uint16_t)(__float_as_uint(x) >> 16); }
// 16-bit order key <-> bf16 bit pattern (same map as order_key on the top half of a float)
__device__ __forceinline__ uint32_t key16_to_bf16_bits(uint32_t k) { return (k & 0x8000u) ? (k & 0x7fffu) : (~k & 0xffffu); }
constexpr uint16_t kBf16NegInf = 0xFF80;

struct WarpRow {
  uint64_t* cand;    // [kWrCap]
  uint16_t* plist;   // [kWrCap] qualifying pieces
  int* count;        // [4]
  uint16_t* pv;      // [n_pieces_pad] piece maxima (bf16 bits), n_pieces_pad a multiple of 8, padding = -inf
};
__host__ __device__ inline size_t warp_row_bytes(int n_pieces_pad) {
  return (size_t)kWrCap * 8 + (size_t)kWrCap * 2 + 16 + (((size_t)n_pieces_pad * 2 + 15) & ~(size_t)15);
}
__device__ __forceinline__ WarpRow warp_row_smem(unsigned char* base, int warp, int n_pieces_pad) {
  unsigned char* p = base + (size_t)warp * warp_row_bytes(n_pieces_pad);
  WarpRow w;
  w.cand = reinterpret_cast<uint64_t*>(p);
  w.plist = reinterpret_cast<uint16_t*>(p + kWrCap * 8);
  w.count = reinterpret_cast<int*>(p + kWrCap * 10);
  w.pv = reinterpret_cast<uint16_t*>(p + kWrCap * 10 + 16);
  return w;
}

// per-halfword (a >= b) as bf16 1.0 / 0.0
__device__ __forceinline__ __nv_bfloat162 ge2(uint32_t a, __nv_bfloat162 b) {
  return __hge2(*reinterpret_cast<const __nv_bfloat162*>(&a), b);
}

// The k-th largest piece maximum (bf16 bits; -inf when there are fewer than k pieces: everything qualifies).
// Bisection over the 16-bit order-key space; a round counts the pieces >= mid with packed compares over the 16-byte
// padded array: ~n / 256 LDS.128 and n / 64 HSET2 + HADD2 per lane, no atomics.  (A shared-memory histogram serialises
// on the one or two exponent bins nearly all piece maxima of a row share: measured 160 us per row.)
__device__ __forceinline__ uint32_t warp_kth_largest_piece(const uint16_t* pv, int n_pad8, int n, int k, int lane) {
  if (n < k) return kBf16NegInf;
  const uint4* kv = reinterpret_cast<const uint4*>(pv);
  const int nv = n_pad8 >> 3;
  uint32_t lo = 0x007fu /* key of -inf */, hi = 0xff7fu /* key of +max finite */;  // invariant: count(>= lo) >= k
#pragma unroll 1
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
    const uint32_t mb = key16_to_bf16_bits(mid) * 0x00010001u;
    const __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&mb);
    __nv_bfloat162 acc0 = __float2bfloat162_rn(0.f), acc1 = acc0;  // <= 2 * ceil(nv / 32) <= 256 per halfword: exact
    for (int v = lane; v < nv; v += 32) {
      const uint4 x = kv[v];
      acc0 = __hadd2(acc0, __hadd2(ge2(x.x, m2), ge2(x.y, m2)));
      acc1 = __hadd2(acc1, __hadd2(ge2(x.z, m2), ge2(x.w, m2)));
    }
    const float2 f0 = __bfloat1622float2(acc0), f1 = __bfloat1622float2(acc1);
    int c = (int)((f0.x + f0.y) + (f1.x + f1.y));
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) lo = mid; else hi = mid - 1u;
  }
  return key16_to_bf16_bits(lo);
}

// visit every element of the pieces whose maximum reaches tf (lanes take pieces round robin; slow path only)
template <typename T, int PIECE, typename F>
__device__ __forceinline__ void for_each_in_pieces(const T* __restrict__ row, int V, const uint16_t* pv, int n_pieces,
                                                   float tf, int lane, F&& fn) {
  for (int i = lane; i < n_pieces; i += 32) {
    if (!(bf16_bits_to_float(pv[i]) >= tf)) continue;
    const int c0 = i * PIECE;
    if (c0 + PIECE <= V) {
      Vec8<T> e[PIECE / 8];
#pragma unroll
      for (int q = 0; q < PIECE / 8; ++q) e[q].load_global(row + c0 + 8 * q);
#pragma unroll
      for (int q = 0; q < PIECE / 8; ++q) {
        float g8[8];
        e[q].unpack(g8);
#pragma unroll
        for (int u = 0; u < 8; ++u) fn(g8[u], c0 + 8 * q + u);
      }
    } else {
      for (int c = c0; c < V; ++c) fn(Elem<T>::to_f(row[c]), c);
    }
  }
}

template <typename T, int PIECE, typename P>
__device__ __forceinline__ int warp_count_if(const T* row, int V, const uint16_t* pv, int n_pieces, float tf, int lane,
                                             P&& pred) {
  int c = 0;
  for_each_in_pieces<T, PIECE>(row, V, pv, n_pieces, tf, lane, [&](float x, int idx) { c += pred(order_key(x), idx) ? 1 : 0; });
  return __reduce_add_sync(0xffffffffu, c);
}

// threshold -> candidates -> (slow path) -> the k outputs of row r; lm / ll = row maximum and log of the exp sum
template <typename T, int PIECE>
__device__ __forceinline__ void warp_select_emit(const T* __restrict__ row, int V, int k, const WarpRow& w, int n_pieces,
                                                 int n_pieces_pad, float lm, float ll, int64_t r,
                                                 __half* __restrict__ out_v, int32_t* __restrict__ out_i, int lane) {
  const uint32_t tb = warp_kth_largest_piece(w.pv, n_pieces_pad, n_pieces, k, lane);
  const float tf = bf16_bits_to_float(tb);  // at least k pieces, hence k elements, are >= tf
  if (lane == 0) {
    w.count[0] = 0;  // candidates
    w.count[1] = 0;  // qualifying pieces
  }
  __syncwarp();
  auto push = [&](float x, int idx) {
    const int slot = atomicAdd(w.count, 1);
    if (slot < kWrCap) w.cand[slot] = ((uint64_t)order_key(x) << 32) | (uint32_t)(0xffffffffu - (uint32_t)idx);
  };
  // (a) list the qualifying pieces (about k of them), (b) one piece per lane: the 64-byte reads of 32 pieces are in
  // flight together.  Visiting them in place costs one dependent global round trip per piece (measured: 70 us a row).
  {
    const uint4* kv = reinterpret_cast<const uint4*>(w.pv);
    const uint32_t t2u = tb * 0x00010001u;
    const __nv_bfloat162 t2 = *reinterpret_cast<const __nv_bfloat162*>(&t2u);
    for (int v = lane; v < (n_pieces_pad >> 3); v += 32) {
      const uint4 x = kv[v];
      const uint32_t wd[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 h = ge2(wd[q], t2);
        const uint32_t hit = *reinterpret_cast<const uint32_t*>(&h);
        if (hit == 0u) continue;
#pragma unroll
        for (int hw = 0; hw < 2; ++hw) {
          const int piece = 8 * v + 2 * q + hw;
          if (((hit >> (16 * hw)) & 0xffffu) && piece < n_pieces) {
            const int slot = atomicAdd(w.count + 1, 1);
            if (slot < kWrCap) w.plist[slot] = (uint16_t)piece;
          }
        }
      }
    }
  }
  __syncwarp();
  const int np = w.count[1];
  if (np <= kWrCap) {
    for (int j = lane; j < np; j += 32) {
      const int c0 = (int)w.plist[j] * PIECE;
      if (c0 + PIECE <= V) {
        Vec8<T> e[PIECE / 8];
#pragma unroll
        for (int q = 0; q < PIECE / 8; ++q) e[q].load_global(row + c0 + 8 * q);
#pragma unroll
        for (int q = 0; q < PIECE / 8; ++q) {
          float g8[8];
          e[q].unpack(g8);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (g8[u] >= tf) push(g8[u], c0 + 8 * q + u);
        }
      } else {
        for (int c = c0; c < V; ++c) {
          const float x = Elem<T>::to_f(row[c]);
          if (x >= tf) push(x, c);
        }
      }
    }
  }
  __syncwarp();
  int count = np <= kWrCap ? w.count[0] : kWrCap + 1;  // too many pieces = too many candidates: slow path
  if (count > kWrCap) {
    // exact slow path (massive ties): bisection on the key, then on the index among the ties at the k-th key
    uint32_t lo = order_key(tf), hi = 0xffffffffu;  // invariant: count(key >= lo) >= k
    while (lo < hi) {
      const uint32_t mid = lo + (uint32_t)(((uint64_t)hi - lo + 1) >> 1);
      const int c = warp_count_if<T, PIECE>(row, V, w.pv, n_pieces, tf, lane, [&](uint32_t key, int) { return key >= mid; });
      if (c >= k) lo = mid; else hi = mid - 1;
    }
    const uint32_t kth = lo;
    const int above = warp_count_if<T, PIECE>(row, V, w.pv, n_pieces, tf, lane, [&](uint32_t key, int) { return key > kth; });
    const int need = k - above;
    int jl = 0, jh = V - 1;
    while (jl < jh) {
      const int mid = jl + ((jh - jl) >> 1);
      const int c = warp_count_if<T, PIECE>(row, V, w.pv, n_pieces, tf, lane,
                                     [&](uint32_t key, int idx) { return key == kth && idx <= mid; });
      if (c >= need) jh = mid; else jl = mid + 1;
    }
    const int jmax = jl;
    __syncwarp();
    if (lane == 0) w.count[0] = 0;
    __syncwarp();
    for_each_in_pieces<T, PIECE>(row, V, w.pv, n_pieces, tf, lane, [&](float x, int idx) {
      const uint32_t key = order_key(x);
      if (key > kth || (key == kth && idx <= jmax)) push(x, idx);
    });
    __syncwarp();
    count = w.count[0];  // == k
  }
  const int n = count < kWrCap ? count : kWrCap;
  for (int c = lane; c < n; c += 32) {
    const uint64_t mine = w.cand[c];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += w.cand[j] > mine ? 1 : 0;
    if (rank < k) {
      const float x = key_to_float((uint32_t)(mine >> 32));
      const int idx = (int)(0xffffffffu - (uint32_t)(mine & 0xffffffffu));
      const float lp = (x - lm) - ll;
      const float lp_r = Elem<T>::to_f(Elem<T>::from_f(lp));  // log_softmax returns the logits' dtype
      out_v[r * k + rank] = __float2half_rn(lp_r);
      out_i[r * k + rank] = idx;
    }
  }
}

// pass 1 of the warp form: one sweep over the row -> piece maxima in shared memory, row maximum and log of the exp sum.
// Two register batches of U vectors per lane: while one batch is reduced the other's loads are in flight, so a
// row-warp always has U..2U 16-byte loads outstanding (a single batch alternated between "all in flight" and "none").
// maximum of the 8 elements of a vector.  16-bit rows: three packed maxima (HMNMX2) on the raw words + one across the
// two halves instead of 8 conversions + 7 FMNMX (pass 1 is issue-bound once enough loads are in flight); NaN handling
// is fmaxf's (the other operand wins).
template <typename T>
__device__ __forceinline__ float vec_max8(const Vec8<T>& v) {
  float f[8];
  v.unpack(f);
  return fmaxf(fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3])), fmaxf(fmaxf(f[4], f[5]), fmaxf(f[6], f[7])));
}
template <>
__device__ __forceinline__ float vec_max8<__nv_bfloat16>(const Vec8<__nv_bfloat16>& v) {
  const __nv_bfloat162 a = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&v.a.x),
                                   *reinterpret_cast<const __nv_bfloat162*>(&v.a.y));
  const __nv_bfloat162 b = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&v.a.z),
                                   *reinterpret_cast<const __nv_bfloat162*>(&v.a.w));
  const __nv_bfloat162 c = __hmax2(a, b);
  const uint32_t w = *reinterpret_cast<const uint32_t*>(&c);
  return fmaxf(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <>
__device__ __forceinline__ float vec_max8<__half>(const Vec8<__half>& v) {
  const __half2 a = __hmax2(*reinterpret_cast<const __half2*>(&v.a.x), *reinterpret_cast<const __half2*>(&v.a.y));
  const __half2 b = __hmax2(*reinterpret_cast<const __half2*>(&v.a.z), *reinterpret_cast<const __half2*>(&v.a.w));
  const float2 c = __half22float2(__hmax2(a, b));
  return fmaxf(c.x, c.y);
}

// packed 16-bit pairs: maximum of two words, low element as a float
template <typename T>
struct Pair16;
template <>
struct Pair16<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  static __device__ __forceinline__ float low(uint32_t w) { return __uint_as_float(w << 16); }
};
template <>
struct Pair16<__half> {
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  static __device__ __forceinline__ float low(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
};

// One register batch of the statistics sweep (shared by the warp-per-row kernel and kd_topk_stats_kernel): lane l holds
// vectors base + 32 u + l (u < U) of a row; four neighbouring lanes hold one 32-element piece.  Writes the piece maxima
// (16-bit keys) to `pieces` (indexed by piece number; shared or global memory) and folds the batch into the lane's
// online (m, s) = (reference maximum, sum e^{x - m}).
// FULL = every vector of the batch lies inside the row: no per-vector predicates.  The sweep is issue-bound, not
// HBM-bound (ncu: 10.6 issued instructions per element in the first version, 66 % issue utilisation at 0.54 of the
// DRAM peak), so the 16-bit FULL path is written for instruction count: piece maxima stay packed (HMNMX2 on the raw
// words, one shuffle for both halves), the lane's reference maximum is the maximum of its PIECES (a superset of its
// own elements - any upper bound taken from the row serves the online sum), and the four keys of a lane group leave
// in one 32-lane store per four vectors.
template <typename T, int U, bool FULL, int PIECE = kWrPiece>
__device__ __forceinline__ void sweep_batch(const Vec8<T> (&v)[U], int base, int nvec, int lane,
                                            uint16_t* __restrict__ pieces, float& m, float& s) {
  constexpr int LP = PIECE / 8;  // lanes (16-byte vectors) per piece: 4 or 8
  static_assert(LP == 4 || LP == 8, "a piece is 32 or 64 elements");
  float vm = -CUDART_INF_F;
  bool skip[U];
  if constexpr (FULL && sizeof(T) == 2 && (U % 4) == 0) {
    float pmf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint32_t c = Pair16<T>::max2(Pair16<T>::max2(v[u].a.x, v[u].a.y), Pair16<T>::max2(v[u].a.z, v[u].a.w));
      c = Pair16<T>::max2(c, __shfl_xor_sync(0xffffffffu, c, 1));
      c = Pair16<T>::max2(c, __shfl_xor_sync(0xffffffffu, c, 2));
      if (LP == 8) c = Pair16<T>::max2(c, __shfl_xor_sync(0xffffffffu, c, 4));
      c = Pair16<T>::max2(c, __byte_perm(c, 0, 0x1032));  // both halves = the piece maximum
      pmf[u] = Pair16<T>::low(c);
      vm = fmaxf(vm, pmf[u]);
      skip[u] = false;
    }
#pragma unroll
    for (int u0 = 0; u0 < U; u0 += 4) {  // lane j of a piece's lane group stores the key of vector u0 + j: one store
      const int j = lane & (LP - 1);     // per 4 vectors (all 32 lanes for 32-element pieces, 16 for 64-element ones)
      const float mine = j == 0 ? pmf[u0] : (j == 1 ? pmf[u0 + 1] : (j == 2 ? pmf[u0 + 2] : pmf[u0 + 3]));
      if (LP == 4 || j < 4) pieces[((base + (u0 + (j & 3)) * 32) / LP) + (lane / LP)] = piece_key<T>(mine);
    }
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 32 + lane;
      float x = (FULL || idx < nvec) ? vec_max8<T>(v[u]) : -CUDART_INF_F;
      skip[u] = !FULL && x == -CUDART_INF_F;  // past the end of the row: the registers hold stale data
      vm = fmaxf(vm, x);
      x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));  // LP lanes = one piece
      x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
      if (LP == 8) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 4));
      if ((lane & (LP - 1)) == 0 && (FULL || idx < nvec)) pieces[idx / LP] = piece_key<T>(x);
    }
  }
  if (vm > m) {
    s *= exp_diff(m, vm, kLog2e);
    m = vm;
  }
  if (m != -CUDART_INF_F) {
    const float off = m * kLog2e;
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (skip[u]) continue;
      float f[8];
      v[u].unpack(f);  // an -inf element gives ex2(-inf) = 0 (m is finite here)
      p0 += ex2(fmaf(f[0], kLog2e, -off)) + ex2(fmaf(f[4], kLog2e, -off));
      p1 += ex2(fmaf(f[1], kLog2e, -off)) + ex2(fmaf(f[5], kLog2e, -off));
      p2 += ex2(fmaf(f[2], kLog2e, -off)) + ex2(fmaf(f[6], kLog2e, -off));
      p3 += ex2(fmaf(f[3], kLog2e, -off)) + ex2(fmaf(f[7], kLog2e, -off));
    }
    s += (p0 + p1) + (p2 + p3);
  }
}

// warp-per-row kernel: 16-byte loads per lane and register batch (two batches), elements per piece, CTAs per SM the
// register budget allows (compile-time A/B switches: speech-distill_b200/build.py --variant NAME -D...).
#ifndef KD_TOPK_WARP_U
#define KD_TOPK_WARP_U 8
#endif
#ifndef KD_TOPK_WARP_PIECE
#define KD_TOPK_WARP_PIECE kWrPiece
#endif
#ifndef KD_TOPK_WARP_MINB
#define KD_TOPK_WARP_MINB 2
#endif

template <typename T, int U, int PIECE>
struct Pass1Batch {
  Vec8<T> v[U];
  __device__ __forceinline__ void load(const T* __restrict__ row, int base, int nvec, int lane) {
    const T* p0 = row + (size_t)(base + lane) * 8;
    if (base + 32 * U <= nvec) {  // whole batch inside the row: one address, immediate offsets, no predicates
#pragma unroll
      for (int u = 0; u < U; ++u) v[u].load_global(p0 + u * 256);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (base + u * 32 + lane < nvec) v[u].load_global(p0 + u * 256);
    }
  }
  __device__ __forceinline__ void reduce(int base, int nvec, int lane, uint16_t* pv, float& m, float& s) {
    if (base + 32 * U <= nvec) sweep_batch<T, U, true, PIECE>(v, base, nvec, lane, pv, m, s);
    else sweep_batch<T, U, false, PIECE>(v, base, nvec, lane, pv, m, s);
  }
};

template <typename T, int U, int PIECE>
__device__ __forceinline__ void warp_pass1(const T* __restrict__ row, int V, uint16_t* pv, int lane, float& lm, float& ll) {
  const int nvec = V >> 3;
  constexpr int kStep = 32 * U;
  float m = -CUDART_INF_F, s = 0.f;
  Pass1Batch<T, U, PIECE> b0, b1;
  b0.load(row, 0, nvec, lane);
#pragma unroll 1
  for (int base = 0; base < nvec; base += 2 * kStep) {
    b1.load(row, base + kStep, nvec, lane);
    b0.reduce(base, nvec, lane, pv, m, s);
    b0.load(row, base + 2 * kStep, nvec, lane);
    b1.reduce(base + kStep, nvec, lane, pv, m, s);
  }
  // the V % 8 elements behind the last whole vector: lane 0, scalar
  const int tail0 = nvec << 3;
  float tmax = -CUDART_INF_F;
  if (lane == 0) {
    for (int c = tail0; c < V; ++c) {
      const float x = Elem<T>::to_f(row[c]);
      tmax = fmaxf(tmax, x);
      if (x > m) {
        s *= exp_diff(m, x, kLog2e);
        m = x;
      }
      if (m != -CUDART_INF_F) s += ex2((x - m) * kLog2e);
    }
  }
  __syncwarp();
  if (lane == 0 && tail0 < V) {
    const int piece = tail0 / PIECE;
    // the tail opens a new piece when the whole vectors fill their pieces, else it joins the last one
    const float prev = (nvec % (PIECE / 8)) ? bf16_bits_to_float(pv[piece]) : -CUDART_INF_F;
    pv[piece] = bf16_bits_rd(fmaxf(prev, tmax));
  }
  const float wm = warp_max(m);
  const float ws = warp_sum(s * exp_diff(m, wm, kLog2e));
  lm = wm;
  ll = ln_acc(ws);
  __syncwarp();
}

// (Measured and not kept, configs[2] size on one B200: 64-element pieces with 4 loads per batch and three CTAs per SM -
//  21 - 24 row-warps per SM instead of 16 - 717 us; a per-lane cp.async ring of 12 x 16 bytes in shared memory instead
//  of the register double buffer 842 us; the configuration below 590 us.  More row-warps lengthen every row in
//  proportion: the SM's XU / issue / load-path mix is saturated at ~0.11 rows per microsecond either way.
//  A bulk L2 prefetch (cp.async.bulk.prefetch.L2 by lane 0, 2 / 4 passes of 8 KB ahead of the register loads - the
//  step that took 7 % off K2) measured 635 / 664 us against 623 us without it on the same box: with 14 row-warps of
//  8 KB in flight each the loads are already deep enough, and the extra requests only compete with them.)
template <typename T>
__global__ void __launch_bounds__(32 * kWrWarps, sizeof(T) == 2 ? KD_TOPK_WARP_MINB : 2) kd_topk_warp_kernel(const T* __restrict__ logits, int64_t R, int V,
                                                                       int64_t row_stride, int k,
                                                                       __half* __restrict__ out_v,
                                                                       int32_t* __restrict__ out_i, int n_pieces_pad) {
  extern __shared__ __align__(16) unsigned char wr_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PIECE = KD_TOPK_WARP_PIECE;
  const WarpRow w = warp_row_smem(wr_smem, warp, n_pieces_pad);
  const int n_pieces = (V + PIECE - 1) / PIECE;
  for (int i = n_pieces + lane; i < n_pieces_pad; i += 32) w.pv[i] = kBf16NegInf;  // padding: never >= a threshold
  __syncwarp();
  const int n_warps = blockDim.x >> 5;  // rows per CTA (warp_form_config)
  for (int64_t r = (int64_t)blockIdx.x * n_warps + warp; r < R; r += (int64_t)gridDim.x * n_warps) {
    const T* row = logits + r * row_stride;
    float lm, ll;
    warp_pass1<T, (sizeof(T) == 2 ? KD_TOPK_WARP_U : (KD_TOPK_WARP_U > 2 ? KD_TOPK_WARP_U / 2 : 2)), PIECE>(row, V, w.pv, lane,
                                                                                                        lm, ll);
    warp_select_emit<T, PIECE>(row, V, k, w, n_pieces, n_pieces_pad, lm, ll, r, out_v, out_i, lane);
    __syncwarp();
  }
}

// ---- selection behind the teacher head GEMM (kd_head_logits_stats) -----------------------------------------------
// Same output as kd_topk_logprobs on the block's bf16 logits, without the sweep over the row: the head GEMM's epilogue
// left (a) the maximum of every 32-column piece and (b) partial (max, sum exp) records, so a row costs the merge of
// n_part records (~2 KB), the piece maxima (~10 KB) and the ~k pieces of 64 bytes that can hold a top-k entry.
template <typename T>
__global__ void __launch_bounds__(32 * kWrWarps) kd_head_select_kernel(const T* __restrict__ logits, int64_t R,
                                                                         int V, int64_t row_stride, int k,
                                                                         const __nv_bfloat16* __restrict__ pmax,
                                                                         int pmax_stride, const float2* __restrict__ part,
                                                                         int part_stride, int n_part,
                                                                         __half* __restrict__ out_v,
                                                                         int32_t* __restrict__ out_i, int n_pieces_pad) {
  extern __shared__ __align__(16) unsigned char wr_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WarpRow w = warp_row_smem(wr_smem, warp, n_pieces_pad);
  const int n_pieces = (V + kWrPiece - 1) / kWrPiece;
  const int n_warps = blockDim.x >> 5;  // rows per CTA (warp_form_config)
  for (int64_t r = (int64_t)blockIdx.x * n_warps + warp; r < R; r += (int64_t)gridDim.x * n_warps) {
    const T* row = logits + r * row_stride;
    // piece maxima (already bf16): straight copy, 8 per 16-byte load; entries past the vocabulary hold -inf
    const uint4* pm = reinterpret_cast<const uint4*>(pmax + r * pmax_stride);
    uint4* dst = reinterpret_cast<uint4*>(w.pv);
    for (int v = lane; v < (n_pieces_pad >> 3); v += 32) dst[v] = pm[v];
    // log-sum-exp from the epilogue's partial records
    float m = -CUDART_INF_F, s = 0.f;
    for (int i = lane; i < n_part; i += 32) {
      const float2 rec = part[r * part_stride + i];
      if (rec.x > m) {
        s *= exp_diff(m, rec.x, kLog2e);
        m = rec.x;
      }
      if (rec.x != -CUDART_INF_F) s += rec.y * exp_diff(rec.x, m, kLog2e);
    }
    const float lm = warp_max(m);
    const float ll = ln_acc(warp_sum(s * exp_diff(m, lm, kLog2e)));
    __syncwarp();
    warp_select_emit<T, kWrPiece>(row, V, k, w, n_pieces, n_pieces_pad, lm, ll, r, out_v, out_i, lane);
    __syncwarp();
  }
}


// ---- two-kernel form of kd_topk_logprobs: statistics sweep + selection ----------------------------------------------
// The warp-per-row kernel above keeps a row's piece maxima in shared memory, so an SM holds 16 rows at a time: the
// rows of a launch advance in lock step (8192 rows on 2368 row slots = 3.46 rounds cost 4), and a slot does not stream
// while it selects.  Here the sweep and the selection are separate kernels that overlap row block by row block:
//   kd_topk_stats_kernel : work item = (row, 8192-element segment), one warp each, taken round robin - fine-grained, no
//                          shared memory, nothing but streaming: piece maxima (bf16, rounded down) and one (max, sum exp)
//                          record per item go to a workspace (1/32 of the logits' bytes);
//   kd_head_select_kernel: the selection behind the teacher head GEMM, unchanged - it reads the piece maxima, the
//                          records and the ~k qualifying 64-byte pieces of every row.
// The sweep of row block b + 1 (internal high-priority stream) runs beside the selection of block b (caller's stream):
// three 256-thread sweep CTAs (<= 64 registers) and one selection CTA fit an SM together.
constexpr int kStatSeg = 8192;                // elements per work item (256 pieces)
constexpr int kStatThreads = 256;
constexpr int kStatCtasPerSm = 3;
constexpr int kStatU = 4;                     // 16-byte loads per lane and register batch (two batches)
constexpr int kStatMaxBlocks = 8;             // row blocks of the sweep / selection pipeline
#ifndef KD_STAT_MINB
#define KD_STAT_MINB 3                        // sweep CTAs the register budget allows per SM (3: <= 80 registers, no spills; 4 = 64 registers spilled and measured 10 % slower)
#endif

// position of a warp in its sequence of batches: item wid + j * wstride, batch b of the item; (row, seg) are carried
// along instead of being divided out of the item number for every batch
struct StatCursor {
  int64_t item, row;
  int seg, b;
};

template <typename T, int U>
struct StatBatch {
  Vec8<T> v[U];
  const T* rowp;
  int64_t row;
  int seg, vbase;
  bool live, last;
  template <int NB>
  __device__ __forceinline__ void load(StatCursor& c, int64_t wstride, int64_t drow, int dseg, int64_t n_items, int n_seg,
                                       int nvec, const T* __restrict__ logits, int64_t row_stride, int lane) {
    live = c.item < n_items;
    if (!live) return;
    last = c.b == NB - 1;
    row = c.row;
    seg = c.seg;
    rowp = logits + row * row_stride;
    vbase = seg * (kStatSeg / 8) + c.b * (32 * U);
    const T* p0 = rowp + (size_t)(vbase + lane) * 8;
    if (vbase + 32 * U <= nvec) {  // whole batch inside the row: one address, immediate offsets, no predicates
#pragma unroll
      for (int u = 0; u < U; ++u) v[u].load_global(p0 + u * 256);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (vbase + u * 32 + lane < nvec) v[u].load_global(p0 + u * 256);
    }
    if (++c.b == NB) {  // next item of this warp
      c.b = 0;
      c.item += wstride;
      c.row += drow;
      c.seg += dseg;
      if (c.seg >= n_seg) {
        c.seg -= n_seg;
        ++c.row;
      }
    }
  }
  __device__ __forceinline__ void reduce(int nvec, int lane, uint16_t* __restrict__ pm_row, float& m, float& s) const {
    if (vbase + 32 * U <= nvec) sweep_batch<T, U, true>(v, vbase, nvec, lane, pm_row, m, s);
    else sweep_batch<T, U, false>(v, vbase, nvec, lane, pm_row, m, s);
  }
};

template <typename T, int U>
__global__ void __launch_bounds__(kStatThreads, KD_STAT_MINB)
kd_topk_stats_kernel(const T* __restrict__ logits, int64_t R, int V, int64_t row_stride, uint16_t* __restrict__ pmax,
                     int pmax_stride, float2* __restrict__ part, int part_stride) {
  constexpr int NB = kStatSeg / 8 / (32 * U);  // batches per item
  static_assert(NB >= 2 && (NB & 1) == 0, "two register batches alternate inside an item");
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (kStatThreads / 32) + (threadIdx.x >> 5);
  const int64_t wstride = (int64_t)gridDim.x * (kStatThreads / 32);
  const int nvec = V >> 3;
  const int n_seg = (V + kStatSeg - 1) / kStatSeg;
  const int n_pieces = (V + kWrPiece - 1) / kWrPiece;
  const int64_t n_items = R * n_seg;
  const int64_t drow = wstride / n_seg;
  const int dseg = (int)(wstride - drow * n_seg);
  StatCursor cur;
  cur.item = wid;
  cur.row = wid / n_seg;
  cur.seg = (int)(wid - cur.row * n_seg);
  cur.b = 0;
  float m = -CUDART_INF_F, s = 0.f;
  StatBatch<T, U> b0, b1;
  b0.template load<NB>(cur, wstride, drow, dseg, n_items, n_seg, nvec, logits, row_stride, lane);
#pragma unroll 1
  while (b0.live) {
    // NB is even: b0 holds the even batches of an item, b1 the odd ones - an item always ends in b1
    b1.template load<NB>(cur, wstride, drow, dseg, n_items, n_seg, nvec, logits, row_stride, lane);
    b0.reduce(nvec, lane, pmax + b0.row * pmax_stride, m, s);
    const int64_t row1 = b1.row;
    const int seg1 = b1.seg;
    const T* rowp1 = b1.rowp;
    const bool last1 = b1.last;
    b1.reduce(nvec, lane, pmax + row1 * pmax_stride, m, s);
    b0.template load<NB>(cur, wstride, drow, dseg, n_items, n_seg, nvec, logits, row_stride, lane);
    if (last1) {  // the item's last batch has been reduced: publish its record
      uint16_t* pm_row = pmax + row1 * pmax_stride;
      if (seg1 == n_seg - 1) {
        // row end: the V % 8 elements behind the last whole vector (lane 0, scalar) and the padding of the piece maxima
        __syncwarp();
        const int tail0 = nvec << 3;
        if (lane == 0 && tail0 < V) {
          const int piece = tail0 / kWrPiece;
          float pmx = -CUDART_INF_F;
          for (int c = piece * kWrPiece; c < V; ++c) {  // the whole last piece again: <= 31 elements
            const float x = Elem<T>::to_f(rowp1[c]);
            pmx = fmaxf(pmx, x);
            if (c >= tail0) {
              if (x > m) {
                s *= exp_diff(m, x, kLog2e);
                m = x;
              }
              if (m != -CUDART_INF_F) s += ex2((x - m) * kLog2e);
            }
          }
          pm_row[piece] = bf16_bits_rd(pmx);
        }
        for (int i = n_pieces + lane; i < pmax_stride; i += 32) pm_row[i] = kBf16NegInf;  // never >= a threshold
      }
      const float wm = warp_max(m);
      const float ws = warp_sum(s * exp_diff(m, wm, kLog2e));
      if (lane == 0) part[row1 * part_stride + seg1] = make_float2(wm, ws);
      m = -CUDART_INF_F;
      s = 0.f;
    }
  }
}

struct TopkPipe {
  std::mutex enqueue;
  cudaStream_t hi = nullptr;
  cudaEvent_t fork = nullptr, swept[kStatMaxBlocks] = {};
  bool ready = false;
};
}  // namespace kd

using namespace kd;

// launch geometry of the warp-per-row kernels: up to 8 row-warps per CTA, as many CTAs per SM as the shared memory
// holds (at most 4); returns false when a row's piece keys do not fit (V beyond ~3.4 M columns).
// The rows of a launch advance in rounds (every row slot takes its next row at about the same time) and the kernels
// are issue-bound, so a round's duration grows with the row-warps per SM: 8192 rows on 148 x 2 x 8 slots are 3.46 -> 4
// rounds of 8-warp CTAs, but also 4 rounds of 7-warp CTAs, which finish 1/8 sooner.  `warps` = the rows-per-CTA in
// [4, 8] that minimises rounds x warps.
template <typename Kern>
static bool warp_form_config(Kern kern, int64_t R, int V, int* grid, size_t* smem, int* n_pieces_pad, int* warps,
                             int piece = kWrPiece, int max_per_sm = 4) {
  const int n_pieces = (V + piece - 1) / piece;
  *n_pieces_pad = (n_pieces + 7) & ~7;
  const size_t smem_max = (size_t)kWrWarps * warp_row_bytes(*n_pieces_pad);
  if (smem_max > 226 * 1024 || n_pieces > 65535) return false;  // piece indices are kept as 16-bit values
  int per_sm = (int)((size_t)232448 / (smem_max + 1024));
  if (per_sm > max_per_sm) per_sm = max_per_sm;
  if (per_sm < 1) per_sm = 1;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int best_w = kWrWarps;
  int64_t best_cost = INT64_MAX;
  for (int w = kWrWarps; w >= 4; --w) {
    const int64_t slots = (int64_t)sms * per_sm * w;
    const int64_t cost = ((R + slots - 1) / slots) * w;
    if (cost < best_cost) {
      best_cost = cost;
      best_w = w;
    }
  }
  *warps = best_w;
  *smem = (size_t)best_w * warp_row_bytes(*n_pieces_pad);
  const int64_t ctas = (R + best_w - 1) / best_w;
  *grid = (int)(ctas < (int64_t)sms * per_sm ? ctas : (int64_t)sms * per_sm);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}
// Which form takes a call of kd_topk_logprobs: a row is one warp's work in the warp form (~0.15 ms per row at
// V = 152,936 however few rows there are), the CTA-per-row form spreads a row over 256 threads.  Measured at k = 64
// (tools/k3_forms.py, us, L2 flushed between launches):  R      256  1024  2048  4096  8192
//                                                        cta     53   125   252   498  1026
//                                                        warp   102   115   158   299   588
//                                                        two     61   104   203   358   670
// -> warp form from half an SM's row slots per SM upwards.  KD_TOPK_FORM=cta / warp forces one of them (A/B runs).
static bool topk_use_warp_form(int64_t R) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("KD_TOPK_FORM");
    v = !e ? 0 : (e[0] == 'c' ? 1 : (e[0] == 'w' ? 2 : 0));
  }
  if (v == 1) return false;
  if (v == 2) return true;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return R >= (int64_t)sms * kWrWarps;
}

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
  cudaStream_t s = (cudaStream_t)stream;
  if (vec_ok && k <= kWrCap / 2 && topk_use_warp_form(R)) {  // warp-per-row form (see above)
    int grid = 0, npp = 0, nw = 0;
    size_t smem = 0;
    bool launched = false;
    switch (dtype) {
      case KD_DTYPE_F32:
        if (warp_form_config(kd_topk_warp_kernel<float>, R, V, &grid, &smem, &npp, &nw, KD_TOPK_WARP_PIECE, KD_TOPK_WARP_MINB)) {
          kd_topk_warp_kernel<float><<<grid, 32 * nw, smem, s>>>((const float*)logits, R, V, row_stride, k,
                                                                       (__half*)out_v, out_i, npp);
          launched = true;
        }
        break;
      case KD_DTYPE_BF16:
        if (warp_form_config(kd_topk_warp_kernel<__nv_bfloat16>, R, V, &grid, &smem, &npp, &nw, KD_TOPK_WARP_PIECE, KD_TOPK_WARP_MINB)) {
          kd_topk_warp_kernel<__nv_bfloat16><<<grid, 32 * nw, smem, s>>>((const __nv_bfloat16*)logits, R, V,
                                                                               row_stride, k, (__half*)out_v, out_i, npp);
          launched = true;
        }
        break;
      case KD_DTYPE_F16:
        if (warp_form_config(kd_topk_warp_kernel<__half>, R, V, &grid, &smem, &npp, &nw, KD_TOPK_WARP_PIECE, KD_TOPK_WARP_MINB)) {
          kd_topk_warp_kernel<__half><<<grid, 32 * nw, smem, s>>>((const __half*)logits, R, V, row_stride, k,
                                                                        (__half*)out_v, out_i, npp);
          launched = true;
        }
        break;
      default:
        set_error("kd_topk_logprobs: unsupported dtype code %d", dtype);
        return 1;
    }
    if (launched) return check_launch("kd_topk (warp form) launch");
  }
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

// ---- two-kernel form: host side -------------------------------------------------------------------------------------
static TopkPipe* get_topk_pipe() {
  static TopkPipe pipes[kMaxDevices];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  TopkPipe& p = pipes[dev];
  if (p.ready) return &p;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  // the sweep's CTAs are placed first; the selection (caller's stream) takes the room they leave on every SM
  if (cudaStreamCreateWithPriority(&p.hi, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  for (int i = 0; i < kStatMaxBlocks; ++i)
    if (cudaEventCreateWithFlags(&p.swept[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  p.ready = true;
  return &p;
}

static inline int topk_pmax_stride(int V) { return (((V + kWrPiece - 1) / kWrPiece) + 7) & ~7; }
static inline int topk_part_stride(int V) { return (V + kStatSeg - 1) / kStatSeg; }
static inline size_t topk_pmax_bytes(int64_t R, int V) {
  return (((size_t)R * topk_pmax_stride(V) * 2) + 255) & ~(size_t)255;
}

extern "C" size_t kd_topk_workspace_bytes(int64_t R, int V) {
  if (R <= 0 || V <= 0) return 0;
  return topk_pmax_bytes(R, V) + (size_t)R * topk_part_stride(V) * sizeof(float2);
}

template <typename T>
static int topk_two_kernel(const T* logits, int64_t R, int V, int64_t row_stride, int k, __half* out_v, int32_t* out_i,
                           uint8_t* ws, cudaStream_t s) {
  TopkPipe* pipe = get_topk_pipe();
  if (!pipe) return -1;
  int sel_grid = 0, npp = 0, sel_warps = 0;
  size_t sel_smem = 0;
  if (!warp_form_config(kd_head_select_kernel<T>, R, V, &sel_grid, &sel_smem, &npp, &sel_warps)) return -1;
  sel_warps = kWrWarps;  // the row blocks differ in size: keep full CTAs
  sel_smem = (size_t)kWrWarps * warp_row_bytes(npp);
  const int pmax_stride = topk_pmax_stride(V), part_stride = topk_part_stride(V);
  uint16_t* pmax = reinterpret_cast<uint16_t*>(ws);
  float2* part = reinterpret_cast<float2*>(ws + topk_pmax_bytes(R, V));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // row blocks: enough of them that the last block's selection (the only one nothing hides) is short, each big enough
  // to fill the GPU's sweep warps a few times over
  int n_blk = (int)(R / 1024);
  n_blk = n_blk < 1 ? 1 : (n_blk > kStatMaxBlocks ? kStatMaxBlocks : n_blk);
  // diagnosis knobs (tools/k3_ab.py): KD_TOPK_BLOCKS, KD_TOPK_SWEEP_CTAS (sweep CTAs per SM), KD_TOPK_PHASE=sweep /
  // select runs one of the two kernels only (the results are then stale or missing)
  static int env_blocks = -1, env_ctas = 0, env_phase = 0;
  if (env_blocks < 0) {
    const char* e = getenv("KD_TOPK_BLOCKS");
    const char* c = getenv("KD_TOPK_SWEEP_CTAS");
    const char* ph = getenv("KD_TOPK_PHASE");
    env_ctas = c ? atoi(c) : 0;
    env_phase = !ph ? 0 : (ph[0] == 's' && ph[1] == 'w' ? 1 : (ph[0] == 's' && ph[1] == 'e' ? 2 : 0));
    env_blocks = e ? atoi(e) : 0;
  }
  if (env_blocks > 0) n_blk = env_blocks > kStatMaxBlocks ? kStatMaxBlocks : env_blocks;
  const int sweep_ctas = env_ctas > 0 ? env_ctas : kStatCtasPerSm;
  const int64_t rows_per_blk = ((R + n_blk - 1) / n_blk + kWrWarps - 1) / kWrWarps * kWrWarps;
  std::lock_guard<std::mutex> lock(pipe->enqueue);
  if (check_cuda(cudaEventRecord(pipe->fork, s), "topk fork")) return 1;
  if (check_cuda(cudaStreamWaitEvent(pipe->hi, pipe->fork, 0), "topk fork")) return 1;
  int b = 0;
  for (int64_t r0 = 0; r0 < R; r0 += rows_per_blk, ++b) {
    const int64_t rows = R - r0 < rows_per_blk ? R - r0 : rows_per_blk;
    const int64_t items = rows * part_stride;
    const int64_t warps = (int64_t)sms * sweep_ctas * (kStatThreads / 32);
    const int grid = (int)((items < warps ? items : warps) + kStatThreads / 32 - 1) / (kStatThreads / 32);
    if (env_phase != 2) {
      kd_topk_stats_kernel<T, (sizeof(T) == 2 ? kStatU : kStatU / 2)><<<grid, kStatThreads, 0, pipe->hi>>>(
          logits + r0 * row_stride, rows, V, row_stride, pmax + r0 * pmax_stride, pmax_stride, part + r0 * part_stride,
          part_stride);
      if (check_launch("kd_topk_stats launch")) return 1;
    }
    if (check_cuda(cudaEventRecord(pipe->swept[b], pipe->hi), "topk swept")) return 1;
    if (check_cuda(cudaStreamWaitEvent(s, pipe->swept[b], 0), "topk swept")) return 1;
    const int64_t ctas = (rows + kWrWarps - 1) / kWrWarps;
    const int g = (int)(ctas < sel_grid ? ctas : sel_grid);
    if (env_phase != 1) {
      kd_head_select_kernel<T><<<g, 32 * kWrWarps, sel_smem, s>>>(
          logits + r0 * row_stride, rows, V, row_stride, k, reinterpret_cast<const __nv_bfloat16*>(pmax + r0 * pmax_stride),
          pmax_stride, part + r0 * part_stride, part_stride, part_stride, out_v + r0 * k, out_i + r0 * k, npp);
      if (check_launch("kd_topk select launch")) return 1;
    }
  }
  return 0;
}

// kd_topk_logprobs with a caller-provided workspace of kd_topk_workspace_bytes(R, V): the two-kernel form (statistics
// sweep beside the selection of the previous row block).  Shapes it does not take (unaligned rows, k > 128) and a
// null / short workspace run the single-kernel forms of kd_topk_logprobs: same results.
extern "C" int kd_topk_logprobs_ws(const void* logits, int dtype, int64_t R, int V, int64_t row_stride, int k, void* out_v,
                                   int32_t* out_i, void* workspace, size_t workspace_bytes, void* stream) {
  kd::DeviceGuard device_guard(logits);
  // KD_TOPK_FORM=cta / warp force a single-kernel form, =two the two-kernel form (A/B runs).  Default: the two-kernel
  // form for launches too small to give every row-warp slot of the warp form a row (its work items are row segments),
  // the warp form above that - measured on B200 at V = 152,936, k = 64 (tools/k3_direct.py): see DESIGN.md, K3.
  static int form_env = -1;
  if (form_env < 0) {
    const char* e = getenv("KD_TOPK_FORM");
    form_env = !e ? 0 : ((e[0] == 'c' || e[0] == 'w') ? 1 : (e[0] == 't' ? 2 : 0));
  }
  int form = form_env;
  if (form == 0) {
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    form = R < (int64_t)sms * kWrWarps ? 2 : 1;
  }
  const size_t es = dtype == KD_DTYPE_F32 ? 4 : 2;
  const bool vec_ok = logits && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (row_stride * es) % 16 == 0;
  const bool ws_ok = workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && R > 0 && V > 0 &&
                     workspace_bytes >= kd_topk_workspace_bytes(R, V);
  if (form == 2 && vec_ok && ws_ok && out_v && out_i && k > 0 && k <= V && k <= kWrCap / 2 && V >= 8) {
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    int rc = -1;
    switch (dtype) {
      case KD_DTYPE_F32:
        rc = topk_two_kernel<float>((const float*)logits, R, V, row_stride, k, (__half*)out_v, out_i, ws, s);
        break;
      case KD_DTYPE_BF16:
        rc = topk_two_kernel<__nv_bfloat16>((const __nv_bfloat16*)logits, R, V, row_stride, k, (__half*)out_v, out_i, ws, s);
        break;
      case KD_DTYPE_F16:
        rc = topk_two_kernel<__half>((const __half*)logits, R, V, row_stride, k, (__half*)out_v, out_i, ws, s);
        break;
      default:
        break;
    }
    if (rc >= 0) return rc;  // -1: this shape / device cannot take the form
  }
  return kd_topk_logprobs(logits, dtype, R, V, row_stride, k, out_v, out_i, stream);
}

extern "C" int kd_head_topk_select(const void* logits, int64_t row_stride, const void* pmax, int pmax_stride,
                                   const void* part, int part_stride, int n_part, int64_t R, int V, int k, void* out_v,
                                   int32_t* out_i, void* stream) {
  kd::DeviceGuard device_guard(logits);
  if (!logits || !pmax || !part || !out_v || !out_i) {
    set_error("kd_head_topk_select: null pointer argument");
    return 1;
  }
  if (R < 0 || V <= 0 || k <= 0 || k > V || k > kTopkMaxThreads || n_part <= 0 || n_part > part_stride ||
      pmax_stride < ((V + 255) / 256) * 8 || (pmax_stride & 7) != 0 || (reinterpret_cast<uintptr_t>(pmax) & 15) != 0) {
    set_error("kd_head_topk_select: bad shape (k=%d V=%d R=%lld n_part=%d pmax_stride=%d)", k, V, (long long)R, n_part,
              pmax_stride);
    return 1;
  }
  if (R == 0) return 0;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (row_stride * 2) % 16 == 0) ? 1 : 0;
  int grid = 0, npp = 0, nw = 0;
  size_t smem = 0;
  if (!vec_ok || k > kWrCap / 2 || !warp_form_config(kd_head_select_kernel<__nv_bfloat16>, R, V, &grid, &smem, &npp, &nw)) {
    // shapes the warp form does not take: the full-row compaction of the scratch logits gives the same result
    return kd_topk_logprobs(logits, KD_DTYPE_BF16, R, V, row_stride, k, out_v, out_i, stream);
  }
  kd_head_select_kernel<__nv_bfloat16><<<grid, 32 * nw, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)logits, R, V, row_stride, k, (const __nv_bfloat16*)pmax, pmax_stride, (const float2*)part,
      part_stride, n_part, (__half*)out_v, out_i, npp);
  return check_launch("kd_head_select launch");
}
