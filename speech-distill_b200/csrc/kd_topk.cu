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
constexpr int kWrPiece = 32;    // elements per piece (the head GEMM's epilogue produces the same pieces)

// Piece maxima live in shared memory as bf16 bit patterns rounded DOWN (exact for bf16 rows), so that counting
// "pieces >= t" is one packed hardware compare (HSET2.BF16) + one packed add per two pieces.
__device__ __forceinline__ uint16_t bf16_bits_rd(float x) {
  const __nv_bfloat16 b = __float2bfloat16_rd(x);
  return *reinterpret_cast<const uint16_t*>(&b);
}
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
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
template <typename T, typename F>
__device__ __forceinline__ void for_each_in_pieces(const T* __restrict__ row, int V, const uint16_t* pv, int n_pieces,
                                                   float tf, int lane, F&& fn) {
  for (int i = lane; i < n_pieces; i += 32) {
    if (!(bf16_bits_to_float(pv[i]) >= tf)) continue;
    const int c0 = i * kWrPiece;
    if (c0 + kWrPiece <= V) {
      Vec8<T> e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) e[q].load_global(row + c0 + 8 * q);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
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

template <typename T, typename P>
__device__ __forceinline__ int warp_count_if(const T* row, int V, const uint16_t* pv, int n_pieces, float tf, int lane,
                                             P&& pred) {
  int c = 0;
  for_each_in_pieces<T>(row, V, pv, n_pieces, tf, lane, [&](float x, int idx) { c += pred(order_key(x), idx) ? 1 : 0; });
  return __reduce_add_sync(0xffffffffu, c);
}

// threshold -> candidates -> (slow path) -> the k outputs of row r; lm / ll = row maximum and log of the exp sum
template <typename T>
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
      const int c0 = (int)w.plist[j] * kWrPiece;
      if (c0 + kWrPiece <= V) {
        Vec8<T> e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q].load_global(row + c0 + 8 * q);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
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
      const int c = warp_count_if<T>(row, V, w.pv, n_pieces, tf, lane, [&](uint32_t key, int) { return key >= mid; });
      if (c >= k) lo = mid; else hi = mid - 1;
    }
    const uint32_t kth = lo;
    const int above = warp_count_if<T>(row, V, w.pv, n_pieces, tf, lane, [&](uint32_t key, int) { return key > kth; });
    const int need = k - above;
    int jl = 0, jh = V - 1;
    while (jl < jh) {
      const int mid = jl + ((jh - jl) >> 1);
      const int c = warp_count_if<T>(row, V, w.pv, n_pieces, tf, lane,
                                     [&](uint32_t key, int idx) { return key == kth && idx <= mid; });
      if (c >= need) jh = mid; else jl = mid + 1;
    }
    const int jmax = jl;
    __syncwarp();
    if (lane == 0) w.count[0] = 0;
    __syncwarp();
    for_each_in_pieces<T>(row, V, w.pv, n_pieces, tf, lane, [&](float x, int idx) {
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
template <typename T, int U>
struct Pass1Batch {
  Vec8<T> v[U];
  __device__ __forceinline__ void load(const T* __restrict__ row, int base, int nvec, int lane) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 32 + lane;
      if (idx < nvec) v[u].load_global(row + (size_t)idx * 8);
    }
  }
  __device__ __forceinline__ void reduce(int base, int nvec, int lane, uint16_t* pv, float& m, float& s) {
    float vmx[U];
    float vm = -CUDART_INF_F;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * 32 + lane;
      float f[8];
      if (idx < nvec) {
        v[u].unpack(f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = -CUDART_INF_F;
      }
      float x = fmaxf(fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3])), fmaxf(fmaxf(f[4], f[5]), fmaxf(f[6], f[7])));
      vmx[u] = x;
      vm = fmaxf(vm, x);
      x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));  // four lanes = one 32-element piece
      x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
      if ((lane & 3) == 0 && idx < nvec) pv[idx >> 2] = bf16_bits_rd(x);
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
        if (vmx[u] == -CUDART_INF_F) continue;  // past the end of the row (or an all -inf vector): contributes 0
        float f[8];
        v[u].unpack(f);
        p0 += ex2(fmaf(f[0], kLog2e, -off)) + ex2(fmaf(f[4], kLog2e, -off));
        p1 += ex2(fmaf(f[1], kLog2e, -off)) + ex2(fmaf(f[5], kLog2e, -off));
        p2 += ex2(fmaf(f[2], kLog2e, -off)) + ex2(fmaf(f[6], kLog2e, -off));
        p3 += ex2(fmaf(f[3], kLog2e, -off)) + ex2(fmaf(f[7], kLog2e, -off));
      }
      s += (p0 + p1) + (p2 + p3);
    }
  }
};

template <typename T, int U>
__device__ __forceinline__ void warp_pass1(const T* __restrict__ row, int V, uint16_t* pv, int lane, float& lm, float& ll) {
  const int nvec = V >> 3;
  constexpr int kStep = 32 * U;
  float m = -CUDART_INF_F, s = 0.f;
  Pass1Batch<T, U> b0, b1;
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
    const int piece = tail0 / kWrPiece;
    // the tail opens a new piece when nvec is a multiple of 4, else it joins the last one
    const float prev = (nvec & 3) ? bf16_bits_to_float(pv[piece]) : -CUDART_INF_F;
    pv[piece] = bf16_bits_rd(fmaxf(prev, tmax));
  }
  const float wm = warp_max(m);
  const float ws = warp_sum(s * exp_diff(m, wm, kLog2e));
  lm = wm;
  ll = ln_acc(ws);
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(32 * kWrWarps) kd_topk_warp_kernel(const T* __restrict__ logits, int64_t R, int V,
                                                                       int64_t row_stride, int k,
                                                                       __half* __restrict__ out_v,
                                                                       int32_t* __restrict__ out_i, int n_pieces_pad) {
  extern __shared__ __align__(16) unsigned char wr_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WarpRow w = warp_row_smem(wr_smem, warp, n_pieces_pad);
  const int n_pieces = (V + kWrPiece - 1) / kWrPiece;
  for (int i = n_pieces + lane; i < n_pieces_pad; i += 32) w.pv[i] = kBf16NegInf;  // padding: never >= a threshold
  __syncwarp();
  for (int64_t r = (int64_t)blockIdx.x * kWrWarps + warp; r < R; r += (int64_t)gridDim.x * kWrWarps) {
    const T* row = logits + r * row_stride;
    float lm, ll;
    warp_pass1<T, (sizeof(T) == 2 ? 4 : 2)>(row, V, w.pv, lane, lm, ll);
    warp_select_emit<T>(row, V, k, w, n_pieces, n_pieces_pad, lm, ll, r, out_v, out_i, lane);
    __syncwarp();
  }
}

// ---- selection behind the teacher head GEMM (kd_head_logits_stats) -----------------------------------------------
// Same output as kd_topk_logprobs on the block's bf16 logits, without the sweep over the row: the head GEMM's epilogue
// left (a) the maximum of every 32-column piece and (b) partial (max, sum exp) records, so a row costs the merge of
// n_part records (~2 KB), the piece maxima (~10 KB) and the ~k pieces of 64 bytes that can hold a top-k entry.
__global__ void __launch_bounds__(32 * kWrWarps) kd_head_select_kernel(const __nv_bfloat16* __restrict__ logits, int64_t R,
                                                                         int V, int64_t row_stride, int k,
                                                                         const __nv_bfloat16* __restrict__ pmax,
                                                                         int pmax_stride, const float2* __restrict__ part,
                                                                         int part_stride, int n_part,
                                                                         __half* __restrict__ out_v,
                                                                         int32_t* __restrict__ out_i, int n_pieces_pad) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(16) unsigned char wr_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WarpRow w = warp_row_smem(wr_smem, warp, n_pieces_pad);
  const int n_pieces = (V + kWrPiece - 1) / kWrPiece;
  for (int64_t r = (int64_t)blockIdx.x * kWrWarps + warp; r < R; r += (int64_t)gridDim.x * kWrWarps) {
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
    warp_select_emit<T>(row, V, k, w, n_pieces, n_pieces_pad, lm, ll, r, out_v, out_i, lane);
    __syncwarp();
  }
}

}  // namespace kd

using namespace kd;

// launch geometry of the warp-per-row kernels: 8 row-warps per CTA, as many CTAs per SM as the shared memory holds
// (at most 4); returns false when a row's piece keys do not fit (V beyond ~3.4 M columns)
template <typename Kern>
static bool warp_form_config(Kern kern, int64_t R, int V, int* grid, size_t* smem, int* n_pieces_pad) {
  const int n_pieces = (V + kWrPiece - 1) / kWrPiece;
  *n_pieces_pad = (n_pieces + 7) & ~7;
  *smem = (size_t)kWrWarps * warp_row_bytes(*n_pieces_pad);
  if (*smem > 226 * 1024 || n_pieces > 65535) return false;  // piece indices are kept as 16-bit values
  int per_sm = (int)((size_t)232448 / (*smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ctas = (R + kWrWarps - 1) / kWrWarps;
  *grid = (int)(ctas < (int64_t)sms * per_sm ? ctas : (int64_t)sms * per_sm);
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}
// Which form takes a call of kd_topk_logprobs: a row is one warp's work in the warp form (~0.2 ms per row at
// V = 152,936), so it needs at least two rows for each of the 16 row-warps an SM holds to beat the CTA-per-row form,
// which spreads a row over 256 threads.  KD_TOPK_FORM=cta / warp forces one of them (A/B runs).
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
  return R >= (int64_t)sms * 2 * kWrWarps * 2;
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
    int grid = 0, npp = 0;
    size_t smem = 0;
    bool launched = false;
    switch (dtype) {
      case KD_DTYPE_F32:
        if (warp_form_config(kd_topk_warp_kernel<float>, R, V, &grid, &smem, &npp)) {
          kd_topk_warp_kernel<float><<<grid, 32 * kWrWarps, smem, s>>>((const float*)logits, R, V, row_stride, k,
                                                                       (__half*)out_v, out_i, npp);
          launched = true;
        }
        break;
      case KD_DTYPE_BF16:
        if (warp_form_config(kd_topk_warp_kernel<__nv_bfloat16>, R, V, &grid, &smem, &npp)) {
          kd_topk_warp_kernel<__nv_bfloat16><<<grid, 32 * kWrWarps, smem, s>>>((const __nv_bfloat16*)logits, R, V,
                                                                               row_stride, k, (__half*)out_v, out_i, npp);
          launched = true;
        }
        break;
      case KD_DTYPE_F16:
        if (warp_form_config(kd_topk_warp_kernel<__half>, R, V, &grid, &smem, &npp)) {
          kd_topk_warp_kernel<__half><<<grid, 32 * kWrWarps, smem, s>>>((const __half*)logits, R, V, row_stride, k,
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
  int grid = 0, npp = 0;
  size_t smem = 0;
  if (!vec_ok || k > kWrCap / 2 || !warp_form_config(kd_head_select_kernel, R, V, &grid, &smem, &npp)) {
    // shapes the warp form does not take: the full-row compaction of the scratch logits gives the same result
    return kd_topk_logprobs(logits, KD_DTYPE_BF16, R, V, row_stride, k, out_v, out_i, stream);
  }
  kd_head_select_kernel<<<grid, 32 * kWrWarps, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)logits, R, V, row_stride, k, (const __nv_bfloat16*)pmax, pmax_stride, (const float2*)part,
      part_stride, n_part, (__half*)out_v, out_i, npp);
  return check_launch("kd_head_select launch");
}
