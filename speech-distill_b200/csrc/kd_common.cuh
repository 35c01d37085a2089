// Shared device helpers for the KD hot-path kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/kd_b200.h"

namespace kd {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kNumPartialSlots = 8;  // floats per partial-sum record

// error plumbing (kd_api.cu)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
// cudaGetLastError() after a kernel launch; also counts the launch (kd_launch_count(), bench.py's gpu_launches)
int check_launch(const char* what);

constexpr int kMaxDevices = 64;
// index of the current device for per-device caches (function attributes, SM counts, stream pools)
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

// Every entry point launches on the device that OWNS its tensors, whatever the calling thread's current device is
// (single-process multi-GPU callers: a teacher on another GPU, one thread per device): the guard looks the device
// of a primary pointer up, switches to it for the duration of the call and restores the caller's device after.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    if (p == nullptr) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();  // not a CUDA pointer: the entry point's own checks report it
      return;
    }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return;
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    if (a.device != prev && cudaSetDevice(a.device) == cudaSuccess) switched = true;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// natural log with full fp32 accuracy for the handful of per-row uses
__device__ __forceinline__ float ln_acc(float x) { return logf(x); }

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kCode = KD_DTYPE_F32;
  __device__ static float to_f(float v) { return v; }
  __device__ static float from_f(float v) { return v; }
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kCode = KD_DTYPE_BF16;
  __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Elem<__half> {
  static constexpr int kCode = KD_DTYPE_F16;
  __device__ static float to_f(__half v) { return __half2float(v); }
  __device__ static __half from_f(float v) { return __float2half_rn(v); }
};

// ---- 8-element vector access (16 B for 16-bit types, 2 x 16 B for fp32) ------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// L2 eviction policies (createpolicy) for the two-sweep row kernel of K2: the first sweep marks a row's lines
// evict_last so they survive until the second sweep, which reads them evict_first (dead afterwards)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg_hint(const void* p, uint64_t policy) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

template <typename T>
struct Vec8;

template <>
struct Vec8<float> {
  uint4 a, b;
  __device__ __forceinline__ void load_global(const float* p) {
    a = ldg_stream(p);
    b = ldg_stream(p + 4);
  }
  __device__ __forceinline__ void load_global_hint(const float* p, uint64_t pol) {
    a = ldg_hint(p, pol);
    b = ldg_hint(p + 4, pol);
  }
  __device__ __forceinline__ void load_shared(const float* p) {
    a = *reinterpret_cast<const uint4*>(p);
    b = *reinterpret_cast<const uint4*>(p + 4);
  }
  __device__ __forceinline__ void store_shared(float* p) const {
    *reinterpret_cast<uint4*>(p) = a;
    *reinterpret_cast<uint4*>(p + 4) = b;
  }
  __device__ __forceinline__ void store_global(float* p) const {
    stg_stream(p, a);
    stg_stream(p + 4, b);
  }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    a = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    b = make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
  }
};

template <>
struct Vec8<__nv_bfloat16> {
  uint4 a;
  __device__ __forceinline__ void load_global(const __nv_bfloat16* p) { a = ldg_stream(p); }
  __device__ __forceinline__ void load_global_hint(const __nv_bfloat16* p, uint64_t pol) { a = ldg_hint(p, pol); }
  __device__ __forceinline__ void load_shared(const __nv_bfloat16* p) { a = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store_shared(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = a; }
  __device__ __forceinline__ void store_global(__nv_bfloat16* p) const { stg_stream(p, a); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    a = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <>
struct Vec8<__half> {
  uint4 a;
  __device__ __forceinline__ void load_global(const __half* p) { a = ldg_stream(p); }
  __device__ __forceinline__ void load_global_hint(const __half* p, uint64_t pol) { a = ldg_hint(p, pol); }
  __device__ __forceinline__ void load_shared(const __half* p) { a = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store_shared(__half* p) const { *reinterpret_cast<uint4*>(p) = a; }
  __device__ __forceinline__ void store_global(__half* p) const { stg_stream(p, a); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    a = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Floor for -inf teacher logits where the unguarded cross term is used.  A POWER OF TWO on purpose: while a thread has
// seen nothing but floored entries the floor is its running maximum, and the exponent's argument
// fma(y, c, -(max * log2e) / tau) must then be exactly 0 - with -1e30 the rounding of max * log2e left +-8e22 there,
// i.e. e^x = inf, and inf * 0 = NaN at the next rescale (a row whose first 16 teacher columns are all -inf).
constexpr float kTeacherFloor = -0x1p100f;  // -1.27e30
// Two packed bf16 values with -inf (0xFF80) replaced by the floor (bf16 0xF180): for negative floats the bit pattern
// grows with the magnitude, positives have a clear sign bit and stay below the bound, so one unsigned 16x2 minimum
// does it (p = e^{(y - max)/tau} is still exactly 0 for such an entry once a real logit has been seen, and
// 0 * (y - z) is then 0 instead of NaN - the xlogy convention of nn.KLDivLoss, distillation_loss.py:68).
__device__ __forceinline__ uint32_t clamp_neg_inf_bf16x2(uint32_t w) { return __vminu2(w, 0xF180F180u); }

// ---- warp helpers --------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// exp(a - b) with the (-inf) - (-inf) = 0 convention used when merging empty partials
__device__ __forceinline__ float exp_diff(float a, float b, float scale_log2e) {
  return (a == b) ? 1.0f : ex2((a - b) * scale_log2e);
}


// ---- online soft-max statistics ------------------------------------------
// student: m = running max, s1 = sum e^{z-m}, st = sum e^{(z-m)/tau}
// teacher: mt, t1, tt likewise and a = sum e^{(y-mt)/tau} (y - z)
template <bool TAU2>
struct ExpPair {
  // e^{(x-m)/tau} and e^{x-m}; tau == 2 needs one MUFU: e^{x-m} = (e^{(x-m)/2})^2
  __device__ static __forceinline__ void eval(float x, float c_tau, float off_tau, float off_one, float& e_tau,
                                              float& e_one) {
    e_tau = ex2(fmaf(x, c_tau, -off_tau));
    if (TAU2) {
      e_one = e_tau * e_tau;
    } else {
      e_one = ex2(fmaf(x, kLog2e, -off_one));
    }
  }
};

// FSQ (tau == 2 only): the tau = 1 sum takes e^{x-m} = et^2 through one FMA (p1 = et * et + p1) instead of a
// multiply and an add - one instruction less per element for the issue-bound streaming kernel (K2)
template <bool TAU2, int N, bool FSQ = false>
__device__ __forceinline__ void student_update(const float (&f)[N], int nvalid, float inv_tau, float& m, float& s1,
                                               float& st) {
  float vm = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < nvalid) vm = fmaxf(vm, f[i]);
  if (vm > m) {
    s1 *= exp_diff(m, vm, kLog2e);
    st *= exp_diff(m, vm, kLog2e * inv_tau);
    m = vm;
  }
  if (m == -CUDART_INF_F) return;  // nothing finite yet
  const float c_tau = kLog2e * inv_tau;
  const float off_one = m * kLog2e, off_tau = off_one * inv_tau;
  // two partial sums per statistic: short dependency chains for the 2 epilogue warps per scheduler
  float pt[2] = {0.f, 0.f}, p1[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < nvalid) {
      float et, e1;
      ExpPair<TAU2>::eval(f[i], c_tau, off_tau, off_one, et, e1);
      pt[i & 1] += et;
      if (TAU2 && FSQ) p1[i & 1] = fmaf(et, et, p1[i & 1]);
      else p1[i & 1] += e1;
    }
  }
  st += pt[0] + pt[1];
  s1 += p1[0] + p1[1];
}

// The two halves of student_update for callers that already know a bound vm >= every value they are about to add
// (the fused forward takes the maximum of a 32-column piece once, for the statistics and for the logit cache alike):
// raise the running maximum to vm, then add values without looking for a new maximum.
__device__ __forceinline__ void student_raise_max(float vm, float inv_tau, float& m, float& s1, float& st) {
  if (vm > m) {
    s1 *= exp_diff(m, vm, kLog2e);
    st *= exp_diff(m, vm, kLog2e * inv_tau);
    m = vm;
  }
}
template <bool TAU2, int N>
__device__ __forceinline__ void student_add(const float (&f)[N], int nvalid, float inv_tau, float m, float& s1,
                                            float& st) {
  if (m == -CUDART_INF_F) return;
  const float c_tau = kLog2e * inv_tau;
  const float off_one = m * kLog2e, off_tau = off_one * inv_tau;
  float pt[2] = {0.f, 0.f}, p1[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < nvalid) {
      float et, e1;
      ExpPair<TAU2>::eval(f[i], c_tau, off_tau, off_one, et, e1);
      pt[i & 1] += et;
      p1[i & 1] += e1;
    }
  }
  st += pt[0] + pt[1];
  s1 += p1[0] + p1[1];
}

// GUARD = 1 keeps p == 0 terms at exactly 0 whatever the student logit is (K1's ragged vocabulary edge: the padding
// columns hold -inf on both sides).  GUARD = 2 (K2, user-supplied logits) does so for finite student logits only: a
// -inf student logit contributes what it does in the reference - inf where the teacher has mass, NaN (0 * inf inside
// kl_div) where it has none.  GUARD = 0 (fused path and K2's hot loop: teacher values already clamped to >=
// kTeacherFloor by the caller, see clamp_neg_inf_bf16x2) runs the cross term as one subtract and one FMA per element,
// with the same results as GUARD = 2.
template <bool TAU2, int N, int GUARD = 1, bool FSQ = false>
__device__ __forceinline__ void teacher_update(const float (&fy)[N], const float (&fz)[N], int nvalid, float inv_tau,
                                               float& mt, float& t1, float& tt, float& a) {
  float vm = -CUDART_INF_F;
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < nvalid) vm = fmaxf(vm, fy[i]);
  if (vm > mt) {
    const float r = exp_diff(mt, vm, kLog2e * inv_tau);
    t1 *= exp_diff(mt, vm, kLog2e);
    tt *= r;
    a *= r;
    mt = vm;
  }
  if (mt == -CUDART_INF_F) return;
  const float c_tau = kLog2e * inv_tau;
  const float off_one = mt * kLog2e, off_tau = off_one * inv_tau;
  float pt[2] = {0.f, 0.f}, p1[2] = {0.f, 0.f}, pa[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < N; ++i) {
    if (i < nvalid) {
      float et, e1;
      ExpPair<TAU2>::eval(fy[i], c_tau, off_tau, off_one, et, e1);
      pt[i & 1] += et;
      if (TAU2 && FSQ) p1[i & 1] = fmaf(et, et, p1[i & 1]);
      else p1[i & 1] += e1;
      if (GUARD) {
        // p = 0 contributes exactly 0 (xlogy semantics of nn.KLDivLoss, distillation_loss.py:68);
        // the clamp keeps 0 * (-inf - z) from producing NaN when the teacher holds -inf
        const float d = fmaxf(fy[i], kTeacherFloor) - fz[i];
        const bool take = GUARD == 2 ? (et > 0.f || fz[i] == -CUDART_INF_F) : (et > 0.f);
        pa[i & 1] = take ? fmaf(et, d, pa[i & 1]) : pa[i & 1];
      } else {
        pa[i & 1] = fmaf(et, fy[i] - fz[i], pa[i & 1]);
      }
    }
  }
  const float pa_sum = pa[0] + pa[1];
  tt += pt[0] + pt[1];
  t1 += p1[0] + p1[1];
  a += pa_sum;
}

__device__ __forceinline__ void merge_student(float& m, float& s1, float& st, float m2, float s12, float st2,
                                              float inv_tau) {
  const float mm = fmaxf(m, m2);
  s1 = s1 * exp_diff(m, mm, kLog2e) + s12 * exp_diff(m2, mm, kLog2e);
  st = st * exp_diff(m, mm, kLog2e * inv_tau) + st2 * exp_diff(m2, mm, kLog2e * inv_tau);
  m = mm;
}
__device__ __forceinline__ void merge_teacher(float& m, float& t1, float& tt, float& a, float m2, float t12,
                                              float tt2, float a2, float inv_tau) {
  const float mm = fmaxf(m, m2);
  const float ra = exp_diff(m, mm, kLog2e * inv_tau), rb = exp_diff(m2, mm, kLog2e * inv_tau);
  t1 = t1 * exp_diff(m, mm, kLog2e) + t12 * exp_diff(m2, mm, kLog2e);
  tt = tt * ra + tt2 * rb;
  a = a * ra + a2 * rb;
  m = mm;
}

// deterministic fixed-order reduction of [n][kNumPartialSlots] partial records -> sums[8] (host launcher, kd_stream.cu)
int reduce_partials(const float* partials, int n, float* sums, cudaStream_t stream);

// row predicate of distillation_loss.py:34-41: row (b,t) scores labels[b,t+1]
__device__ __forceinline__ bool row_is_valid(const int64_t* __restrict__ labels, const uint8_t* __restrict__ mask,
                                             int T, int b, int t, int64_t ignore_index, int64_t* label_out) {
  if (t >= T - 1) return false;
  const int64_t idx = (int64_t)b * T + t + 1;
  const int64_t l = labels[idx];
  *label_out = l;
  if (l == ignore_index) return false;
  if (mask != nullptr && mask[idx] == 0) return false;
  return true;
}

}  // namespace kd
