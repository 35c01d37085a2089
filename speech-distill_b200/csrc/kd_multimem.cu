// Token-shard dW all-reduce through the NVSwitch (NVLS multicast) instead of NCCL's ring: one small kernel per
// vocabulary range and GPU, beside the backward's GEMMs.
//
// The LM-head gradient of every rank lives in a symmetric buffer (same layout on all GPUs, one multicast address that
// maps to all replicas).  For a finished row range, rank g owns 1/G of the rows: it reads them with
// multimem.ld_reduce (the switch adds the G replicas in fp32 and returns one value: 1/G of the range enters this GPU
// once, already reduced) and writes the sum back with multimem.st (the switch stores it into every replica).  Per GPU
// the kernel moves 2/G of the range through its SMs - a ring all-reduce moves 2 (G-1)/G of it, 7x as much at G = 8 -
// so a handful of CTAs do in ~0.1 ms what NCCL's 32 CTAs held 32 SMs ~1 ms for.
// The cross-GPU barriers before (all replicas written) and after (all sums stored) are the caller's
// (torch's symmetric-memory handle: hdl.barrier on the same stream).
// Reference semantics: the SUM of train.py's gradient synchronisation (accelerate / DDP all-reduce of lm_head.weight.grad).
#include "kd_common.cuh"

namespace kd {

__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ uint4 multimem_ld_reduce_f32x4(const void* mc) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_16(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

constexpr int kMmThreads = 512;
constexpr int kMmUnroll = 4;

// mc = multicast address of the first byte of this rank's slice, n_vec = 16-byte vectors in the slice
template <bool F32>
__global__ void __launch_bounds__(kMmThreads) kd_multimem_allreduce_kernel(uint8_t* __restrict__ mc, size_t n_vec) {
  const size_t stride = (size_t)gridDim.x * kMmThreads;
  size_t i = (size_t)blockIdx.x * kMmThreads + threadIdx.x;
  for (; i + (kMmUnroll - 1) * stride < n_vec; i += kMmUnroll * stride) {
    uint4 v[kMmUnroll];
#pragma unroll
    for (int u = 0; u < kMmUnroll; ++u)
      v[u] = F32 ? multimem_ld_reduce_f32x4(mc + (i + u * stride) * 16) : multimem_ld_reduce_bf16x8(mc + (i + u * stride) * 16);
#pragma unroll
    for (int u = 0; u < kMmUnroll; ++u) multimem_st_16(mc + (i + u * stride) * 16, v[u]);
  }
  for (; i < n_vec; i += stride) {
    const uint4 v = F32 ? multimem_ld_reduce_f32x4(mc + i * 16) : multimem_ld_reduce_bf16x8(mc + i * 16);
    multimem_st_16(mc + i * 16, v);
  }
}

}  // namespace kd

using namespace kd;

extern "C" int kd_multimem_allreduce(void* multicast_base, size_t byte_offset, size_t bytes, int dtype, int rank,
                                     int world, int ctas, void* stream) {
  if (!multicast_base || world < 1 || rank < 0 || rank >= world || (byte_offset & 15) != 0 || (bytes & 15) != 0 ||
      (dtype != KD_DTYPE_BF16 && dtype != KD_DTYPE_F32)) {
    set_error("kd_multimem_allreduce: need a multicast address, 16-byte aligned offset / size, bf16 or fp32, 0 <= rank < world");
    return 1;
  }
  if (bytes == 0) return 0;
  // this rank's share: whole 16-byte vectors, the last rank takes the remainder
  const size_t n_vec_all = bytes / 16;
  const size_t per = (n_vec_all + world - 1) / world;
  const size_t v0 = per * rank < n_vec_all ? per * rank : n_vec_all;
  const size_t v1 = v0 + per < n_vec_all ? v0 + per : n_vec_all;
  if (v1 <= v0) return 0;
  if (ctas < 1) ctas = 16;
  uint8_t* mc = reinterpret_cast<uint8_t*>(multicast_base) + byte_offset + v0 * 16;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == KD_DTYPE_F32) kd_multimem_allreduce_kernel<true><<<ctas, kMmThreads, 0, s>>>(mc, v1 - v0);
  else kd_multimem_allreduce_kernel<false><<<ctas, kMmThreads, 0, s>>>(mc, v1 - v0);
  return check_launch("kd_multimem_allreduce launch");
}
