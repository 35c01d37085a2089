// Valid-row compaction for K1.  The collator marks the text prefix and the padding of every
// sequence with -100 (data.py:246-251, 273-276, 350-387), typically a large share of the rows; the reference
// drops them with a boolean gather (distillation_loss.py:37-45).  Here the same thing happens on the device
// without a host sync: valid rows are moved to the front (order preserved), the GEMM kernels read the live-row
// count from device memory and skip every tile behind it, and dH is scattered back through the inverse map.
#include "kd_common.cuh"

namespace kd {

constexpr int kCompactThreads = 1024;

// perm[j] = original row of the j-th valid row (j < N), -1 behind; inv[r] = rank of row r among the valid rows
// or -1; target_c[j] = row_target[perm[j]] (-1 behind).  One CTA: R is a few tens of thousands at most.
__global__ void __launch_bounds__(kCompactThreads) kd_compact_rows_kernel(const int32_t* __restrict__ row_target, int R,
                                                                         int32_t* __restrict__ perm,
                                                                         int32_t* __restrict__ inv,
                                                                         int32_t* __restrict__ target_c,
                                                                         int32_t* __restrict__ n_valid) {
  __shared__ int warp_tot[kCompactThreads / 32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int r0 = 0; r0 < R; r0 += kCompactThreads) {
    const int r = r0 + tid;
    const int t = r < R ? row_target[r] : -1;
    const bool ok = t >= 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    const int before = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (r < R) {
      if (ok) {
        const int j = off + before;
        perm[j] = r;
        target_c[j] = t;
        inv[r] = j;
      } else {
        inv[r] = -1;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int s = 0;
      for (int w = 0; w < kCompactThreads / 32; ++w) s += warp_tot[w];
      base += s;
    }
    __syncthreads();
  }
  const int n = base;
  for (int j = n + tid; j < R; j += kCompactThreads) {
    perm[j] = -1;
    target_c[j] = -1;
  }
  if (tid == 0 && n_valid) *n_valid = n;
}

// dst[j, :] = src[map[j], :] for map[j] >= 0, else zeros (zero_fill) or untouched.  Rows are row_bytes bytes,
// copied in 16-byte pieces when everything is aligned.  grid.x = rows, grid.y splits long rows.
__global__ void __launch_bounds__(256) kd_gather_rows_kernel(const uint8_t* __restrict__ src, int64_t src_stride,
                                                            const int32_t* __restrict__ map, int R,
                                                            uint8_t* __restrict__ dst, int64_t dst_stride,
                                                            int64_t row_bytes, int zero_fill, int vec_ok) {
  for (int j = blockIdx.x; j < R; j += gridDim.x) {
    const int s = map[j];
    if (s < 0 && !zero_fill) continue;
    uint8_t* d = dst + (int64_t)j * dst_stride;
    const uint8_t* p = s >= 0 ? src + (int64_t)s * src_stride : nullptr;
    if (vec_ok) {
      const int64_t n16 = row_bytes >> 4;
      for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.y * blockDim.x) {
        const uint4 v = p ? ldg_stream(p + (i << 4)) : make_uint4(0, 0, 0, 0);
        stg_stream(d + (i << 4), v);
      }
    } else {
      for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < row_bytes; i += (int64_t)gridDim.y * blockDim.x)
        d[i] = p ? p[i] : (uint8_t)0;
    }
  }
}

// zero `n16` 16-byte pieces at `dst` iff *n_rows == 0: with no live row the dW GEMM has nothing to contract and is
// skipped, but the caller still expects a zero gradient (distillation_loss.py:47-53); otherwise exit at once
__global__ void __launch_bounds__(256) kd_zero_if_empty_kernel(uint4* __restrict__ dst, int64_t n16,
                                                              const int32_t* __restrict__ n_rows) {
  if (*n_rows > 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace kd

using namespace kd;

extern "C" int kd_zero_if_empty(void* dst, int64_t bytes, const int32_t* n_rows, void* stream) {
  kd::DeviceGuard device_guard(dst);
  if (!dst || !n_rows || bytes < 0 || (bytes & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0) {
    set_error("kd_zero_if_empty: bad arguments (16-byte aligned buffer and size required)");
    return 1;
  }
  if (bytes == 0) return 0;
  kd_zero_if_empty_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(dst), bytes >> 4, n_rows);
  return check_launch("kd_zero_if_empty launch");
}

extern "C" int kd_compact_rows(const int32_t* row_target, int R, int32_t* perm, int32_t* inv, int32_t* target_c,
                               int32_t* n_valid, void* stream) {
  kd::DeviceGuard device_guard(row_target);
  if (!row_target || !perm || !inv || !target_c || R <= 0) {
    set_error("kd_compact_rows: bad arguments");
    return 1;
  }
  kd_compact_rows_kernel<<<1, kCompactThreads, 0, (cudaStream_t)stream>>>(row_target, R, perm, inv, target_c, n_valid);
  return check_launch("kd_compact_rows launch");
}

extern "C" int kd_gather_rows(const void* src, int64_t src_stride_bytes, const int32_t* map, int R, void* dst,
                              int64_t dst_stride_bytes, int64_t row_bytes, int zero_fill, void* stream) {
  kd::DeviceGuard device_guard(src);
  if (!src || !map || !dst || R <= 0 || row_bytes <= 0) {
    set_error("kd_gather_rows: bad arguments");
    return 1;
  }
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)src_stride_bytes |
                        (uintptr_t)dst_stride_bytes | (uintptr_t)row_bytes) & 15) == 0;
  int gy = (int)((row_bytes / 16 + 256 * 8 - 1) / (256 * 8));  // ~8 pieces per thread
  if (gy < 1) gy = 1;
  if (gy > 64) gy = 64;
  dim3 grid(R < 65535 ? R : 65535, gy);
  kd_gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint8_t*>(src), src_stride_bytes, map, R, reinterpret_cast<uint8_t*>(dst),
      dst_stride_bytes, row_bytes, zero_fill, vec_ok ? 1 : 0);
  return check_launch("kd_gather_rows launch");
}
