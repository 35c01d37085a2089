// Read-bandwidth probe: what a read-only streaming kernel can reach on this GPU, by load mechanism.  The HBM figure in
// MEASURED_PEAKS.json is a copy (half reads, half writes); K2's forward and K3 are read-only streams, and their
// roofline fractions are read against this probe as well (bench.py `read_probe`, tools/read_probe.py).
//   mode 0: ld.global.nc.L1::no_allocate.v4 (what Vec8::load_global issues), `unroll` loads in flight per thread
//   mode 1: the same with an L2 evict_first policy
//   mode 2: cp.async.bulk (TMA, 1-D) global -> shared ring of `unroll` stages x 16 KB per CTA, one elected producer;
//           the consumers only wait on the stage's mbarrier and read one word (no LSU pressure from the stream itself)
#include "kd_common.cuh"
#include "kd_umma.cuh"

namespace kd {

template <int U, bool HINT>
__global__ void __launch_bounds__(256) kd_probe_ldg_kernel(const uint4* __restrict__ p, size_t n_vec, uint32_t* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  const uint64_t pol = l2_policy_evict_first();
  for (; i + (U - 1) * stride < n_vec; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = HINT ? ldg_hint(p + i + u * stride, pol) : ldg_stream(p + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n_vec; i += stride) {
    const uint4 v = ldg_stream(p + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x9e3779b9u) out[0] = acc;  // keeps the loads alive
}

constexpr uint32_t kProbeStage = 16384;

__global__ void __launch_bounds__(128) kd_probe_bulk_kernel(const uint8_t* __restrict__ p, size_t n_chunks, int stages,
                                                           uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t probe_smem[];
  const uint32_t base = umma::smem_u32(probe_smem);
  const uint32_t bars = base + (uint32_t)stages * kProbeStage;  // full[stages], empty[stages]
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      umma::mbar_init(bars + 8u * s, 1);
      umma::mbar_init(bars + 8u * (stages + s), 96);
    }
    umma::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        umma::mbar_wait(bars + 8u * (stages + s), ph ^ 1u);
        umma::mbar_expect_tx(bars + 8u * s, kProbeStage);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         base + (uint32_t)s * kProbeStage),
                     "l"(p + c * kProbeStage), "r"(kProbeStage), "r"(bars + 8u * s)
                     : "memory");
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    int s = 0;
    uint32_t ph = 0, acc = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      umma::mbar_wait(bars + 8u * s, ph);
      acc ^= *reinterpret_cast<const volatile uint32_t*>(probe_smem + (size_t)s * kProbeStage + 4 * threadIdx.x);
      umma::mbar_arrive(bars + 8u * (stages + s));
      if (++s == stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (acc == 0x9e3779b9u) out[0] = acc;
  }
}

}  // namespace kd

using namespace kd;

extern "C" int kd_probe_read_bandwidth(const void* p, size_t bytes, int mode, int ctas_per_sm, int unroll, void* scratch4,
                                       void* stream) {
  kd::DeviceGuard device_guard(p);
  if (!p || !scratch4 || bytes < (1u << 20) || (reinterpret_cast<uintptr_t>(p) & 15) != 0) {
    set_error("kd_probe_read_bandwidth: need a 16-byte aligned buffer of at least 1 MiB and a 4-byte scratch");
    return 1;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t s = (cudaStream_t)stream;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (mode == 2) {
    int stages = unroll < 2 ? 2 : (unroll > 12 ? 12 : unroll);
    const size_t smem = (size_t)stages * kProbeStage + 16 * stages + 64;
    if (ctas_per_sm * (smem + 1024) > 232448) ctas_per_sm = (int)(232448 / (smem + 1024));
    if (check_cuda(cudaFuncSetAttribute(kd_probe_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                   "probe smem"))
      return 2;
    kd_probe_bulk_kernel<<<sms * ctas_per_sm, 128, smem, s>>>((const uint8_t*)p, bytes / kProbeStage, stages,
                                                              (uint32_t*)scratch4);
    return check_launch("kd_probe_bulk launch");
  }
  const size_t n_vec = bytes / 16;
  const int grid = sms * ctas_per_sm;
#define KD_PROBE(U)                                                                                        \
  do {                                                                                                     \
    if (mode == 1) kd_probe_ldg_kernel<U, true><<<grid, 256, 0, s>>>((const uint4*)p, n_vec, (uint32_t*)scratch4); \
    else kd_probe_ldg_kernel<U, false><<<grid, 256, 0, s>>>((const uint4*)p, n_vec, (uint32_t*)scratch4);  \
  } while (0)
  if (unroll >= 16) KD_PROBE(16);
  else if (unroll >= 8) KD_PROBE(8);
  else if (unroll >= 4) KD_PROBE(4);
  else KD_PROBE(1);
#undef KD_PROBE
  return check_launch("kd_probe_ldg launch");
}
