// Debug probe: tcgen05.mma with an MN-major (pixels contiguous) SWIZZLE_128B A operand written by threads,
// K-major SWIZZLE_128B B operand, fp16 x fp16 -> fp32.  Checks the shared-memory layout / descriptor fields
// (LBO = stride between 64-element M atoms, SBO = stride between 8-row K groups) that the channel-per-lane
// fused block uses for its depthwise output, against a host matmul (tools/mn_probe.py).
#include "../ptx_sm100.cuh"
#include "../../../include/audiomuse_b200_debug.h"

namespace am {
using namespace ptx;

// A element (m, k) of a [128 x K] fp16 matrix, MN-major SW128: atoms of 64 m x 8 k (1024 B), rows of an atom are
// the 8 k values (128 B = 64 m each), 16-byte chunks (8 m) XOR-swizzled with the row
__device__ __forceinline__ uint32_t mn_offset(int m, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(m >> 6) * lbo + (uint32_t)(k >> 3) * sbo + (uint32_t)(k & 7) * 128u +
         ((((uint32_t)(m >> 3) & 7u) ^ ((uint32_t)k & 7u)) << 4) + (uint32_t)(m & 7) * 2u;
}

__global__ void __launch_bounds__(128, 1)
mn_probe_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, int N, int K, uint32_t lbo, uint32_t sbo,
                int m_rows, int swap, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;               // 32 KB
  uint8_t* sB = smem + 32768;       // K/64 k-blocks of [N rows x 128 B]
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32768 + 2 * 256 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < m_rows * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    *reinterpret_cast<uint16_t*>(sA + mn_offset(m, k, lbo, sbo)) = A[i];
  }
  const int nb_bytes = N * 128;
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    *reinterpret_cast<uint16_t*>(sB + (k >> 6) * nb_bytes + sw128_offset((uint32_t)n, (uint32_t)((k & 63) >> 3)) + (k & 7) * 2) = B[i];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    if (elect_one_sync()) {
      // fp16 operands, A MN-major (bit 15), B K-major
      // swap bit 1: fp16 ACCUMULATORS (D format field 0 instead of 1) -- how does TMEM hold them?
      const uint32_t idesc = ((swap & 2) ? (make_idesc_f16(128, N) & ~(1u << 4)) : make_idesc_f16(128, N)) | (1u << 15);
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)ks * 2u * sbo;
        uint64_t da = 0;
        da |= (uint64_t)((a_addr >> 4) & 0x3fffu);
        da |= (uint64_t)((((swap & 1) ? sbo : lbo) >> 4) & 0x3fffu) << 16;   // swap bit 0: the two stride fields exchanged
        da |= (uint64_t)((((swap & 1) ? lbo : sbo) >> 4) & 0x3fffu) << 32;
        da |= (uint64_t)1 << 46;
        da |= (uint64_t)2 << 61;
        const uint64_t db = make_smem_desc(smem_u32(sB) + (uint32_t)((ks >> 2) * nb_bytes) + (uint32_t)((ks & 3) * 32));
        umma_f16(tmem, da, db, idesc, (uint32_t)(ks > 0));
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    const int m = warp * 32 + (tid & 31);
    for (int e = 0; e < 16; ++e) D[(size_t)m * N + c0 + e] = __uint_as_float(v[e]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}
}  // namespace am

extern "C" AM_API int am_probe_mn_major(const uint16_t* a_f16, const uint16_t* b_f16, int N, int K, int lbo_bytes,
                                        int sbo_bytes, int m_rows, int swap, float* d_out) {
  using namespace am;
  AM_CHECK(a_f16 && b_f16 && d_out && N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K <= 128 && K % 16 == 0 &&
               (m_rows == 64 || m_rows == 128) && lbo_bytes >= 0 && sbo_bytes > 0,
           "am_probe_mn_major: bad argument");
  AM_CHECK((size_t)(m_rows > 64 ? lbo_bytes : 0) + (size_t)(K / 8) * sbo_bytes <= 32768 + 1024, "am_probe_mn_major: layout exceeds 32 KB");
  AM_TRY(ensure_init());
  DevBuf<uint16_t> dA, dB;
  DevBuf<float> dD;
  AM_TRY(dA.alloc((size_t)m_rows * K));
  AM_TRY(dB.alloc((size_t)N * K));
  AM_TRY(dD.alloc((size_t)128 * N));
  AM_CUDA(cudaMemcpy(dA.p, a_f16, (size_t)m_rows * K * 2, cudaMemcpyHostToDevice));
  AM_CUDA(cudaMemcpy(dB.p, b_f16, (size_t)N * K * 2, cudaMemcpyHostToDevice));
  const size_t smem = 32768 + 2 * 256 * 128 + 1024;
  AM_CUDA(cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mn_probe_kernel<<<1, 128, smem>>>(dA.p, dB.p, N, K, (uint32_t)lbo_bytes, (uint32_t)sbo_bytes, m_rows, swap, dD.p);
  AM_CUDA(cudaGetLastError());
  AM_CUDA(cudaDeviceSynchronize());
  AM_CUDA(cudaMemcpy(d_out, dD.p, (size_t)128 * N * 4, cudaMemcpyDeviceToHost));
  return AM_OK;
}
