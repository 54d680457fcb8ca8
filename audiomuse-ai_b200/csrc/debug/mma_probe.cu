// Debug probe: issue cost of tcgen05.mma (kind::f16, cta_group::1, M = 128, K = 16) from K-major
// SWIZZLE_128B shared-memory operands, alone or under shared-memory traffic from 16 other warps.
// Answers "what does one MMA of the fused block cost" without the rest of that kernel around it.
#include "../ptx_sm100.cuh"
#include "../../../include/audiomuse_b200_debug.h"

namespace am {
using namespace ptx;

constexpr int kProbeThreads = 17 * 32;

__global__ void __launch_bounds__(kProbeThreads, 1)
mma_probe_kernel(int N, int iters, int d_tiles, int traffic, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  uint8_t* sA = smem;                // 3 x 16 KB: three M tiles like the fused block's X halo
  uint8_t* sB = smem + 3 * 16384;    // N rows x 128 B
  uint8_t* sT = sB + 32768;          // 64 KB traffic scratch
  for (int i = threadIdx.x; i < (3 * 16384 + 32768 + 65536) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    done = 0;
  }
  if (warp == 16) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 16) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(128, N);
      const uint64_t da = make_smem_desc(smem_u32(sA)), db = make_smem_desc(smem_u32(sB));
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int t = (i >> 2) % d_tiles;  // 4 k-steps into one accumulator, then the next tile
        umma_f16(tmem + (uint32_t)(t * N), da + (uint64_t)(t * 1024) + (uint64_t)((i & 3) * 2),
                 db + (uint64_t)((i & 3) * 2), idesc, (uint32_t)(i >= 4 * d_tiles));
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
      done = 1;
    }
    __syncwarp();
  } else if (traffic) {
    // 16 warps hammering shared memory: LDS.128 + STS.128 on conflict-free addresses
    const uint32_t base = smem_u32(sT) + (uint32_t)threadIdx.x * 16u;
    uint4 acc = make_uint4(0, 0, 0, 0);
    int guard = 0;
    while (!done && guard < (1 << 22)) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(base + (uint32_t)(r * 8192)));
        acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
        if (traffic > 1)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)(r * 8192)), "r"(acc.x),
                       "r"(acc.y), "r"(acc.z), "r"(acc.w));
      }
      ++guard;
    }
    if (acc.x == 0x12345u) out[0] = acc.y;  // keep the loop alive
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// TMEM read bandwidth: `warps` warps (4 per lane group first) each issue `iters` tcgen05.ld of 32 lanes x
// `cols` columns, `depth` loads in flight before each wait.  out[cta] = cycles for the whole loop.
template <int kCols, int kDepth>
__global__ void __launch_bounds__(544, 1) tmem_ld_probe_kernel(int iters, int n_mma, long long* __restrict__ out) {
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (n_mma > 0) {
    for (int i = threadIdx.x; i < (16384 + 8192) / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_barrier_init();
    }
    fence_proxy_async();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 16) {  // optional: an MMA stream (M128 x N64 x K16, accumulating into columns 448..511) beside the loads
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(128, 64);
      const uint64_t da = make_smem_desc(smem_u32(smem)), db = make_smem_desc(smem_u32(smem + 16384));
#pragma unroll 4
      for (int i = 0; i < n_mma; ++i)
        umma_f16(tmem_slot + 448u, da + (uint64_t)((i & 3) * 2), db + (uint64_t)((i & 3) * 2), idesc, (uint32_t)(i > 0));
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[gridDim.x + blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  } else
  for (int i = 0; i < iters; i += kDepth) {
    uint32_t v[kDepth][kCols];
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      {
        const uint32_t col = (uint32_t)(((i + d) * kCols + (warp >> 2) * 64) & 511) & ~(uint32_t)(kCols - 1);
        if constexpr (kCols == 16) tmem_ld_x16(base + col, v[d]);
        else tmem_ld_x32(base + col, v[d]);
      }
    }
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
#pragma unroll
      for (int e = 0; e < kCols; ++e) acc ^= v[d][e];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x1234567u) out[0] = 0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}
}  // namespace am

extern "C" AM_API int am_probe_mma(int N, int iters, int d_tiles, int traffic, double* issue_cycles,
                                   double* total_cycles) {
  using namespace am;
  AM_CHECK(issue_cycles && total_cycles && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && d_tiles >= 1 &&
               d_tiles * N <= 512 && d_tiles <= 3,
           "am_probe_mma: bad argument");
  AM_TRY(ensure_init());
  const int grid = sm_count();
  DevBuf<long long> out;
  AM_TRY(out.alloc((size_t)grid * 2));
  const size_t smem = 3 * 16384 + 32768 + 65536 + 1024;
  AM_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_probe_kernel<<<grid, kProbeThreads, smem>>>(N, iters, d_tiles, traffic, out.p);
  AM_CUDA(cudaGetLastError());
  AM_CUDA(cudaDeviceSynchronize());
  std::vector<long long> h((size_t)grid * 2);
  AM_CUDA(cudaMemcpy(h.data(), out.p, h.size() * 8, cudaMemcpyDeviceToHost));
  double a = 0, b = 0;
  for (int i = 0; i < grid; ++i) {
    a += (double)h[2 * i];
    b += (double)h[2 * i + 1];
  }
  *issue_cycles = a / grid / iters;
  *total_cycles = b / grid / iters;
  return AM_OK;
}

extern "C" AM_API int am_probe_tmem_ld(int warps, int cols, int depth, int iters, int n_mma, double* bytes_per_cycle,
                                       double* cycles_per_mma) {
  using namespace am;
  AM_CHECK(bytes_per_cycle && warps >= 1 && warps <= 16 && (cols == 16 || cols == 32) && depth >= 1 && depth <= 4 &&
               iters > 0 && iters % depth == 0,
           "am_probe_tmem_ld: bad argument");
  AM_TRY(ensure_init());
  const int grid = sm_count();
  DevBuf<long long> out;
  AM_TRY(out.alloc((size_t)grid * 2));
  AM_CUDA(cudaMemset(out.p, 0, (size_t)grid * 16));
  AM_CHECK(n_mma >= 0 && cycles_per_mma != nullptr, "am_probe_tmem_ld: bad argument");
  AM_CHECK(!(cols == 32 && depth == 4), "am_probe_tmem_ld: x32 supports depth 1 or 2");
  const int thr = warps * 32 + (n_mma > 0 ? 32 : 0);
  AM_CHECK(n_mma == 0 || warps == 16, "am_probe_tmem_ld: the MMA stream needs warps == 16");
  const size_t smem = n_mma > 0 ? 16384 + 8192 + 1024 : 0;
  if (cols == 16 && depth == 1) tmem_ld_probe_kernel<16, 1><<<grid, thr, smem>>>(iters, n_mma, out.p);
  else if (cols == 16 && depth == 2) tmem_ld_probe_kernel<16, 2><<<grid, thr, smem>>>(iters, n_mma, out.p);
  else if (cols == 16 && depth == 4) tmem_ld_probe_kernel<16, 4><<<grid, thr, smem>>>(iters, n_mma, out.p);
  else if (cols == 32 && depth == 1) tmem_ld_probe_kernel<32, 1><<<grid, thr, smem>>>(iters, n_mma, out.p);
  else if (cols == 32 && depth == 2) tmem_ld_probe_kernel<32, 2><<<grid, thr, smem>>>(iters, n_mma, out.p);
  else AM_CHECK(false, "am_probe_tmem_ld: depth must be 1, 2 or 4");
  AM_CUDA(cudaGetLastError());
  AM_CUDA(cudaDeviceSynchronize());
  std::vector<long long> h((size_t)grid * 2);
  AM_CUDA(cudaMemcpy(h.data(), out.p, h.size() * 8, cudaMemcpyDeviceToHost));
  double cyc = 0, mcyc = 0;
  for (int i = 0; i < grid; ++i) {
    cyc += (double)h[i];
    mcyc += (double)h[grid + i];
  }
  cyc /= grid;
  *cycles_per_mma = n_mma > 0 ? mcyc / grid / n_mma : 0.0;
  *bytes_per_cycle = (double)warps * iters * 32.0 * cols * 4.0 / cyc;
  return AM_OK;
}
