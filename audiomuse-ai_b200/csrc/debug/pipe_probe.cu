// Debug probe: issue rate of the fp16x2 / pack / permute instructions the channel-per-lane depthwise is made of
// (HFMA2 register form, HFMA2 with an immediate, HFMA2.SAT, HMNMX2, PRMT, F2FP pack, and an HFMA2 + PRMT mix), as
// warp-instructions per cycle per SM sub-partition with `warps` resident warps per SM (tools/pipe_probe.py).
#include <cuda_fp16.h>

#include <vector>

#include "../common.cuh"
#include "../../../include/audiomuse_b200_debug.h"

namespace am {

template <int kOp>
__global__ void __launch_bounds__(1024, 1) pipe_probe_kernel(int iters, float seed, long long* __restrict__ out, uint32_t* __restrict__ sink) {
  constexpr int kAcc = 8;
  __half2 acc[kAcc], x[kAcc];
  const __half2 w = __floats2half2_rn(seed, seed * 0.5f), k = __floats2half2_rn(0.1666f, 0.1666f);
  float f[2 * kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    acc[i] = __floats2half2_rn(0.001f * (float)(threadIdx.x + i), 0.002f * (float)i);
    x[i] = __floats2half2_rn(0.5f + 0.01f * (float)i, 0.25f);
    f[2 * i] = seed * (float)i;
    f[2 * i + 1] = seed + (float)threadIdx.x;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < kAcc; ++i) {
        if constexpr (kOp == 0) acc[i] = __hfma2(x[i], w, acc[i]);                              // HFMA2 R, R, R, R
        else if constexpr (kOp == 1) acc[i] = __hfma2(acc[i], __floats2half2_rn(0.999f, 0.999f), x[i]);   // immediate multiplier
        else if constexpr (kOp == 2) acc[i] = __hfma2_sat(acc[i], k, x[i]);                     // HFMA2.SAT
        else if constexpr (kOp == 3) acc[i] = __hmin2(__hmax2(acc[i], x[i]), w);                 // 2 x HMNMX2
        else if constexpr (kOp == 4) {                                                           // PRMT
          uint32_t a = *reinterpret_cast<uint32_t*>(&acc[i]), b = *reinterpret_cast<uint32_t*>(&x[i]);
          a = __byte_perm(a, b, 0x5432);
          acc[i] = *reinterpret_cast<__half2*>(&a);
        } else if constexpr (kOp == 5) {                                                         // F2FP pack
          acc[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
          f[2 * i] += __low2float(acc[i]);                                                        // (keeps the pack alive: adds 1 HADD2.F32 / FADD)
        } else {                                                                                 // HFMA2 + PRMT mix (1 : 1)
          acc[i] = __hfma2(x[i], w, acc[i]);
          uint32_t a = *reinterpret_cast<uint32_t*>(&x[i]), b = *reinterpret_cast<uint32_t*>(&acc[i]);
          a = __byte_perm(a, b, 0x5432);
          x[i] = *reinterpret_cast<__half2*>(&a);
        }
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) s ^= *reinterpret_cast<uint32_t*>(&acc[i]) ^ *reinterpret_cast<uint32_t*>(&x[i]) ^ __float_as_uint(f[2 * i]);
  if (s == 0x12345u) sink[0] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
}  // namespace am

extern "C" AM_API int am_probe_pipe(int op, int warps, int iters, double* cycles_per_warp_instr_per_smsp) {
  using namespace am;
  AM_CHECK(cycles_per_warp_instr_per_smsp && op >= 0 && op <= 6 && warps >= 1 && warps <= 32 && iters > 0, "am_probe_pipe: bad argument");
  AM_TRY(ensure_init());
  const int grid = sm_count();
  DevBuf<long long> out;
  DevBuf<uint32_t> sink;
  AM_TRY(out.alloc((size_t)grid));
  AM_TRY(sink.alloc(1));
  const int thr = warps * 32;
  switch (op) {
    case 0: pipe_probe_kernel<0><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
    case 1: pipe_probe_kernel<1><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
    case 2: pipe_probe_kernel<2><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
    case 3: pipe_probe_kernel<3><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
    case 4: pipe_probe_kernel<4><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
    case 5: pipe_probe_kernel<5><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
    default: pipe_probe_kernel<6><<<grid, thr>>>(iters, 0.5f, out.p, sink.p); break;
  }
  AM_CUDA(cudaGetLastError());
  AM_CUDA(cudaDeviceSynchronize());
  std::vector<long long> h((size_t)grid);
  AM_CUDA(cudaMemcpy(h.data(), out.p, h.size() * 8, cudaMemcpyDeviceToHost));
  double cyc = 0;
  for (long long v : h) cyc += (double)v;
  cyc /= grid;
  // instructions of the probed kind issued per SM sub-partition: warps / 4 warps x iters x 4 x 8 (x 2 for ops 3, 5, 6)
  const double per = (op == 3 || op == 5 || op == 6) ? 2.0 : 1.0;
  *cycles_per_warp_instr_per_smsp = cyc / ((double)warps / 4.0 * iters * 32.0 * per);
  return AM_OK;
}
