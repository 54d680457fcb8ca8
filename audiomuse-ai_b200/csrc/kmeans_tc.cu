// K5 on the tensor cores: one Lloyd E-step + M-step partial sums for k <= 128 clusters (sm_100a).
//
// Replaces cuml.cluster.KMeans's assignment GEMM (tasks/clustering_gpu.py:100-123).
//
//   split (once per data set)   X f32 [N, d] -> Xs bf16 [N, 2*dp] = [hi | lo]  (x = hi + lo + O(2^-18 x)),
//                               xn[N] = ||x||^2;  dp = d rounded up to 64 (zero padded).
//   per iteration
//     centre prep               C f32 [k, d] -> Cs bf16 [2*kp, dp] = rows [0, kp): hi, [kp, 2kp): lo;  cn[j] = ||c_j||^2
//                               (+inf for the padded rows j >= k), cmax = max ||c_j||.
//     assign_tc_kernel          persistent warp-specialised CTAs, a tile = 128 points:
//                                 warp 0   TMA: per 64-wide K chunk the A_hi, A_lo [128 x 64] and B [2kp x 64] tiles
//                                          (SWIZZLE_128B) into a 3-stage mbarrier ring;
//                                 warp 1   one elected thread issues tcgen05.mma.kind::f16 (M128, N = kp, K16):
//                                          D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo   -- an fp32-class dot product
//                                          (dropped term lo.lo <= 2^-18 |x||c|), accumulators double-buffered in TMEM;
//                                 warps 2-5 FUSED ARGMIN EPILOGUE: thread = point; tcgen05.ld 32 columns at a time,
//                                          v_j = cn_j - 2 D_j, running best / second best; writes label and distance,
//                                          nothing else -- the [N, k] score matrix never exists in memory.
//                               A point whose runner-up is within the proven error band of the best
//                               (2^-11 ||x|| max||c||) is appended to a recheck list.
//     recheck_kernel            exact fp32 argmin (the CUDA-core arithmetic of kmeans.cu's assign_kernel) for the
//                               listed points only: labels equal the exact-arithmetic labels for EVERY point.
//     accumulate_sorted_kernel  M-step partial sums: a CTA counting-sorts the labels of its 2048-point slab in
//                               shared memory, then each warp walks one label segment: rows are read once, as whole
//                               coalesced rows, summed in registers, the exact ||x - c||^2 taken on the way (inertia),
//                               and flushed with one red.global.add per (cluster, column) per slab -- no
//                               shared-memory atomics on the data path.
//
// HBM traffic per iteration: Xs once (assign) + X once (accumulate) = 2 * N * d * 4 bytes.
#include "kmeans_tc.cuh"

#include "gemm_tcgen05.cuh"
#include "ptx_sm100.cuh"

#include <algorithm>
#include <cmath>

namespace am {
namespace kmtc {

using namespace ptx;

constexpr int kTileM = 128;
constexpr int kChunkK = 64;
constexpr int kThreads = 64 + 4 * 32;  // TMA warp, MMA warp, 4 epilogue warps
constexpr int kATile = kTileM * kChunkK * 2;  // 16 KiB
constexpr int kSlabMax = 4096;                // most points one accumulate CTA sorts at a time

// ---------------------------------------------------------------- split passes
// one warp per row: Xs[row] = [hi | lo], xn[row] = ||x||^2 (fp32)
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ X, int64_t N, int d, int dp, __nv_bfloat16* __restrict__ Xs,
                  float* __restrict__ xn) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < N; row += warps) {
    const float* x = X + row * d;
    __nv_bfloat16* o = Xs + row * 2 * dp;
    float acc = 0.f;
    for (int i = lane; i < dp; i += 32) {
      const float v = i < d ? x[i] : 0.f;
      acc = fmaf(v, v, acc);
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      o[i] = hi;
      o[dp + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
    acc = warp_sum(acc);
    if (lane == 0 && xn) xn[row] = acc;
  }
}

// one warp per centre row (incl. the padded rows): Cs, cn, and the max norm (as int bits of a non-negative float)
__global__ void __launch_bounds__(256)
split_centers_kernel(const float* __restrict__ C, int k, int d, int kp, int dp, __nv_bfloat16* __restrict__ Cs,
                     float* __restrict__ cn, int* __restrict__ cmax2_bits) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= kp) return;
  float acc = 0.f;
  for (int i = lane; i < dp; i += 32) {
    const float v = (j < k && i < d) ? C[(int64_t)j * d + i] : 0.f;
    acc = fmaf(v, v, acc);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    Cs[(int64_t)j * dp + i] = hi;
    Cs[(int64_t)(kp + j) * dp + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    cn[j] = j < k ? acc : INFINITY;
    if (j < k) atomicMax(cmax2_bits, __float_as_int(acc));
  }
}

// ---------------------------------------------------------------- assignment GEMM with the fused argmin epilogue
struct AssignArgs {
  int64_t N;
  int kp, dp, k;
  int tiles;
  const float* cn;        // [kp]
  const float* xn;        // [N]
  const int* cmax2_bits;  // max ||c||^2
  int32_t* labels;        // [N]
  float* dist;            // [N] or NULL: max(best + xn, 0)
  int* n_recheck;         // counter
  int32_t* recheck;       // [N] row list
  float band_scale;       // 2^-11
};

__global__ void __launch_bounds__(kThreads, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c, const AssignArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int kStages = 3;
  const int b_bytes = 2 * args.kp * kChunkK * 2;  // hi rows then lo rows
  const int stage_bytes = 2 * kATile + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_cn = reinterpret_cast<float*>(tmem_ptr + 4);  // [kp]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = args.dp / kChunkK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_c);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 256);
  for (int i = threadIdx.x; i < args.kp; i += kThreads) s_cn[i] = args.cn[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < args.tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          tma_load_2d(sa, &map_x, &full_bar[stage], kb * kChunkK, tile * kTileM);                    // hi
          tma_load_2d(sa + kATile, &map_x, &full_bar[stage], args.dp + kb * kChunkK, tile * kTileM);  // lo
          tma_load_2d(sa + 2 * kATile, &map_c, &full_bar[stage], kb * kChunkK, 0);                    // centres hi | lo
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(kTileM, args.kp);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < args.tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 128);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t d_ahi = make_smem_desc(sa), d_alo = make_smem_desc(sa + kATile);
          const uint64_t d_bhi = make_smem_desc(sa + 2 * kATile);
          const uint64_t d_blo = make_smem_desc(sa + 2 * kATile + (uint32_t)args.kp * 128u);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks) {
            const uint64_t o = (uint64_t)(ks * 2);  // 16 bf16 = 32 bytes along K inside the swizzle atom
            umma_f16(tmem_d, d_ahi + o, d_bhi + o, idesc, (kb | ks) ? 1u : 0u);
            umma_f16(tmem_d, d_alo + o, d_bhi + o, idesc, 1u);
            umma_f16(tmem_d, d_ahi + o, d_blo + o, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== fused argmin epilogue: thread = point =====================
    const int lane_grp = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const float cmax = sqrtf(__int_as_float(*args.cmax2_bits));
    for (int tile = blockIdx.x; tile < args.tiles; tile += gridDim.x) {
      const int64_t row = (int64_t)tile * kTileM + lane_grp * 32 + lane;
      const bool row_ok = row < args.N;
      const float xn = row_ok ? __ldg(&args.xn[row]) : 0.f;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * 128);
      float best = INFINITY, second = INFINITY;
      int best_j = 0;
      for (int c = 0; c < args.kp; c += 32) {
        uint32_t v[32];
        if (args.kp - c >= 32) {
          tmem_ld_x32(taddr0 + (uint32_t)c, v);
        } else {
          uint32_t lo[16];
          tmem_ld_x16(taddr0 + (uint32_t)c, lo);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = lo[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0u;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c + j < args.kp) {
            const float s = fmaf(-2.0f, __uint_as_float(v[j]), s_cn[c + j]);  // +inf for padded centres
            if (s < best) {
              second = best;
              best = s;
              best_j = c + j;
            } else if (s < second) {
              second = s;
            }
          }
        }
      }
      // release the accumulator before the global writes
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (row_ok) {
        args.labels[row] = best_j;
        if (args.dist) args.dist[row] = fmaxf(best + xn, 0.f);
        // |v~ - v| <= 2^-13 ||x|| max||c|| per centre (split-bf16 residuals + fp32 accumulation over dp terms);
        // runner-up within 4x that of the winner: let the exact kernel decide
        const float band = args.band_scale * sqrtf(xn) * cmax;
        if (second - best <= band) {
          const int slot = atomicAdd(args.n_recheck, 1);
          args.recheck[slot] = (int32_t)row;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// exact fp32 argmin for the listed rows, the arithmetic of kmeans.cu's assign_kernel (v_j = cn_j - 2 x.c_j with
// lane-strided fp32 FMAs + a shuffle tree; strict "<" in index order, so the lowest index wins ties).  One CTA per
// row, the centres dealt round-robin to its 8 warps (a single warp walking all k centres took ~0.3 ms per row, and
// a kernel is as slow as its slowest row); each warp keeps four dot products in flight.
__global__ void __launch_bounds__(256)
recheck_kernel(const float* __restrict__ X, int d, const float* __restrict__ C, const float* __restrict__ cn, int k,
               const int* __restrict__ n_recheck, const int32_t* __restrict__ recheck, int32_t* __restrict__ labels,
               float* __restrict__ dist) {
  __shared__ float s_best[8];
  __shared__ int s_idx[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = *n_recheck;
  for (int q = blockIdx.x; q < n; q += gridDim.x) {
    const int64_t row = recheck[q];
    const float* x = X + row * d;
    float xn = 0.f;
    for (int i = lane; i < d; i += 32) xn = fmaf(x[i], x[i], xn);
    xn = warp_sum(xn);
    float best = INFINITY;
    int best_j = 0x7fffffff;
    for (int j0 = warp; j0 < k; j0 += 32) {  // centres j0, j0 + 8, j0 + 16, j0 + 24 together
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = lane; i < d; i += 32) {
        const float xv = __ldg(&x[i]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 8 * u;
          if (j < k) acc[u] = fmaf(xv, __ldg(&C[(int64_t)j * d + i]), acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 8 * u;
        const float a = warp_sum(acc[u]);
        if (j < k) {
          const float v = cn[j] - 2.0f * a;
          if (v < best) {  // j increases inside a warp: ties keep the lower index
            best = v;
            best_j = j;
          }
        }
      }
    }
    if (lane == 0) {
      s_best[warp] = best;
      s_idx[warp] = best_j;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = s_best[0];
      int bj = s_idx[0];
      for (int w = 1; w < 8; ++w)
        if (s_best[w] < b || (s_best[w] == b && s_idx[w] < bj)) {
          b = s_best[w];
          bj = s_idx[w];
        }
      labels[row] = bj;
      if (dist) dist[row] = fmaxf(b + xn, 0.0f);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- M-step partial sums
// sums[j, :] += sum of the rows labelled j; counts[j] += their number; inertia += sum ||x - c_label||^2 (exact, fp32
// per row, float64 across rows).  kMaxChunks 128-column float4 chunks per lane cover d <= 512 in registers.
template <int kVec>
__global__ void __launch_bounds__(256)
accumulate_sorted_kernel(const float* __restrict__ X, int64_t N, int d, const int32_t* __restrict__ labels, int k,
                         const float* __restrict__ C, int slab, float* __restrict__ sums, float* __restrict__ counts,
                         double* __restrict__ inertia) {
  extern __shared__ int s_mem[];
  int* s_hist = s_mem;            // [k + 1] start offsets after the scan
  int* s_cursor = s_hist + k + 1;  // [k]
  int* s_order = s_cursor + k;    // [slab] rows of the slab grouped by label
  __shared__ double s_inertia[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double local = 0.0;
  // the grid is sized to ONE resident wave (a second, partial wave doubled the kernel's time): every CTA takes the same
  // number of slabs
  for (int64_t p0 = (int64_t)blockIdx.x * slab; p0 < N; p0 += (int64_t)gridDim.x * slab) {
    const int np = (int)min((int64_t)slab, N - p0);
    __syncthreads();
    for (int i = threadIdx.x; i <= k; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += blockDim.x) atomicAdd(&s_hist[labels[p0 + i] + 1], 1);
    __syncthreads();
    if (warp == 0) {  // inclusive scan of the k + 1 bins
      int carry = 0;
      for (int base = 0; base <= k; base += 32) {
        const int idx = base + lane;
        int v = idx <= k ? s_hist[idx] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, v, o);
          if (lane >= o) v += t;
        }
        v += carry;
        if (idx <= k) s_hist[idx] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) s_cursor[i] = s_hist[i];
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
      const int slot = atomicAdd(&s_cursor[labels[p0 + i]], 1);
      s_order[slot] = i;
    }
    __syncthreads();
    constexpr int kMaxChunks = 4;  // 4 x 32 lanes x kVec columns per pass (512 columns with float4 loads)
    for (int col0 = 0; col0 < d; col0 += kMaxChunks * 32 * kVec) {
      for (int j = warp; j < k; j += 8) {
        const int s0 = s_hist[j], s1 = s_hist[j + 1];
        if (s0 == s1) continue;
        float acc[kMaxChunks][kVec], cj[kMaxChunks][kVec];
#pragma unroll
        for (int q = 0; q < kMaxChunks; ++q)
#pragma unroll
          for (int e = 0; e < kVec; ++e) {
            const int col = col0 + (q * 32 + lane) * kVec + e;
            acc[q][e] = 0.f;
            cj[q][e] = col < d ? __ldg(&C[(int64_t)j * d + col]) : 0.f;
          }
        float dsum = 0.f;
        auto consume = [&](const float (&v)[kMaxChunks][kVec]) {
          float dloc = 0.f;
#pragma unroll
          for (int q = 0; q < kMaxChunks; ++q)
#pragma unroll
            for (int e = 0; e < kVec; ++e) {
              acc[q][e] += v[q][e];
              const float a = v[q][e] - cj[q][e];
              dloc = fmaf(a, a, dloc);
            }
          dsum += dloc;
        };
        auto load = [&](int s, float (&v)[kMaxChunks][kVec]) {
          const float* x = X + (p0 + s_order[s]) * d;
#pragma unroll
          for (int q = 0; q < kMaxChunks; ++q) {
            const int col = col0 + (q * 32 + lane) * kVec;
            if constexpr (kVec == 4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (col < d) t = __ldg(reinterpret_cast<const float4*>(x + col));
              v[q][0] = t.x; v[q][1] = t.y; v[q][2] = t.z; v[q][3] = t.w;
            } else {
              v[q][0] = col < d ? __ldg(x + col) : 0.f;
            }
          }
        };
        // padded columns load 0 and their centre entry is 0: they add nothing to sums or distances
        int s = s0;
        for (; s + 1 < s1; s += 2) {  // two rows in flight per warp
          float va[kMaxChunks][kVec], vb[kMaxChunks][kVec];
          load(s, va);
          load(s + 1, vb);
          consume(va);
          consume(vb);
        }
        if (s < s1) {
          float va[kMaxChunks][kVec];
          load(s, va);
          consume(va);
        }
#pragma unroll
        for (int q = 0; q < kMaxChunks; ++q)
#pragma unroll
          for (int e = 0; e < kVec; ++e) {
            const int col = col0 + (q * 32 + lane) * kVec + e;
            if (col < d) atomicAdd(&sums[(int64_t)j * d + col], acc[q][e]);
          }
        dsum = warp_sum(dsum);
        if (lane == 0) {
          local += (double)dsum;
          if (col0 == 0 && counts) atomicAdd(&counts[j], (float)(s1 - s0));
        }
      }
    }
  }
  if (lane == 0) s_inertia[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0 && inertia) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_inertia[w];
    atomicAdd(inertia, t);
  }
}

// ---------------------------------------------------------------- host
bool usable(int64_t N, int d, int k) {
  return gemm::available() && k >= 1 && k <= 128 && d >= 1 && d <= 4096 && N >= 1 && N < ((int64_t)1 << 31) &&
         std::getenv("AM_KMEANS_SIMT") == nullptr;
}

int Plan::create(const float* X_dev, int64_t N_, int d_, int k_, cudaStream_t st) {
  N = N_;
  d = d_;
  k = k_;
  X = X_dev;
  dp = (int)round_up((size_t)d, 64);
  kp = (int)round_up((size_t)k, 16);
  AM_TRY(Xs.alloc((size_t)N * 2 * dp));
  AM_TRY(xn.alloc((size_t)N));
  AM_TRY(Cs.alloc((size_t)2 * kp * dp));
  AM_TRY(cn.alloc((size_t)kp));
  AM_TRY(scal.alloc(2));  // [0] max ||c||^2 bits, [1] recheck counter
  AM_TRY(recheck.alloc((size_t)N));
  AM_TRY(inertia64.alloc(1));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((N + 7) / 8, (int64_t)sm_count() * 16));
  AM_LAUNCH(split_rows_kernel, grid, 256, 0, st, X, N, d, dp, Xs.p, xn.p);
  {
    const uint64_t dims[2] = {(uint64_t)(2 * dp), (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)(2 * dp) * 2};
    const uint32_t box[2] = {(uint32_t)kChunkK, (uint32_t)kTileM};
    AM_TRY(gemm::encode_map_bf16(map_x, Xs.p, 2, dims, strides, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)dp, (uint64_t)(2 * kp)};
    const uint64_t strides[1] = {(uint64_t)dp * 2};
    const uint32_t box[2] = {(uint32_t)kChunkK, (uint32_t)(2 * kp)};
    AM_TRY(gemm::encode_map_bf16(map_c, Cs.p, 2, dims, strides, box));
  }
  return AM_OK;
}

int Plan::launch_accumulate(float* sums, float* counts, double* inertia_dev, const float* C_dev, const int32_t* labels,
                            cudaStream_t st) {
  // one resident wave: 3 CTAs per SM (register limited), each takes ceil(N / grid) points in slabs of <= kSlabMax
  const int64_t ctas = (int64_t)sm_count() * 3;
  int slab = (int)std::min<int64_t>(kSlabMax, std::max<int64_t>(256, (N + ctas - 1) / ctas));
  const int64_t n_slabs = (N + slab - 1) / slab;
  const unsigned grid = (unsigned)std::min<int64_t>(n_slabs, ctas);
  const size_t smem = (size_t)(2 * k + 1 + slab) * sizeof(int);
  if (d % 4 == 0) {
    AM_LAUNCH(accumulate_sorted_kernel<4>, grid, 256, smem, st, X, N, d, labels, k, C_dev, slab, sums, counts, inertia_dev);
  } else {
    AM_LAUNCH(accumulate_sorted_kernel<1>, grid, 256, smem, st, X, N, d, labels, k, C_dev, slab, sums, counts, inertia_dev);
  }
  return AM_OK;
}

int Plan::step(const float* C_dev, int32_t* labels, float* sums, float* counts, double* inertia_dev, float* dist,
               cudaStream_t st) {
  AM_CUDA(cudaMemsetAsync(scal.p, 0, 2 * sizeof(int), st));
  AM_LAUNCH(split_centers_kernel, ceil_div(kp, 8), 256, 0, st, C_dev, k, d, kp, dp, Cs.p, cn.p, scal.p);
  AssignArgs a{};
  a.N = N;
  a.kp = kp;
  a.dp = dp;
  a.k = k;
  a.tiles = (int)((N + kTileM - 1) / kTileM);
  a.cn = cn.p;
  a.xn = xn.p;
  a.cmax2_bits = scal.p;
  a.labels = labels;
  a.dist = dist;
  a.n_recheck = scal.p + 1;
  a.recheck = recheck.p;
  a.band_scale = 1.0f / 2048.0f;
  const size_t smem = 1024 + 3 * (size_t)(2 * kATile + 2 * kp * kChunkK * 2) + 256 + (size_t)kp * 4;
  static size_t attr = 0;
  if (smem > attr) {
    AM_CUDA(cudaFuncSetAttribute(assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int grid = std::min(a.tiles, sm_count());
  AM_LAUNCH(assign_tc_kernel, grid, kThreads, smem, st, *reinterpret_cast<const CUtensorMap*>(map_x),
            *reinterpret_cast<const CUtensorMap*>(map_c), a);
  AM_LAUNCH(recheck_kernel, sm_count() * 8, 256, 0, st, X, d, C_dev, cn.p, k, scal.p + 1, recheck.p, labels, dist);
  if (sums) {
    AM_CUDA(cudaMemsetAsync(sums, 0, (size_t)k * d * 4, st));
    if (counts) AM_CUDA(cudaMemsetAsync(counts, 0, (size_t)k * 4, st));
    if (inertia_dev) AM_CUDA(cudaMemsetAsync(inertia_dev, 0, 8, st));
    AM_TRY(launch_accumulate(sums, counts, inertia_dev, C_dev, labels, st));
  } else if (inertia_dev) {  // final E-step: inertia only (sums go to scratch-free path: counts ignored)
    AM_CUDA(cudaMemsetAsync(inertia_dev, 0, 8, st));
    AM_TRY(scratch_sums.ensure((size_t)k * d));
    AM_CUDA(cudaMemsetAsync(scratch_sums.p, 0, (size_t)k * d * 4, st));
    AM_TRY(launch_accumulate(scratch_sums.p, nullptr, inertia_dev, C_dev, labels, st));
  }
  return AM_OK;
}

int Plan::last_recheck_count(cudaStream_t st, int* out) {
  AM_CUDA(cudaMemcpyAsync(out, scal.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  AM_CUDA(cudaStreamSynchronize(st));
  return AM_OK;
}

}  // namespace kmtc
}  // namespace am
