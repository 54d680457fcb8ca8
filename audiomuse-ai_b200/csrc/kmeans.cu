// K5: Lloyd k-means (sm_100a).  Replaces cuml.cluster.KMeans(...).fit_predict as called by
// tasks/clustering_gpu.py:100-123 (k-means++ init, n_init restarts, labels + cluster_centers_).
//
// Per Lloyd iteration, k <= 128 (kmeans_tc.cu): the assignment is a split-bf16 tcgen05 GEMM with a fused argmin
// epilogue (+ an exact fp32 recheck of near-ties), the partial sums a sort-by-label pass that reads every row
// once -- two passes over the data per iteration, HBM-bound.  This file keeps the driver (k-means++ seeding,
// sklearn's tolerance / empty-cluster rules, restarts) and the CUDA-core kernels that serve k > 128 and small
// problems:
//   assign   one warp per point; centres streamed through L1/L2; argmin_c (||c||^2 - 2 x.c)
//            with fp32 FMAs, lowest index wins ties; inertia accumulated in float64  (compute-bound on CUDA cores);
//   update   label-segmented column sums in shared memory ([k, W] slab per CTA, W columns),
//            flushed with one global atomicAdd per (centre, column) per CTA.
// Multi-GPU (dist.py): rows stay sharded; am_kmeans_plan_step / am_kmeans_assign_dev produce per-rank partial
// sums / counts which the host all-reduces (NCCL) before dividing.
#include "common.cuh"
#include "kmeans_tc.cuh"

#include <algorithm>
#include <cmath>
#include <memory>

namespace am {

__global__ void center_norms_kernel(const float* __restrict__ C, int k, int d, float* __restrict__ cn) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= k) return;
  float acc = 0.f;
  for (int i = lane; i < d; i += 32) acc = fmaf(C[(int64_t)j * d + i], C[(int64_t)j * d + i], acc);
  acc = warp_sum(acc);
  if (lane == 0) cn[j] = acc;
}

// one warp per point; each lane keeps up to 16 features of x in registers per 512-chunk
__global__ void __launch_bounds__(256)
assign_kernel(const float* __restrict__ X, int64_t N, int d, const float* __restrict__ C,
              const float* __restrict__ cn, int k, int32_t* __restrict__ labels,
              float* __restrict__ counts, double* __restrict__ inertia, float* __restrict__ dist) {
  __shared__ double s_inertia[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  double local = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * warps + warp; row < N; row += (int64_t)gridDim.x * warps) {
    const float* x = X + row * d;
    float best = INFINITY;
    int best_j = 0;
    float xn = 0.f;
    for (int i = lane; i < d; i += 32) xn = fmaf(x[i], x[i], xn);
    xn = warp_sum(xn);
    for (int j = 0; j < k; ++j) {
      const float* c = C + (int64_t)j * d;
      float acc = 0.f;
      for (int i = lane; i < d; i += 32) acc = fmaf(__ldg(&x[i]), __ldg(&c[i]), acc);
      acc = warp_sum(acc);
      const float v = cn[j] - 2.0f * acc;
      if (v < best) {
        best = v;
        best_j = j;
      }
    }
    if (lane == 0) {
      labels[row] = best_j;
      if (counts) atomicAdd(&counts[best_j], 1.0f);
      if (dist) dist[row] = fmaxf(best + xn, 0.0f);
      local += (double)fmaxf(best + xn, 0.0f);
    }
  }
  if (lane == 0) s_inertia[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0 && inertia) {
    double t = 0.0;
    for (int w = 0; w < warps; ++w) t += s_inertia[w];
    atomicAdd(inertia, t);
  }
}

// sums[label[i], c0:c0+W] += X[i, c0:c0+W] for a chunk of points, via a smem slab [k, W]
__global__ void __launch_bounds__(256)
accumulate_kernel(const float* __restrict__ X, int64_t N, int d, const int32_t* __restrict__ labels, int k,
                  int W, int points_per_cta, float* __restrict__ sums) {
  extern __shared__ float s_acc[];  // [k * W]
  const int c0 = blockIdx.y * W;
  const int wcols = min(W, d - c0);
  for (int i = threadIdx.x; i < k * W; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t p0 = (int64_t)blockIdx.x * points_per_cta;
  const int64_t p1 = min(N, p0 + points_per_cta);
  const int rows_per_iter = blockDim.x / W;  // W divides blockDim.x
  const int col = threadIdx.x % W, r = threadIdx.x / W;
  for (int64_t p = p0 + r; p < p1; p += rows_per_iter) {
    if (col < wcols) atomicAdd(&s_acc[labels[p] * W + col], X[p * d + c0 + col]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k * W; i += blockDim.x) {
    const int j = i / W, c = i % W;
    const float v = s_acc[i];
    if (c < wcols && v != 0.f) atomicAdd(&sums[(int64_t)j * d + c0 + c], v);
  }
}

// C_new = sums / counts (empty clusters keep their centre); shift2 += ||C_new - C||^2
__global__ void update_centers_kernel(const float* __restrict__ sums, const float* __restrict__ counts, int k,
                                      int d, float* __restrict__ C, double* __restrict__ shift2) {
  const int j = blockIdx.x;
  const float cnt = counts[j];
  double local = 0.0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float old = C[(int64_t)j * d + i];
    const float nw = cnt > 0.f ? sums[(int64_t)j * d + i] / cnt : old;
    C[(int64_t)j * d + i] = nw;
    const double df = (double)nw - (double)old;
    local += df * df;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) atomicAdd(shift2, local);
}

// sklearn's _relocate_empty_clusters_dense (sklearn/cluster/_k_means_common.pyx): each empty cluster takes the
// point that is farthest from its own centre (next farthest for the next empty cluster, ...); that point's row
// leaves the sums / counts of the cluster it was assigned to.  One CTA, the reassignments are sequential because
// two of them may hit the same donor cluster.  Labels are NOT changed here (as in sklearn: the next E-step does).
__global__ void relocate_empty_kernel(const float* __restrict__ X, int d, const int32_t* __restrict__ labels,
                                      const int64_t* __restrict__ far_rows, const int32_t* __restrict__ empty_ids,
                                      int n_empty, float* __restrict__ sums, float* __restrict__ counts) {
  for (int e = 0; e < n_empty; ++e) {
    const int64_t row = far_rows[e];
    const int donor = labels[row], target = empty_ids[e];
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      const float v = X[row * d + i];
      sums[(int64_t)donor * d + i] -= v;
      sums[(int64_t)target * d + i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      counts[target] = 1.0f;
      counts[donor] -= 1.0f;
    }
    __syncthreads();
  }
}

// per-feature variance of X, summed (for sklearn's tol scaling): out[0] += sum_j var_j
__global__ void column_moments_kernel(const float* __restrict__ X, int64_t N, int d, double* __restrict__ s1,
                                      double* __restrict__ s2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int64_t chunk = (N + gridDim.y - 1) / gridDim.y;
  const int64_t p0 = (int64_t)blockIdx.y * chunk, p1 = min(N, p0 + chunk);
  double a = 0.0, b = 0.0;
  for (int64_t p = p0; p < p1; ++p) {
    const double v = (double)X[p * d + c];
    a += v;
    b += v * v;
  }
  atomicAdd(&s1[c], a);
  atomicAdd(&s2[c], b);
}

// k-means++: mind2[i] = min(mind2[i], ||x_i - c||^2); per-block partial sums of mind2
__global__ void __launch_bounds__(256)
pp_update_kernel(const float* __restrict__ X, int64_t N, int d, const float* __restrict__ c, int first,
                 float* __restrict__ mind2, double* __restrict__ block_sums) {
  __shared__ double s_part[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  double part = 0.0;
  for (int r = warp; r < 1024; r += 8) {
    const int64_t row = base + r;
    if (row >= N) break;
    const float* x = X + row * d;
    float acc = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float df = x[i] - __ldg(&c[i]);
      acc = fmaf(df, df, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float m = first ? acc : fminf(mind2[row], acc);
      mind2[row] = m;
      part += (double)m;
    }
  }
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    block_sums[blockIdx.x] = t;
  }
}

// greedy k-means++ (sklearn's _kmeans_plusplus): for each of L candidate rows, the potential
// sum_i min(mind2[i], ||x_i - c_l||^2); per-block partial sums -> block_pot[block][L]
constexpr int kMaxTrials = 8;
__global__ void __launch_bounds__(256)
pp_trial_kernel(const float* __restrict__ X, int64_t N, int d, const float* __restrict__ cand, int L,
                const float* __restrict__ mind2, double* __restrict__ block_pot) {
  __shared__ double s_part[8][kMaxTrials];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * 1024;
  double part[kMaxTrials];
#pragma unroll
  for (int l = 0; l < kMaxTrials; ++l) part[l] = 0.0;
  for (int r = warp; r < 1024; r += 8) {
    const int64_t row = base + r;
    if (row >= N) break;
    const float* x = X + row * d;
    float acc[kMaxTrials];
#pragma unroll
    for (int l = 0; l < kMaxTrials; ++l) acc[l] = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float xv = x[i];
#pragma unroll
      for (int l = 0; l < kMaxTrials; ++l) {
        if (l < L) {
          const float df = xv - __ldg(&cand[(int64_t)l * d + i]);
          acc[l] = fmaf(df, df, acc[l]);
        }
      }
    }
    const float m = mind2[row];
#pragma unroll
    for (int l = 0; l < kMaxTrials; ++l) {
      if (l < L) {
        const float v = warp_sum(acc[l]);
        part[l] += (double)fminf(m, v);
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < kMaxTrials; ++l) s_part[warp][l] = part[l];
  }
  __syncthreads();
  if (threadIdx.x < kMaxTrials) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
    block_pot[(int64_t)blockIdx.x * kMaxTrials + threadIdx.x] = t;
  }
}

struct SplitMix {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

static int accumulate_width(int k) { return k <= 128 ? 128 : (k <= 256 ? 64 : 32); }

static int launch_accumulate(const float* X, int64_t N, int d, const int32_t* labels, int k, float* sums,
                             cudaStream_t st) {
  const int W = accumulate_width(k);
  const size_t smem = (size_t)k * W * 4;
  AM_CHECK(smem <= 200 * 1024, "kmeans: k=%d too large for the shared-memory accumulation slab", k);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AM_CUDA(cudaFuncSetAttribute(accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int ppc = 4096;
  dim3 grid((unsigned)((N + ppc - 1) / ppc), (unsigned)ceil_div(d, W));
  AM_LAUNCH(accumulate_kernel, grid, 256, smem, st, X, N, d, labels, k, W, ppc, sums);
  return AM_OK;
}

static int assign_pass(const float* X, int64_t N, int d, const float* C, float* cn, int k, int32_t* labels,
                       float* sums, float* counts, double* inertia, cudaStream_t st, float* dist = nullptr) {
  AM_LAUNCH(center_norms_kernel, ceil_div(k, 8), 256, 0, st, C, k, d, cn);
  if (counts) AM_CUDA(cudaMemsetAsync(counts, 0, (size_t)k * 4, st));
  if (inertia) AM_CUDA(cudaMemsetAsync(inertia, 0, 8, st));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((N + 7) / 8, (int64_t)sm_count() * 8));
  AM_LAUNCH(assign_kernel, grid, 256, 0, st, X, N, d, C, cn, k, labels, counts, inertia, dist);
  if (sums) {
    AM_CUDA(cudaMemsetAsync(sums, 0, (size_t)k * d * 4, st));
    AM_TRY(launch_accumulate(X, N, d, labels, k, sums, st));
  }
  return AM_OK;
}

__global__ void f64_to_f32_kernel(const double* in, float* out) { out[0] = (float)in[0]; }

}  // namespace am

using namespace am;

extern "C" int am_kmeans_assign_dev(const float* X_dev, int64_t N, int d, const float* centers_dev, int k,
                                    int32_t* labels_dev, float* sums_dev, float* counts_dev, float* inertia_dev,
                                    void* stream) {
  AM_CHECK(X_dev && centers_dev && labels_dev, "am_kmeans_assign_dev: NULL argument");
  AM_CHECK(N > 0 && d > 0 && k > 0, "am_kmeans_assign_dev: bad shape");
  AM_TRY(ensure_init());
  cudaStream_t st = (cudaStream_t)stream;
  DevBuf<float> cn;
  DevBuf<double> inert;
  AM_TRY(cn.alloc(k));
  AM_TRY(inert.alloc(1));
  if (kmtc::usable(N, d, k) && (double)N * k * d >= 2e9) {  // big one-shot call: the split pass pays for itself
    kmtc::Plan plan;
    AM_TRY(plan.create(X_dev, N, d, k, st));
    AM_TRY(plan.step(centers_dev, labels_dev, sums_dev, counts_dev, inertia_dev ? inert.p : nullptr, nullptr, st));
    if (inertia_dev) AM_LAUNCH(f64_to_f32_kernel, 1, 1, 0, st, inert.p, inertia_dev);
    AM_CUDA(cudaStreamSynchronize(st));
    return AM_OK;
  }
  AM_TRY(assign_pass(X_dev, N, d, centers_dev, cn.p, k, labels_dev, sums_dev, counts_dev,
                     inertia_dev ? inert.p : nullptr, st));
  if (inertia_dev) AM_LAUNCH(f64_to_f32_kernel, 1, 1, 0, st, inert.p, inertia_dev);
  AM_CUDA(cudaStreamSynchronize(st));  // scratch is freed on return
  return AM_OK;
}

// ---- iterative device API: the split-bf16 copy of the rows is built once and reused by every Lloyd step
struct am_kmeans_plan {
  kmtc::Plan tc;
  bool use_tc = false;
  const float* X = nullptr;
  int64_t N = 0;
  int d = 0, k = 0;
  DevBuf<float> cn;
  DevBuf<double> inert;
};

extern "C" int am_kmeans_plan_create(const float* X_dev, int64_t N, int d, int k, void* stream, am_kmeans_plan** out) {
  AM_CHECK(out != nullptr, "am_kmeans_plan_create: out is NULL");
  *out = nullptr;
  AM_CHECK(X_dev && N > 0 && d > 0 && k > 0, "am_kmeans_plan_create: bad argument");
  AM_TRY(ensure_init());
  auto p = std::make_unique<am_kmeans_plan>();
  p->X = X_dev;
  p->N = N;
  p->d = d;
  p->k = k;
  AM_TRY(p->cn.alloc(k));
  AM_TRY(p->inert.alloc(1));
  p->use_tc = kmtc::usable(N, d, k);
  if (p->use_tc) AM_TRY(p->tc.create(X_dev, N, d, k, (cudaStream_t)stream));
  *out = p.release();
  return AM_OK;
}

extern "C" void am_kmeans_plan_free(am_kmeans_plan* p) {
  if (p) cudaDeviceSynchronize();
  delete p;
}

// rows the last step handed to the exact recheck kernel (near-ties inside the tensor-core error band); synchronises
extern "C" int am_kmeans_plan_last_recheck(am_kmeans_plan* p, void* stream, int* n_rows) {
  AM_CHECK(p && n_rows, "am_kmeans_plan_last_recheck: NULL argument");
  *n_rows = 0;
  if (!p->use_tc) return AM_OK;
  return p->tc.last_recheck_count((cudaStream_t)stream, n_rows);
}

extern "C" int am_kmeans_plan_uses_tensor_cores(const am_kmeans_plan* p) { return p && p->use_tc ? 1 : 0; }

extern "C" int am_kmeans_plan_step(am_kmeans_plan* p, const float* centers_dev, int32_t* labels_dev, float* sums_dev,
                                   float* counts_dev, float* inertia_dev, float* dist_dev, void* stream) {
  AM_CHECK(p && centers_dev && labels_dev, "am_kmeans_plan_step: NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->use_tc) {
    AM_TRY(p->tc.step(centers_dev, labels_dev, sums_dev, counts_dev, inertia_dev ? p->inert.p : nullptr, dist_dev, st));
  } else {
    AM_TRY(assign_pass(p->X, p->N, p->d, centers_dev, p->cn.p, p->k, labels_dev, sums_dev, counts_dev,
                       inertia_dev ? p->inert.p : nullptr, st, dist_dev));
  }
  if (inertia_dev) AM_LAUNCH(f64_to_f32_kernel, 1, 1, 0, st, p->inert.p, inertia_dev);
  return AM_OK;  // stream-ordered: no synchronisation
}

extern "C" int am_kmeans_fit(const float* X, int64_t N, int d, int k, int n_init, int max_iter, float tol,
                             uint64_t seed, const float* init_centers, float* centers, int32_t* labels,
                             float* inertia, int* n_iter) {
  AM_CHECK(X && centers && labels, "am_kmeans_fit: NULL buffer");
  AM_CHECK(N > 0 && d > 0 && k > 0 && k <= N, "am_kmeans_fit: need 0 < k <= N (N=%lld k=%d)", (long long)N, k);
  AM_CHECK(max_iter > 0 && n_init > 0, "am_kmeans_fit: max_iter and n_init must be positive");
  AM_TRY(ensure_init());
  Stream st;
  AM_TRY(st.create());
  DevBuf<float> dX, dC, dBestC, cn, sums, counts, mind2;
  DevBuf<int32_t> dL, dBestL;
  DevBuf<double> scal, bsum, mom;
  AM_TRY(dX.alloc((size_t)N * d));
  AM_TRY(dC.alloc((size_t)k * d));
  AM_TRY(dBestC.alloc((size_t)k * d));
  AM_TRY(cn.alloc(k));
  AM_TRY(sums.alloc((size_t)k * d));
  AM_TRY(counts.alloc(k));
  AM_TRY(dL.alloc(N));
  AM_TRY(dBestL.alloc(N));
  AM_TRY(scal.alloc(2));  // [0] inertia, [1] shift2
  AM_TRY(mom.alloc((size_t)2 * d));
  AM_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)N * d * 4, cudaMemcpyHostToDevice, st.s));
  std::unique_ptr<kmtc::Plan> plan;
  if (kmtc::usable(N, d, k) && (double)N * k * d >= 5e7) {
    plan = std::make_unique<kmtc::Plan>();
    AM_TRY(plan->create(dX.p, N, d, k, st.s));
  }

  // sklearn: tol_ = mean(var(X, axis=0)) * tol
  AM_CUDA(cudaMemsetAsync(mom.p, 0, (size_t)2 * d * 8, st.s));
  {
    dim3 grid(ceil_div(d, 128), (unsigned)std::max<int64_t>(1, std::min<int64_t>(256, N / 1024)));
    AM_LAUNCH(column_moments_kernel, grid, 128, 0, st.s, dX.p, N, d, mom.p, mom.p + d);
  }
  std::vector<double> hm((size_t)2 * d);
  AM_CUDA(cudaMemcpyAsync(hm.data(), mom.p, hm.size() * 8, cudaMemcpyDeviceToHost, st.s));
  AM_CUDA(cudaStreamSynchronize(st.s));
  double var_mean = 0.0;
  for (int c = 0; c < d; ++c) {
    const double mu = hm[c] / N;
    var_mean += hm[d + c] / N - mu * mu;
  }
  var_mean /= d;
  const double tol_abs = var_mean * tol;

  SplitMix rng{seed ^ 0x5851f42d4c957f2dull};
  double best_inertia = INFINITY;
  int best_iters = 0;
  const int restarts = init_centers ? 1 : n_init;
  const int64_t nblk = (N + 1023) / 1024;
  DevBuf<float> cand;
  DevBuf<double> bpot;
  if (!init_centers) {
    AM_TRY(mind2.alloc(N));
    AM_TRY(bsum.alloc(nblk));
    AM_TRY(cand.alloc((size_t)kMaxTrials * d));
    AM_TRY(bpot.alloc((size_t)nblk * kMaxTrials));
  }
  std::vector<double> hbs(nblk), hpot((size_t)nblk * kMaxTrials);
  std::vector<float> hcounts((size_t)k), hdist;
  DevBuf<float> pdist;
  DevBuf<int64_t> far_dev;
  DevBuf<int32_t> empty_dev;
  AM_TRY(pdist.alloc((size_t)N));
  std::vector<float> hblock(1024);

  for (int run = 0; run < restarts; ++run) {
    if (init_centers) {
      AM_CUDA(cudaMemcpyAsync(dC.p, init_centers, (size_t)k * d * 4, cudaMemcpyHostToDevice, st.s));
    } else {
      // greedy k-means++ (D^2 sampling with 2 + ln k local trials, as sklearn / cuML do)
      const int L = std::min(kMaxTrials, 2 + (int)std::log((double)k));
      int64_t pick = std::min<int64_t>((int64_t)(rng.uniform() * N), N - 1);
      AM_CUDA(cudaMemcpyAsync(dC.p, dX.p + (size_t)pick * d, (size_t)d * 4, cudaMemcpyDeviceToDevice, st.s));
      AM_LAUNCH(pp_update_kernel, (unsigned)nblk, 256, 0, st.s, dX.p, N, d, dC.p, 1, mind2.p, bsum.p);
      for (int j = 1; j < k; ++j) {
        AM_CUDA(cudaMemcpyAsync(hbs.data(), bsum.p, nblk * 8, cudaMemcpyDeviceToHost, st.s));
        AM_CUDA(cudaStreamSynchronize(st.s));
        double total = 0.0;
        for (double v : hbs) total += v;
        int64_t cand_row[kMaxTrials];
        for (int l = 0; l < L; ++l) {
          double r = rng.uniform() * total;
          int64_t b = 0;
          for (; b < nblk - 1; ++b) {
            if (r < hbs[b]) break;
            r -= hbs[b];
          }
          const int64_t b0 = b * 1024, cnt = std::min<int64_t>(1024, N - b0);
          AM_CUDA(cudaMemcpyAsync(hblock.data(), mind2.p + b0, cnt * 4, cudaMemcpyDeviceToHost, st.s));
          AM_CUDA(cudaStreamSynchronize(st.s));
          int64_t o = 0;
          for (; o < cnt - 1; ++o) {
            if (r < hblock[o]) break;
            r -= hblock[o];
          }
          cand_row[l] = b0 + o;
          AM_CUDA(cudaMemcpyAsync(cand.p + (size_t)l * d, dX.p + (size_t)cand_row[l] * d, (size_t)d * 4,
                                  cudaMemcpyDeviceToDevice, st.s));
        }
        AM_LAUNCH(pp_trial_kernel, (unsigned)nblk, 256, 0, st.s, dX.p, N, d, cand.p, L, mind2.p, bpot.p);
        AM_CUDA(cudaMemcpyAsync(hpot.data(), bpot.p, (size_t)nblk * kMaxTrials * 8, cudaMemcpyDeviceToHost, st.s));
        AM_CUDA(cudaStreamSynchronize(st.s));
        int best = 0;
        double best_pot = INFINITY;
        for (int l = 0; l < L; ++l) {
          double pot = 0.0;
          for (int64_t b = 0; b < nblk; ++b) pot += hpot[(size_t)b * kMaxTrials + l];
          if (pot < best_pot) {
            best_pot = pot;
            best = l;
          }
        }
        AM_CUDA(cudaMemcpyAsync(dC.p + (size_t)j * d, cand.p + (size_t)best * d, (size_t)d * 4,
                                cudaMemcpyDeviceToDevice, st.s));
        if (j < k - 1)
          AM_LAUNCH(pp_update_kernel, (unsigned)nblk, 256, 0, st.s, dX.p, N, d, dC.p + (size_t)j * d, 0, mind2.p,
                    bsum.p);
      }
    }
    int it = 0;
    for (it = 1; it <= max_iter; ++it) {
      if (plan) AM_TRY(plan->step(dC.p, dL.p, sums.p, counts.p, nullptr, pdist.p, st.s));
      else AM_TRY(assign_pass(dX.p, N, d, dC.p, cn.p, k, dL.p, sums.p, counts.p, nullptr, st.s, pdist.p));
      // empty clusters are relocated the way sklearn's Lloyd does it (rare: costs one [k] read-back per
      // iteration, and the [N] distances only when a cluster actually emptied)
      AM_CUDA(cudaMemcpyAsync(hcounts.data(), counts.p, (size_t)k * 4, cudaMemcpyDeviceToHost, st.s));
      AM_CUDA(cudaStreamSynchronize(st.s));
      std::vector<int32_t> empty;
      for (int j = 0; j < k; ++j)
        if (hcounts[j] == 0.f) empty.push_back(j);
      if (!empty.empty() && (int64_t)empty.size() < N) {
        hdist.resize((size_t)N);
        AM_CUDA(cudaMemcpyAsync(hdist.data(), pdist.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st.s));
        AM_CUDA(cudaStreamSynchronize(st.s));
        std::vector<int64_t> order((size_t)N);
        for (int64_t i = 0; i < N; ++i) order[(size_t)i] = i;
        std::partial_sort(order.begin(), order.begin() + (int64_t)empty.size(), order.end(), [&](int64_t a, int64_t b) {
          return hdist[(size_t)a] > hdist[(size_t)b] || (hdist[(size_t)a] == hdist[(size_t)b] && a < b);
        });
        AM_TRY(far_dev.alloc(empty.size()));
        AM_TRY(empty_dev.alloc(empty.size()));
        AM_CUDA(cudaMemcpyAsync(far_dev.p, order.data(), empty.size() * 8, cudaMemcpyHostToDevice, st.s));
        AM_CUDA(cudaMemcpyAsync(empty_dev.p, empty.data(), empty.size() * 4, cudaMemcpyHostToDevice, st.s));
        AM_LAUNCH(relocate_empty_kernel, 1, 256, 0, st.s, dX.p, d, dL.p, far_dev.p, empty_dev.p, (int)empty.size(), sums.p,
                  counts.p);
        AM_CUDA(cudaStreamSynchronize(st.s));  // far_dev / empty_dev are reused next time
      }
      AM_CUDA(cudaMemsetAsync(scal.p + 1, 0, 8, st.s));
      AM_LAUNCH(update_centers_kernel, k, 128, 0, st.s, sums.p, counts.p, k, d, dC.p, scal.p + 1);
      double shift2 = 0.0;
      AM_CUDA(cudaMemcpyAsync(&shift2, scal.p + 1, 8, cudaMemcpyDeviceToHost, st.s));
      AM_CUDA(cudaStreamSynchronize(st.s));
      if (shift2 <= tol_abs) break;
    }
    it = std::min(it, max_iter);
    // final E-step: labels and inertia consistent with the returned centres
    if (plan) AM_TRY(plan->step(dC.p, dL.p, nullptr, nullptr, scal.p, nullptr, st.s));
    else AM_TRY(assign_pass(dX.p, N, d, dC.p, cn.p, k, dL.p, nullptr, nullptr, scal.p, st.s));
    double inert = 0.0;
    AM_CUDA(cudaMemcpyAsync(&inert, scal.p, 8, cudaMemcpyDeviceToHost, st.s));
    AM_CUDA(cudaStreamSynchronize(st.s));
    if (inert < best_inertia) {
      best_inertia = inert;
      best_iters = it;
      AM_CUDA(cudaMemcpyAsync(dBestC.p, dC.p, (size_t)k * d * 4, cudaMemcpyDeviceToDevice, st.s));
      AM_CUDA(cudaMemcpyAsync(dBestL.p, dL.p, (size_t)N * 4, cudaMemcpyDeviceToDevice, st.s));
    }
  }
  AM_CUDA(cudaMemcpyAsync(centers, dBestC.p, (size_t)k * d * 4, cudaMemcpyDeviceToHost, st.s));
  AM_CUDA(cudaMemcpyAsync(labels, dBestL.p, (size_t)N * 4, cudaMemcpyDeviceToHost, st.s));
  AM_CUDA(cudaStreamSynchronize(st.s));
  if (inertia) *inertia = (float)best_inertia;
  if (n_iter) *n_iter = best_iters;
  return AM_OK;
}
