// tcgen05 / TMEM / TMA GEMM for sm_100a (see gemm_tcgen05.cuh for the contract).
//
// Persistent, warp-specialised CTA of 320 threads, one CTA per SM:
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the A tile [128 x 64] and the
//               B tile [BLOCK_N x 64] (bf16, 128-byte swizzle) into a 4-stage smem ring;
//   warp 1      TMEM allocator + MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::f16
//               (M = 128, N = BLOCK_N <= 256, K = 16 per instruction), accumulators in TMEM,
//               double-buffered (2 x BLOCK_N of the 512 columns) so the epilogue of tile i
//               overlaps the MMAs of tile i+1; tcgen05.commit releases smem stages / signals
//               the epilogue through mbarriers;
//   warps 2..9  epilogue: the tile's bias slice is staged once in smem; tcgen05.ld (32 lanes x 32
//               columns per instruction, two warps per lane group taking alternate chunks) ->
//               registers -> alpha / bias / ReLU6 / residual -> bf16 or fp32 -> 16-byte global stores.
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
#include "gemm_tcgen05.cuh"
#include "ptx_sm100.cuh"

#include <mutex>

namespace am {
namespace gemm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int kStages = 4;             // at most; 3 when a 256-wide B tile and the store staging would not fit
constexpr int kEpiWarps = 8;           // 2 per TMEM lane group (even / odd 32-column chunks)
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxBlockN = 256;
constexpr int kTmemCols = 512;
constexpr int kStoreStageBytes = 32 * 128;  // per epilogue warp: 32 rows x 128 B (fp32) or 64 B (bf16) for the TMA store

using namespace ptx;

struct KernelArgs {
  int64_t M, N;
  int K;
  int block_n;
  int tiles_m, tiles_n;
  int m_fastest;
  void* D;
  int64_t ldd;
  int d_is_f32;
  float alpha;
  const float* bias;
  const float* col_sub;
  int act;
  const __nv_bfloat16* residual;
  int64_t ld_res;
  float* chunk_max;  // k-NN: per-32-column maxima beside D (see Epilogue::chunk_max)
  int64_t ld_cm;
  int stages;     // smem ring depth (3 or 4)
  int tma_store;  // epilogue writes D through shared memory + TMA tile stores (coalesced, asynchronous)
};

__device__ __forceinline__ void tile_coords(const KernelArgs& a, int tile, int& m_blk, int& n_blk) {
  if (a.m_fastest) {
    m_blk = tile % a.tiles_m;
    n_blk = tile / a.tiles_m;
  } else {
    n_blk = tile % a.tiles_n;
    m_blk = tile / a.tiles_n;
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_d, const KernelArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_tile_bytes = args.block_n * kBlockK * 2;
  const int stage_bytes = kATileBytes + b_tile_bytes;
  const int n_stages = args.stages;
  uint8_t* store_stage = smem + n_stages * stage_bytes;  // [kEpiWarps][kStoreStageBytes], 1024-byte aligned
  const int store_stage_bytes = args.d_is_f32 ? kStoreStageBytes : kStoreStageBytes / 2;  // per epilogue warp
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + (args.tma_store ? kEpiWarps * store_stage_bytes : 0));
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 4);  // [2][kMaxBlockN]: bias - col_sub per column

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = args.tiles_m * args.tiles_n;
  const int num_kb = (args.K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    if (args.tma_store) prefetch_tensormap(&map_d);
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(args, tile, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + kATileBytes;
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
          tma_load_2d(sb, &map_b, &full_bar[stage], kb * kBlockK, n_blk * args.block_n);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(kBlockM, args.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * args.block_n);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + kATileBytes;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
          const int ksteps = min(kBlockK, args.K - kb * kBlockK + 15) / 16;  // skip all-zero K tails
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
            umma_f16(tmem_d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kb | ks) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int epi_warp = warp - 2;
    const int lane_grp = warp & 3;        // TMEM lanes [32*lane_grp, +32) are accessible to this warp
    const int chunk_par = epi_warp >> 2;  // 0: even 32-column chunks, 1: odd
    const int epi_tid = threadIdx.x - 64;
    const bool has_cols = (args.bias != nullptr) || (args.col_sub != nullptr);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(args, tile, m_blk, n_blk);
      const int64_t n_base = (int64_t)n_blk * args.block_n;
      float* sb = s_bias + acc * kMaxBlockN;
      if (has_cols) {  // stage this tile's per-column constants once (one global round trip per tile)
        for (int c = epi_tid; c < args.block_n; c += kEpiWarps * 32) {
          const int64_t n = n_base + c;
          float v = 0.f;
          if (n < args.N) {
            if (args.bias) v += __ldg(&args.bias[n]);
            if (args.col_sub) v -= __ldg(&args.col_sub[n]);
          }
          sb[c] = v;
        }
      }
      named_bar_sync_1<kEpiWarps * 32>();
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int64_t row = (int64_t)m_blk * kBlockM + lane_grp * 32 + lane;
      const bool row_ok = row < args.M;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * args.block_n);
      for (int c = chunk_par * 32; c < args.block_n; c += 64) {
        const int width = min(32, args.block_n - c);  // 32, or 16 for the last chunk (block_n % 16 == 0)
        uint32_t v[32];
        if (width == 32) {
          tmem_ld_x32(taddr0 + (uint32_t)c, v);
        } else {
          uint32_t lo[16];
          tmem_ld_x16(taddr0 + (uint32_t)c, lo);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = lo[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0u;
        }
        // residual for this chunk: issue the loads before waiting on TMEM
        const int64_t n0 = n_base + c;
        uint4 rres[4];
        const bool full_bf16 = !args.d_is_f32 && row_ok && (n0 + width <= args.N);
        if (args.residual && full_bf16) {
          const uint4* r = reinterpret_cast<const uint4*>(args.residual + row * args.ld_res + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j * 8 < width) rres[j] = __ldg(r + j);
        }
        tmem_ld_wait();
        // TMA-store path: the whole warp stages its 32 x 32 sub-tile (rows / columns outside D are clipped by
        // the store), so only warp-uniform conditions may skip; the direct path skips per thread
        const bool via_tma = args.tma_store && width == 32 && (!args.residual || n0 + 32 <= args.N);
        if (n0 >= args.N || (via_tma ? (row - lane >= args.M) : !row_ok)) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (args.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= args.alpha;
        }
        if (has_cols) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + c + j);  // smem broadcast
            f[j] += b4.x;
            f[j + 1] += b4.y;
            f[j + 2] += b4.z;
            f[j + 3] += b4.w;
          }
        }
        if (args.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = relu6f(f[j]);
        } else if (args.act == 2) {  // ReLU
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        } else if (args.act == 3) {  // HardSwish: x * clamp(x / 6 + 0.5, 0, 1)
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= fminf(fmaxf(fmaf(f[j], 1.0f / 6.0f, 0.5f), 0.f), 1.f);
        }
        if (args.chunk_max && row_ok) {  // k-NN: the maximum of this 32-column chunk (columns >= N excluded)
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < width && n0 + j < args.N) m = fmaxf(m, f[j]);
          args.chunk_max[row * args.ld_cm + (n0 >> 5)] = m;
        }
        if (via_tma) {
          if (args.residual && full_bf16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rres[j]);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 rv = __bfloat1622float2(h2[q]);
                f[j * 8 + 2 * q] += rv.x;
                f[j * 8 + 2 * q + 1] += rv.y;
              }
            }
          }
          // the previous store issued from this warp's staging buffer must have read it
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          const uint32_t stg = smem_u32(store_stage + epi_warp * store_stage_bytes);
          if (args.d_is_f32) {  // 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row r at (j ^ (r & 7))
            const uint32_t rowa = stg + ((uint32_t)lane << 7);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((((uint32_t)j) ^ ((uint32_t)lane & 7u)) << 4)),
                           "f"(f[4 * j]), "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                           : "memory");
            }
          } else {  // 64-byte rows, SWIZZLE_64B: chunk j of row r at (j ^ ((r >> 1) & 3))
            const uint32_t rowa = stg + ((uint32_t)lane << 6);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                               rowa + ((((uint32_t)j) ^ (((uint32_t)lane >> 1) & 3u)) << 4)),
                           "r"(pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1])), "r"(pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3])),
                           "r"(pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5])), "r"(pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]))
                           : "memory");
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_d, stg, (int)n0, (int)(row));  // row of lane 0 == first row of the sub-tile
            tma_store_commit();
          }
        } else if (full_bf16) {
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + row * args.ldd + n0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j * 8 < width) {
              if (args.residual) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rres[j]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 rv = __bfloat1622float2(h2[q]);
                  f[j * 8 + 2 * q] += rv.x;
                  f[j * 8 + 2 * q + 1] += rv.y;
                }
              }
              uint4 o;
              o.x = pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]);
              o.y = pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]);
              o.z = pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]);
              o.w = pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]);
              reinterpret_cast<uint4*>(d)[j] = o;
            }
          }
        } else if (args.d_is_f32) {
          float* d = reinterpret_cast<float*>(args.D) + row * args.ldd + n0;
          if (n0 + width <= args.N && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < width) *reinterpret_cast<float4*>(d + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            const int nvalid = (int)min((int64_t)width, args.N - n0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) d[j] = f[j];
          }
        } else {  // ragged bf16 tail (N not a multiple of 16): scalar, never hit by the encoder
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + row * args.ldd + n0;
          const int nvalid = (int)min((int64_t)width, args.N - n0);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              float o = f[j];
              if (args.residual) o += __bfloat162float(args.residual[row * args.ld_res + n0 + j]);
              d[j] = __float2bfloat16_rn(o);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  if (args.tma_store && warp >= 2 && lane == 0) tma_store_wait_all();  // this lane issued the warp's tile stores
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- SIMT reference (self test only)
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A, int64_t M, int64_t lda,
                                 const __nv_bfloat16* __restrict__ B, int64_t N, int64_t ldb, int K,
                                 KernelArgs args) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(__bfloat162float(A[m * lda + k]), __bfloat162float(B[n * ldb + k]), acc);
  float v = acc * args.alpha;
  if (args.bias) v += args.bias[n];
  if (args.col_sub) v -= args.col_sub[n];
  if (args.act == 1) v = relu6f(v);
  else if (args.act == 2) v = fmaxf(v, 0.f);
  else if (args.act == 3) v *= fminf(fmaxf(fmaf(v, 1.0f / 6.0f, 0.5f), 0.f), 1.f);
  if (args.d_is_f32) {
    reinterpret_cast<float*>(args.D)[m * args.ldd + n] = v;
  } else {
    if (args.residual) v += __bfloat162float(args.residual[m * args.ld_res + n]);
    reinterpret_cast<__nv_bfloat16*>(args.D)[m * args.ldd + n] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;
static bool g_attr_set = false;
static std::mutex g_attr_mu;

static EncodeTiledFn get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    cudaGetLastError();
  });
  return g_encode;
}

bool available() { return ensure_init() == AM_OK && device_cc() / 10 == 10 && get_encode() != nullptr; }

static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t ld_elems, int K, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld ld=%lld K=%d box_rows=%d", (int)r, (long long)rows,
              (long long)ld_elems, K, box_rows);
    return AM_ERR_CUDA;
  }
  return AM_OK;
}

// generic bf16 tiled tensor map (rank <= 4), SWIZZLE_128B, zero OOB fill (used by fused_block.cu)
int encode_map_bf16(void* map_out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  AM_CHECK(get_encode() != nullptr, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t d[4], st[3];
  cuuint32_t b[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = get_encode()(reinterpret_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                            const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r);
    return AM_ERR_CUDA;
  }
  return AM_OK;
}

static int pick_block_n(int64_t N) {
  const int64_t n16 = (int64_t)round_up((size_t)N, 16);
  if (n16 <= kMaxBlockN) return (int)n16;
  const int64_t tiles = (n16 + kMaxBlockN - 1) / kMaxBlockN;
  return (int)round_up((size_t)((n16 + tiles - 1) / tiles), 16);
}

static KernelArgs make_args(int64_t M, int64_t N, int K, void* D, int64_t ldd, bool d_is_f32, const Epilogue& ep,
                            bool m_fastest) {
  KernelArgs a{};
  a.M = M;
  a.N = N;
  a.K = K;
  a.block_n = pick_block_n(N);
  if (ep.chunk_max) a.block_n = std::min(kMaxBlockN, (int)round_up((size_t)a.block_n, 32));  // chunk c <-> columns [32c, 32c + 32)
  a.tiles_m = (int)((M + kBlockM - 1) / kBlockM);
  a.tiles_n = (int)((N + a.block_n - 1) / a.block_n);
  a.m_fastest = m_fastest ? 1 : 0;
  a.D = D;
  a.ldd = ldd;
  a.d_is_f32 = d_is_f32 ? 1 : 0;
  a.alpha = ep.alpha;
  a.bias = ep.bias;
  a.col_sub = ep.col_sub;
  a.act = ep.act;
  a.chunk_max = ep.chunk_max;
  a.ld_cm = ep.ld_cm;
  a.residual = ep.residual;
  a.ld_res = ep.ld_res;
  return a;
}

int gemm_bf16(const __nv_bfloat16* A, int64_t M, int64_t lda, const __nv_bfloat16* B, int64_t N, int64_t ldb, int K,
              void* D, int64_t ldd, bool d_is_f32, const Epilogue& ep, bool m_fastest, cudaStream_t st) {
  AM_CHECK(A && B && D, "gemm: NULL operand");
  AM_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%lld N=%lld K=%d", (long long)M, (long long)N, K);
  AM_CHECK(lda % 8 == 0 && ldb % 8 == 0, "gemm: lda/ldb must be multiples of 8 elements (TMA 16-byte pitch)");
  AM_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
           "gemm: operands must be 16-byte aligned");
  AM_CHECK(available(), "gemm: tcgen05 path unavailable (needs sm_100 and cuTensorMapEncodeTiled)");
  KernelArgs args = make_args(M, N, K, D, ldd, d_is_f32, ep, m_fastest);
  CUtensorMap map_a, map_b, map_d;
  AM_TRY(make_map(&map_a, A, M, lda, K, kBlockM));
  AM_TRY(make_map(&map_b, B, N, ldb, K, args.block_n));
  // D through TMA tile stores when its pitch / base allow a tensor map (16-byte multiples)
  const size_t esz = d_is_f32 ? 4 : 2;
  static const bool no_tma_store = std::getenv("AM_GEMM_NO_TMA_STORE") != nullptr;
  args.tma_store = (!no_tma_store && (ldd * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0 &&
                    M < (int64_t)1 << 31 && N < (int64_t)1 << 31) ? 1 : 0;
  if (args.tma_store) {
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)ldd * esz};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&map_d, d_is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, D,
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              d_is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) args.tma_store = 0;  // odd pitch etc.: the direct-store epilogue handles it
  }
  if (!args.tma_store) map_d = map_a;
  const size_t stage_bytes = (size_t)kATileBytes + (size_t)args.block_n * kBlockK * 2;
  const size_t tail = (args.tma_store ? (size_t)kEpiWarps * (d_is_f32 ? kStoreStageBytes : kStoreStageBytes / 2) : 0) +
                      1024 + 256 + 2 * kMaxBlockN * 4;
  constexpr size_t kSmemMax = 232448;
  args.stages = (kStages * stage_bytes + tail <= kSmemMax) ? kStages : kStages - 1;
  const size_t smem = (size_t)args.stages * stage_bytes + tail;
  {
    std::lock_guard<std::mutex> lk(g_attr_mu);
    if (!g_attr_set) {
      AM_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
      g_attr_set = true;
    }
  }
  const int tiles = args.tiles_m * args.tiles_n;
  const int grid = std::max(1, std::min(tiles, sm_count()));
  AM_LAUNCH(gemm_tcgen05_kernel, grid, kThreads, smem, st, map_a, map_b, map_d, args);
  return AM_OK;
}

int gemm_bf16_simt(const __nv_bfloat16* A, int64_t M, int64_t lda, const __nv_bfloat16* B, int64_t N, int64_t ldb,
                   int K, void* D, int64_t ldd, bool d_is_f32, const Epilogue& ep, cudaStream_t st) {
  AM_CHECK(A && B && D && M > 0 && N > 0 && K > 0, "gemm_simt: bad problem");
  AM_CHECK(M <= 65535, "gemm_simt: M too large for the self-test kernel");
  KernelArgs args = make_args(M, N, K, D, ldd, d_is_f32, ep, false);
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)M);
  AM_LAUNCH(gemm_simt_kernel, grid, 128, 0, st, A, M, lda, B, N, ldb, K, args);
  return AM_OK;
}

}  // namespace gemm
}  // namespace am
