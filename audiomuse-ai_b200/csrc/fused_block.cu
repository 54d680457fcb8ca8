// Fused PhiNet inverted-residual block (sm_100a):
//
//     Y = [X +] project( relu6( dw3x3_s( relu6( expand(X) ) ) ) )        all BatchNorms folded
//
// in ONE kernel, so the expanded tensor (up to 5.8x the block input) never touches HBM.
// Replaces, for the wide early blocks, the three-kernel sequence GEMM -> depthwise -> GEMM of
// encoder.cu (student_clap/models/student_onnx_model.py:95-148 block shape; micromind PhiNetConvBlock).
//
// One CTA (12 warps) owns TH output rows x the full width of one window:
//   TMA      X halo tile  [(TH-1)s+3 rows x W x Cin]  -> smem, K-major SWIZZLE_128B (zero rows = conv padding)
//   per 64-channel chunk j of the expanded dimension:
//     tcgen05.mma   D1[t] = X[t] . W1_j^T            (M1 halo pixels in 128-row tiles, N = 64, TMEM)
//     epilogue 1    TMEM -> +b1, ReLU6, zero outside the image -> bf16 -> smem E (same swizzled layout)
//     depthwise     3x3 stride s over E (+bd, ReLU6) -> bf16 -> smem A2, written directly in the
//                   K-major SWIZZLE_128B operand layout (fence.proxy.async before the MMA reads it)
//     tcgen05.mma   D2 += A2 . W2_j^T                (M = 128 output pixels, N = Cout, TMEM)
//   epilogue 2      TMEM -> +b2 (+ residual read from the X tile in smem) -> bf16 -> global Y
// Weights stream through a 2-stage TMA ring; the MMA of chunk j+1's expansion is issued before the
// CUDA-core phases of chunk j+1 start, so tensor and CUDA-core work of adjacent chunks overlap.
#include <algorithm>

#include "fused_block.cuh"
#include "gemm_tcgen05.cuh"
#include "ptx_sm100.cuh"

namespace am {
namespace fused {

using namespace ptx;

constexpr int kThreads = 384;
constexpr int kWarps = kThreads / 32;
constexpr int kCK = 64;                 // expanded channels per chunk = one 128-byte swizzle row
constexpr int kTileBytes = 128 * 128;   // one [128 rows x 64 ch] bf16 operand tile
constexpr int kTmemCols = 512;

struct Args {
  int B, H, W, Ho, Wo, stride;
  int cin_p, cmid_p, cout_p;
  int has_expand, residual;
  int TH, IH, M1, m1_tiles, M2;
  int kb_in, n_chunks;
  int tiles_per_window, total_tiles;
  const float* b1;   // [cmid_p]
  const float* wd;   // [9, cmid_p]
  const float* bd;   // [cmid_p]
  const float* b2;   // [cout_p]
  __nv_bfloat16* Y;  // [B, Ho, Wo, cout_p]
  // smem byte offsets (from the 1024-aligned base)
  uint32_t off_x, off_e, off_a2, off_w1, off_w2, off_small, off_bar;
  uint32_t w1_stage_bytes, w2_stage_bytes;
};

struct SmallVecs {   // staged per chunk
  float b1[kCK];
  float bd[kCK];
  float wd[9][kCK];
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1)
fused_block_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                   const __grid_constant__ CUtensorMap map_w2, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_x = smem + a.off_x;
  uint8_t* s_e = smem + a.off_e;
  uint8_t* s_a2 = smem + a.off_a2;
  uint8_t* s_w1 = smem + a.off_w1;
  uint8_t* s_w2 = smem + a.off_w2;
  SmallVecs* sv = reinterpret_cast<SmallVecs*>(smem + a.off_small);
  float* s_b2 = reinterpret_cast<float*>(sv + 1);                        // [cout_p]
  uint64_t* bar_x = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint64_t* bar_w = bar_x + 1;       // [2]
  uint64_t* bar_mma1 = bar_w + 2;
  uint64_t* bar_mma2 = bar_mma1 + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_mma2 + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lane_grp = warp & 3;     // TMEM lanes [32*lane_grp, +32)
  const int grp_rank = warp >> 2;    // 0..2: which of the three warps sharing that lane group

  if (tid == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    prefetch_tensormap(&map_w2);
    mbar_init(bar_x, 1);
    mbar_init(&bar_w[0], 1);
    mbar_init(&bar_w[1], 1);
    mbar_init(bar_mma1, 1);
    mbar_init(bar_mma2, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  for (int i = tid; i < a.cout_p; i += kThreads) s_b2[i] = a.b2[i];
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_d2 = tmem_base + (uint32_t)(a.has_expand ? a.m1_tiles * kCK : 0);

  // running completion counters -> mbarrier parities (uniform across the CTA)
  uint32_t n_x = 0, n_w0 = 0, n_w1 = 0, n_m1 = 0, n_m2 = 0;

  const uint32_t idesc1 = make_idesc(128, kCK);
  const uint32_t idesc2 = make_idesc(128, a.cout_p);
  const uint32_t x_kb_bytes = (uint32_t)a.m1_tiles * kTileBytes;          // smem pitch of one X k-block
  const uint32_t x_box_bytes = (uint32_t)a.M1 * 128u;                     // bytes one TMA box delivers
  const uint32_t w_bytes = (a.has_expand ? (uint32_t)a.kb_in * kCK * 128u : 0u) + (uint32_t)a.cout_p * 128u;

  auto load_weights = [&](int j) {  // thread 0 only: W1 chunk j (kb_in boxes) + W2 chunk j -> stage j & 1
    const int stg = j & 1;
    mbar_expect_tx(&bar_w[stg], w_bytes);
    if (a.has_expand)
      for (int kb = 0; kb < a.kb_in; ++kb)
        tma_load_2d(s_w1 + stg * a.w1_stage_bytes + kb * (kCK * 128), &map_w1, &bar_w[stg], kb * 64, j * kCK);
    tma_load_2d(s_w2 + stg * a.w2_stage_bytes, &map_w2, &bar_w[stg], j * kCK, 0);
  };
  auto wait_weights = [&](int j) {  // thread 0 only
    if (j & 1) {
      mbar_wait(&bar_w[1], n_w1 & 1);
    } else {
      mbar_wait(&bar_w[0], n_w0 & 1);
    }
  };
  auto issue_mma1 = [&](int j) {  // thread 0 only: D1[t] = X[t] . W1_j^T for every halo M-tile
    const int stg = j & 1;
    for (int t = 0; t < a.m1_tiles; ++t) {
      const uint32_t d = tmem_base + (uint32_t)(t * kCK);
      for (int kb = 0; kb < a.kb_in; ++kb) {
        const uint64_t da = make_smem_desc(smem_u32(s_x + kb * x_kb_bytes + t * kTileBytes));
        const uint64_t db = make_smem_desc(smem_u32(s_w1 + stg * a.w1_stage_bytes + kb * (kCK * 128)));
        const int ksteps = min(64, a.cin_p - kb * 64 + 15) / 16;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_f16(d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc1, (kb | ks) ? 1u : 0u);
      }
    }
    umma_commit(bar_mma1);
  };

  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_window;
    const int ho0 = (tile - b * a.tiles_per_window) * a.TH;
    const int h0 = ho0 * a.stride - 1;  // first input row of the halo tile (may be -1: zero filled)

    // ---------------- tile prologue: X halo tile + weights of chunk 0, first expansion MMA
    if (tid == 0) {
      mbar_expect_tx(bar_x, x_box_bytes * (uint32_t)a.kb_in);
      for (int kb = 0; kb < a.kb_in; ++kb) tma_load_4d(s_x + kb * x_kb_bytes, &map_x, bar_x, kb * 64, 0, h0, b);
      load_weights(0);
      mbar_wait(bar_x, n_x & 1);
      if (a.has_expand) {
        wait_weights(0);
        tcgen05_fence_after();
        issue_mma1(0);
      }
    }
    if (tid != 0) mbar_wait(bar_x, n_x & 1);  // everyone needs X (depthwise of block 1, residual)
    n_x++;

    for (int j = 0; j < a.n_chunks; ++j) {
      const int c_base = j * kCK;
      // ---- A. per-chunk constants
      for (int i = tid; i < kCK; i += kThreads) {
        const int c = c_base + i;
        const bool ok = c < a.cmid_p;
        sv->b1[i] = (ok && a.has_expand) ? a.b1[c] : 0.f;
        sv->bd[i] = ok ? a.bd[c] : 0.f;
      }
      for (int i = tid; i < 9 * kCK; i += kThreads) {
        const int t = i / kCK, c = c_base + (i - t * kCK);
        sv->wd[t][i - t * kCK] = c < a.cmid_p ? a.wd[t * a.cmid_p + c] : 0.f;
      }
      __syncthreads();  // S1

      // ---- B. expansion epilogue: TMEM -> +b1, ReLU6, zero outside the image -> bf16 -> E (swizzled)
      const uint8_t* dw_src = s_x + (size_t)j * x_kb_bytes;  // no-expand block: depthwise reads X k-block j
      if (a.has_expand) {
        mbar_wait(bar_mma1, n_m1 & 1);
        n_m1++;
        tcgen05_fence_after();
        const int items = a.m1_tiles * 2;  // (M-tile, 32-column half)
        for (int it = grp_rank; it < items; it += 3) {
          const int t = it >> 1, half = it & 1;
          uint32_t v[32];
          tmem_ld_x32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t * kCK + half * 32), v);
          tmem_ld_wait();
          const int p = t * 128 + lane_grp * 32 + lane;  // halo pixel
          const int ih = p / a.W;
          const bool inside = (p < a.M1) && (h0 + ih >= 0) && (h0 + ih < a.H);
          uint8_t* dst = s_e + ((uint32_t)p >> 3) * 1024u + ((uint32_t)p & 7u) * 128u;  // row base; chunk XOR below
          const uint32_t r7 = (uint32_t)p & 7u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (inside) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                f[e] = relu6f(__uint_as_float(v[q * 8 + e]) + sv->b1[half * 32 + q * 8 + e]);
              o.x = pack2(f[0], f[1]);
              o.y = pack2(f[2], f[3]);
              o.z = pack2(f[4], f[5]);
              o.w = pack2(f[6], f[7]);
            }
            const uint32_t chunk = (uint32_t)(half * 4 + q);
            *reinterpret_cast<uint4*>(dst + ((chunk ^ r7) << 4)) = o;
          }
        }
        dw_src = s_e;
        tcgen05_fence_before();
      }
      __syncthreads();  // S2: E complete

      // ---- D. the previous chunk's projection MMA must have finished reading A2 / its W2 stage
      if (j > 0) {
        mbar_wait(bar_mma2, n_m2 & 1);
        n_m2++;
      }
      if (tid == 0 && j + 1 < a.n_chunks) load_weights(j + 1);

      // ---- E. depthwise 3x3 (+bd, ReLU6) -> A2 in the MMA operand layout
      for (int it = tid; it < a.M2 * 8; it += kThreads) {
        const int g = it & 7, o = it >> 3;
        const int oh = o / a.Wo, ow = o - oh * a.Wo;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = sv->bd[g * 8 + e];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int ih = oh * a.stride + dy;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int iw = ow * a.stride + dx - 1;
            if (iw < 0 || iw >= a.W) continue;
            const uint32_t p = (uint32_t)(ih * a.W + iw);
            const uint4 raw = *reinterpret_cast<const uint4*>(dw_src + sw128_offset(p, (uint32_t)g));
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
            const float* wv = &sv->wd[dy * 3 + dx][g * 8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f2 = __bfloat1622float2(h2[q]);
              acc[2 * q] = fmaf(f2.x, wv[2 * q], acc[2 * q]);
              acc[2 * q + 1] = fmaf(f2.y, wv[2 * q + 1], acc[2 * q + 1]);
            }
          }
        }
        uint4 pk;
        pk.x = pack2(relu6f(acc[0]), relu6f(acc[1]));
        pk.y = pack2(relu6f(acc[2]), relu6f(acc[3]));
        pk.z = pack2(relu6f(acc[4]), relu6f(acc[5]));
        pk.w = pack2(relu6f(acc[6]), relu6f(acc[7]));
        *reinterpret_cast<uint4*>(s_a2 + sw128_offset((uint32_t)o, (uint32_t)g)) = pk;
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncthreads();      // S3

      // ---- F. projection MMA of this chunk, then the expansion MMA of the next one
      if (tid == 0) {
        wait_weights(j);
        tcgen05_fence_after();
        const int stg = j & 1;
        const uint64_t da = make_smem_desc(smem_u32(s_a2));
        const uint64_t db = make_smem_desc(smem_u32(s_w2 + stg * a.w2_stage_bytes));
        const int ksteps = min(64, a.cmid_p - c_base + 15) / 16;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_f16(tmem_d2, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc2, (j | ks) ? 1u : 0u);
        umma_commit(bar_mma2);
        if (a.has_expand && j + 1 < a.n_chunks) {
          wait_weights(j + 1);
          tcgen05_fence_after();
          issue_mma1(j + 1);
        }
      }
      if (j & 1) n_w1++; else n_w0++;
    }

    // ---------------- tile epilogue: D2 -> +b2 (+ residual from the X tile) -> bf16 -> Y
    mbar_wait(bar_mma2, n_m2 & 1);
    n_m2++;
    tcgen05_fence_after();
    {
      const int o = lane_grp * 32 + lane;  // output pixel of this thread's TMEM lane
      const int oh = o / a.Wo, ow = o - oh * a.Wo;
      const int ho = ho0 + oh;
      const bool valid = (o < a.M2) && (ho < a.Ho);
      const int n_items = (a.cout_p + 31) / 32;
      __nv_bfloat16* yrow = a.Y + (((int64_t)b * a.Ho + ho) * a.Wo + ow) * a.cout_p;
      const uint32_t pc = (uint32_t)((oh + 1) * a.W + ow);  // centre input pixel (stride-1 residual blocks)
      for (int it = grp_rank; it < n_items; it += 3) {
        const int c0 = it * 32;
        const int width = min(32, a.cout_p - c0);  // 32 or 16 (cout_p % 16 == 0)
        uint32_t v[32];
        if (width == 32) {
          tmem_ld_x32(tmem_d2 + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, v);
        } else {
          uint32_t lo[16];
          tmem_ld_x16(tmem_d2 + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)c0, lo);
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = lo[e];
#pragma unroll
          for (int e = 16; e < 32; ++e) v[e] = 0u;
        }
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 8 < width) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]) + s_b2[c0 + q * 8 + e];
              if (a.residual) {
                const int c = c0 + q * 8;
                const uint4 raw = *reinterpret_cast<const uint4*>(s_x + (size_t)(c >> 6) * x_kb_bytes +
                                                                  sw128_offset(pc, (uint32_t)((c & 63) >> 3)));
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f2 = __bfloat1622float2(h2[e]);
                  f[2 * e] += f2.x;
                  f[2 * e + 1] += f2.y;
                }
              }
              uint4 pk;
              pk.x = pack2(f[0], f[1]);
              pk.y = pack2(f[2], f[3]);
              pk.z = pack2(f[4], f[5]);
              pk.w = pack2(f[6], f[7]);
              *reinterpret_cast<uint4*>(yrow + c0 + q * 8) = pk;
            }
          }
        }
      }
    }
    tcgen05_fence_before();
    __syncthreads();  // S4: X tile / TMEM free for the next tile
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- host
static size_t layout_smem(Args& a) {
  size_t off = 0;
  a.off_x = (uint32_t)off;
  off += (size_t)a.kb_in * a.m1_tiles * kTileBytes;
  a.off_e = (uint32_t)off;
  if (a.has_expand) off += (size_t)a.m1_tiles * kTileBytes;
  a.off_a2 = (uint32_t)off;
  off += kTileBytes;
  a.off_w1 = (uint32_t)off;
  a.w1_stage_bytes = a.has_expand ? (uint32_t)a.kb_in * kCK * 128u : 0u;
  off += 2 * (size_t)a.w1_stage_bytes;
  a.off_w2 = (uint32_t)off;
  a.w2_stage_bytes = (uint32_t)round_up((size_t)a.cout_p * 128u, 1024);
  off += 2 * (size_t)a.w2_stage_bytes;
  a.off_small = (uint32_t)off;
  off += sizeof(SmallVecs) + (size_t)a.cout_p * 4;
  off = round_up(off, 16);
  a.off_bar = (uint32_t)off;
  off += 64;
  return off + 1024;  // alignment slack
}

bool plan(const BlockDesc& d, Plan* out) {
  if (d.cout_p > 256 || d.cout_p % 16 || d.cin_p % 16 || d.cmid_p % 16) return false;
  if (d.W > 256 || d.W < 1) return false;
  if (!d.has_expand && d.cmid_p != d.cin_p) return false;
  if (d.residual && (d.stride != 1 || d.cin_p != d.cout_p)) return false;
  const int Ho = (d.H + 2 - 3) / d.stride + 1, Wo = (d.W + 2 - 3) / d.stride + 1;
  const int kb_in = (d.cin_p + 63) / 64;
  for (int TH = std::min(Ho, 128 / std::max(Wo, 1)); TH >= 1; --TH) {
    Args a{};
    a.has_expand = d.has_expand;
    a.cout_p = d.cout_p;
    a.kb_in = kb_in;
    a.IH = (TH - 1) * d.stride + 3;
    a.M1 = a.IH * d.W;
    a.m1_tiles = (a.M1 + 127) / 128;
    if (a.IH > 256) continue;
    const int tmem = (d.has_expand ? a.m1_tiles * kCK : 0) + d.cout_p;
    if (tmem > kTmemCols) continue;
    const size_t smem = layout_smem(a);
    if (smem > 220 * 1024) continue;
    out->TH = TH;
    out->smem_bytes = smem;
    return true;
  }
  return false;
}

int run(const BlockDesc& d, const Plan& p, const __nv_bfloat16* X, const __nv_bfloat16* W1, const float* b1,
        const float* wd, const float* bd, const __nv_bfloat16* W2, const float* b2, __nv_bfloat16* Y, int B,
        cudaStream_t st) {
  Args a{};
  a.B = B;
  a.H = d.H;
  a.W = d.W;
  a.stride = d.stride;
  a.Ho = (d.H + 2 - 3) / d.stride + 1;
  a.Wo = (d.W + 2 - 3) / d.stride + 1;
  a.cin_p = d.cin_p;
  a.cmid_p = d.cmid_p;
  a.cout_p = d.cout_p;
  a.has_expand = d.has_expand;
  a.residual = d.residual;
  a.TH = p.TH;
  a.IH = (p.TH - 1) * d.stride + 3;
  a.M1 = a.IH * d.W;
  a.m1_tiles = (a.M1 + 127) / 128;
  a.M2 = p.TH * a.Wo;
  a.kb_in = (d.cin_p + 63) / 64;
  a.n_chunks = (d.cmid_p + kCK - 1) / kCK;
  a.tiles_per_window = (a.Ho + p.TH - 1) / p.TH;
  a.total_tiles = a.tiles_per_window * B;
  a.b1 = b1;
  a.wd = wd;
  a.bd = bd;
  a.b2 = b2;
  a.Y = Y;
  const size_t smem = layout_smem(a);
  AM_CHECK(smem == p.smem_bytes, "fused block: plan / launch smem mismatch");

  CUtensorMap mx, mw1, mw2;
  {
    const uint64_t dims[4] = {(uint64_t)d.cin_p, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)d.cin_p * 2, (uint64_t)d.W * d.cin_p * 2, (uint64_t)d.H * d.W * d.cin_p * 2};
    const uint32_t box[4] = {64, (uint32_t)d.W, (uint32_t)a.IH, 1};
    AM_TRY(gemm::encode_map_bf16(&mx, X, 4, dims, str, box));
  }
  if (d.has_expand) {
    const uint64_t dims[2] = {(uint64_t)d.cin_p, (uint64_t)d.cmid_p};
    const uint64_t str[1] = {(uint64_t)d.cin_p * 2};
    const uint32_t box[2] = {64, (uint32_t)kCK};
    AM_TRY(gemm::encode_map_bf16(&mw1, W1, 2, dims, str, box));
  } else {
    mw1 = mx;  // unused
  }
  {
    const uint64_t dims[2] = {(uint64_t)d.cmid_p, (uint64_t)d.cout_p};
    const uint64_t str[1] = {(uint64_t)d.cmid_p * 2};
    const uint32_t box[2] = {64, (uint32_t)d.cout_p};
    AM_TRY(gemm::encode_map_bf16(&mw2, W2, 2, dims, str, box));
  }
  static size_t attr = 0;
  if (smem > attr) {
    AM_CUDA(cudaFuncSetAttribute(fused_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int grid = std::max(1, std::min(a.total_tiles, sm_count()));
  AM_LAUNCH(fused_block_kernel, grid, kThreads, smem, st, mx, mw1, mw2, a);
  return AM_OK;
}

}  // namespace fused
}  // namespace am
