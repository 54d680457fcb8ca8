// K2 + K3: student CLAP audio encoder, batched (sm_100a).
//
// Replaces the per-segment onnxruntime call of tasks/clap_analyzer.py:534 and the numpy pooling
// at :552-562.  Architecture: student_clap/models/student_onnx_model.py (bn0 over mel bins ->
// PhiNet inverted-residual trunk (ReLU6, no SE: compatibility=True) -> 1x1 stride-2 conv to 2048
// -> spatial mean -> Projection (linear1, GELU, linear2, residual, LayerNorm) -> L2).
//
// Data layout in HBM: activations are NHWC bf16 [B, H, W, Cp] with Cp = channels rounded up to 16
// (padded channels are exact zeros: zero weight rows + zero bias), so every 1x1 convolution is
// one K-major GEMM  D[B*H*W, Cout] = A[B*H*W, Cin] x W[Cout, Cin]^T  on the tcgen05 path
// (gemm.cu) with the folded-BatchNorm bias, ReLU6 and the residual add fused in its epilogue.
// Depthwise 3x3 (+BN+ReLU6) is a bandwidth kernel over 8-channel (16-byte) vectors.  The stem
// (bn0 + pad + 3x3 stride-2 on the single input channel + 1x1 + BN + ReLU6) reads the fp32
// log-mel directly and is computed in fp32.  The head (<= 2.5 MMAC per window) runs in fp32.
#include "common.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>

#include "encoder_generic.cuh"
#include "fused_block.cuh"
#include "gemm_tcgen05.cuh"
#include "model_spec.cuh"

namespace am {

struct Layer {
  int type = 0;
  int cin = 0, cout = 0, cin_p = 0, cout_p = 0;
  int kh = 1, kw = 1;
  int stride = 1, act = 0, block_start = 0, residual = 0;
  int pad_t = 0, pad_b = 0, pad_l = 0, pad_r = 0;
  int h_is_time = 1, gate_act = 0, cmid = 0;
  DevBuf<__nv_bfloat16> w_bf16;  // pointwise: [cout_p, cin_p]
  DevBuf<__half> w_f16;          // projection (act == 0) pointwise, same layout: the fused block's MMA2 runs fp16 x fp16
  DevBuf<uint32_t> tpack;        // depthwise of a block with an expansion conv: fusedt::pack_consts (taps, biases)
  DevBuf<__nv_bfloat16> w1s;     // stem: [cout_p, 16] rows (s_hi, s_hi, s_lo, t_hi, t_lo, 0 ...), the stem as an expansion GEMM (stem_x0_kernel)
  DevBuf<float> w_f32;           // depthwise: [k*k, c_p]; stem: dw[9]; first conv: [kh*kw, c_p]; squeeze-excite: fc1 [cmid, c]
  DevBuf<float> bias;            // [cout_p]  (squeeze-excite: fc1 bias [cmid])
  DevBuf<float> aux0, aux1, aux2;  // stem / first conv: per-mel scale / shift (+ stem pw scale); squeeze-excite: fc2 [c, cmid], fc2 bias
  // true for the 3x3 / pad 1 / ReLU6 depthwise the packed-fp16 kernels and the fused block kernel implement
  bool dw_fast() const {
    return type == kDepthwise && kh == 3 && kw == 3 && pad_t == 1 && pad_b == 1 && pad_l == 1 && pad_r == 1 && act == kActRelu6;
  }
};

// one operation of the head's row program (model_spec.cuh: VecOp) with its weights on the device
struct HeadOp {
  int kind = 0, a = -1, b = -1, dst = -1, K = 0, N = 0, act = 0, stride = 1;
  float eps = 0.f, eps2 = 0.f;
  DevBuf<float> w, bias;          // fp32 weights (SIMT path, LayerNorm gain / shift, affine scale / shift)
  DevBuf<__nv_bfloat16> w3;       // kVecLinear: bf16 [N, 3*Kp] = [hi | lo | hi] (see split3_kernel)
  bool has_w = false, has_bias = false;
};

static inline int pad16(int c) { return (int)round_up((size_t)c, 16); }

// ---------------------------------------------------------------- kernels
// stem: mel f32 [B, n_mels, T] -> NHWC 16-bit [B, Ho, Wo, Cp];  image H = time, W = mel bin.
// A CTA owns a tile of 32 output rows (time) x 8 output columns (mel) of one window.
//   phase 1: one thread per pixel computes the single-channel 3x3 stride-2 response; lanes of a warp
//            walk the TIME axis (contiguous in the mel layout), so each tap is a 256-byte strided read
//            instead of 32 scattered sectors;
//   phase 2: the 256 responses are expanded to Cp channels, consecutive threads writing consecutive
//            16-byte groups (8 pixels x Cp x 2 B contiguous runs per output row).
constexpr int kStemTH = 32, kStemTW = 8;

template <bool kHalfOut>
__global__ void __launch_bounds__(256)
stem_kernel(const float* __restrict__ mel, int B, int n_mels, int T, int Ho, int Wo, int pad_t, int pad_l,
            const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
            const float* __restrict__ dw, const float* __restrict__ pw_scale,
            const float* __restrict__ pw_shift, int cp, __nv_bfloat16* __restrict__ out) {
  __shared__ float s_v[kStemTH * kStemTW];
  extern __shared__ float s_pw[];  // [2 * cp]: scale, shift
  const int groups = cp >> 3;
  const int tiles_h = (Ho + kStemTH - 1) / kStemTH, tiles_w = (Wo + kStemTW - 1) / kStemTW;
  const int64_t n_tiles = (int64_t)B * tiles_h * tiles_w;
  for (int i = threadIdx.x; i < cp; i += 256) {
    s_pw[i] = pw_scale[i];
    s_pw[cp + i] = pw_shift[i];
  }
  float wdw[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wdw[i] = __ldg(&dw[i]);
  // phase-2 mapping: a thread keeps ONE channel group (its 8 scales + 8 shifts live in registers) and
  // walks the tile's pixels; consecutive threads = consecutive groups of one pixel, then the next pixel,
  // so a warp still writes contiguous 16-byte runs.  (Reading the 16 constants from shared memory per
  // item made the kernel LSU-bound at 3x its write roofline.)
  const int g_step = groups < 256 ? groups : 256;
  const int pix_lanes = 256 / g_step;
  const int g0 = (int)threadIdx.x % g_step, pl = (int)threadIdx.x / g_step;
  float sc[8], sh[8];
  int g_regs = -1;
  __syncthreads();
  if (pl < pix_lanes) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = s_pw[g0 * 8 + e];
      sh[e] = s_pw[cp + g0 * 8 + e];
    }
    g_regs = g0;
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const int64_t t1 = tile / tiles_w;
    const int th = (int)(t1 % tiles_h);
    const int b = (int)(t1 / tiles_h);
    const int ho0 = th * kStemTH, wo0 = tw * kStemTW;
    __syncthreads();
    {
      const int hl = threadIdx.x & (kStemTH - 1), wl = threadIdx.x / kStemTH;  // lanes walk the time axis
      const int ho = ho0 + hl, wo = wo0 + wl;
      float v = 0.f;
      if (ho < Ho && wo < Wo) {
        const float* m = mel + (int64_t)b * n_mels * T;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int w = 2 * wo + dx - pad_l;
          if (w < 0 || w >= n_mels) continue;
          const float bsc = __ldg(&bn_scale[w]), bsh = __ldg(&bn_shift[w]);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int h = 2 * ho + dy - pad_t;
            if (h < 0 || h >= T) continue;
            v = fmaf(wdw[dy * 3 + dx], fmaf(__ldg(&m[(int64_t)w * T + h]), bsc, bsh), v);
          }
        }
      }
      s_v[hl * kStemTW + wl] = v;
    }
    __syncthreads();
    const int nw = min(kStemTW, Wo - wo0), nh = min(kStemTH, Ho - ho0);
    for (int g = g0; g < groups && pl < pix_lanes; g += g_step) {
      if (g != g_regs) {  // only when groups > 256 (never for the shipped students)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          sc[e] = s_pw[g * 8 + e];
          sh[e] = s_pw[cp + g * 8 + e];
        }
        g_regs = g;
      }
      for (int p = pl; p < nh * nw; p += pix_lanes) {
        const int hl = (nw == kStemTW) ? (p >> 3) : p / nw;  // ragged right-edge tiles only
        const int wl = p - hl * nw;
        const float v = s_v[hl * kStemTW + wl];
      float o8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o8[e] = relu6f(fmaf(v, sc[e], sh[e]));
      uint4 pk;
      if constexpr (kHalfOut) {  // consumed only by the fused first block, whose depthwise runs in fp16
        __half2 t;
        t = __floats2half2_rn(o8[0], o8[1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(o8[2], o8[3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(o8[4], o8[5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(o8[6], o8[7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
      } else {
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(o8[0], o8[1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(o8[2], o8[3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(o8[4], o8[5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(o8[6], o8[7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
      }
      const int64_t pix = ((int64_t)b * Ho + ho0 + hl) * Wo + wo0 + wl;
      *reinterpret_cast<uint4*>(out + (pix * groups + g) * 8) = pk;
      }
    }
  }
}

// stem, first half only: the single-channel 3x3 stride-2 response v per output pixel, written as the 16-column bf16
// row  (v_hi, v_lo, v_hi, 1, 1, 0 ...)  of a [B, Ho, Wo, 16] tensor.  The channel-per-lane fused block then produces the
// stem's per-channel affine + ReLU6 as its "expansion" GEMM against rows (s_hi, s_hi, s_lo, t_hi, t_lo, 0 ...):
// v s + t to fp32-class accuracy on the tensor pipe (split bf16), and the 144-channel stem output never exists in HBM
// (it was 2.4 GB written + read per 256 windows).  Same tile walk as stem_kernel phase 1.
__global__ void __launch_bounds__(256)
stem_x0_kernel(const float* __restrict__ mel, int B, int n_mels, int T, int Ho, int Wo, int pad_t, int pad_l,
               const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, const float* __restrict__ dw,
               __nv_bfloat16* __restrict__ out) {
  __shared__ float s_v[kStemTH * kStemTW];
  const int tiles_h = (Ho + kStemTH - 1) / kStemTH, tiles_w = (Wo + kStemTW - 1) / kStemTW;
  const int64_t n_tiles = (int64_t)B * tiles_h * tiles_w;
  float wdw[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) wdw[i] = __ldg(&dw[i]);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const int64_t t1 = tile / tiles_w;
    const int th = (int)(t1 % tiles_h);
    const int b = (int)(t1 / tiles_h);
    const int ho0 = th * kStemTH, wo0 = tw * kStemTW;
    __syncthreads();
    {
      const int hl = threadIdx.x & (kStemTH - 1), wl = threadIdx.x / kStemTH;  // lanes walk the time axis
      const int ho = ho0 + hl, wo = wo0 + wl;
      float v = 0.f;
      if (ho < Ho && wo < Wo) {
        const float* m = mel + (int64_t)b * n_mels * T;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int w = 2 * wo + dx - pad_l;
          if (w < 0 || w >= n_mels) continue;
          const float bsc = __ldg(&bn_scale[w]), bsh = __ldg(&bn_shift[w]);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int h = 2 * ho + dy - pad_t;
            if (h < 0 || h >= T) continue;
            v = fmaf(wdw[dy * 3 + dx], fmaf(__ldg(&m[(int64_t)w * T + h]), bsc, bsh), v);
          }
        }
      }
      s_v[hl * kStemTW + wl] = v;
    }
    __syncthreads();
    {
      const int hl = threadIdx.x >> 3, wl = threadIdx.x & 7;   // consecutive threads = consecutive pixels of a row: 256 B runs
      const int ho = ho0 + hl, wo = wo0 + wl;
      if (ho < Ho && wo < Wo) {
        const float v = s_v[hl * kStemTW + wl];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const uint32_t uh = (uint32_t)__bfloat16_as_ushort(hi), ul = (uint32_t)__bfloat16_as_ushort(lo);
        uint4* dst = reinterpret_cast<uint4*>(out + (((int64_t)b * Ho + ho) * Wo + wo) * 16);
        dst[0] = make_uint4(uh | (ul << 16), uh | (0x3f80u << 16), 0x3f80u, 0u);   // v_hi v_lo | v_hi 1 | 1 0 | 0 0
        dst[1] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

// depthwise 3x3, pad 1, stride s, folded BN, ReLU6.  NHWC bf16.
// A thread owns 8 channels (one 16-byte vector) of a vertical strip of kDwRows output rows at one
// output column: the 9 x 8 folded weights live in registers, each input row (3 taps) is loaded once
// and scattered into the (up to three) output rows it feeds, so the input is read ~(s*R+2)/R times
// from L1/L2 instead of 9.  Consecutive threads = consecutive channel groups (coalesced 16 B).
constexpr int kDwRows = 4;
constexpr int kDwThreads = 128;

__device__ __forceinline__ void dw_load_row(const __nv_bfloat16* base, int h, int H, int W, int cp, int x0,
                                            uint4 (&raw)[3]) {
#pragma unroll
  for (int dx = 0; dx < 3; ++dx) {
    const int x = x0 + dx;
    raw[dx] = make_uint4(0u, 0u, 0u, 0u);
    if (h >= 0 && h < H && x >= 0 && x < W)
      raw[dx] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)h * W + x) * cp));
  }
}

template <int kStride>
__global__ void __launch_bounds__(kDwThreads, 3)
depthwise_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int cp, int Ho, int Wo,
                 const float* __restrict__ w /* [9, cp] */, const float* __restrict__ bias,
                 __nv_bfloat16* __restrict__ out) {
  const int groups = cp >> 3;
  const int strips = (Ho + kDwRows - 1) / kDwRows;
  const int64_t total = (int64_t)B * strips * Wo * groups;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  int64_t r = idx / groups;
  const int wo = (int)(r % Wo);
  r /= Wo;
  const int strip = (int)(r % strips);
  const int b = (int)(r / strips);
  const int c0 = g * 8;
  const int ho0 = strip * kDwRows;
  const int nrows = min(kDwRows, Ho - ho0);

  // packed fp16 arithmetic (HFMA2), like the depthwise stage of the fused blocks: inputs are post-ReLU6
  // (|x| <= 6, bf16 -> fp16 is exact there), 9 taps; half the FMAs and half the weight registers of fp32
  __half2 wt[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + t * cp + c0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + t * cp + c0 + 4));
    wt[t][0] = __floats2half2_rn(w0.x, w0.y);
    wt[t][1] = __floats2half2_rn(w0.z, w0.w);
    wt[t][2] = __floats2half2_rn(w1.x, w1.y);
    wt[t][3] = __floats2half2_rn(w1.z, w1.w);
  }
  __half2 acc[kDwRows][4];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
#pragma unroll
    for (int i = 0; i < kDwRows; ++i) {
      acc[i][0] = __floats2half2_rn(b0.x, b0.y);
      acc[i][1] = __floats2half2_rn(b0.z, b0.w);
      acc[i][2] = __floats2half2_rn(b1.x, b1.y);
      acc[i][3] = __floats2half2_rn(b1.z, b1.w);
    }
  }
  const int x0 = wo * kStride - 1;
  const int h_first = ho0 * kStride - 1;
  constexpr int kInRows = (kDwRows - 1) * kStride + 3;
  const __nv_bfloat16* base = in + ((int64_t)b * H) * W * cp + c0;
  // software pipeline: the loads of input row ir+1 are in flight while row ir is consumed
  uint4 cur[3], nxt[3];
  dw_load_row(base, h_first, H, W, cp, x0, cur);
#pragma unroll
  for (int ir = 0; ir < kInRows; ++ir) {
    if (ir + 1 < kInRows) dw_load_row(base, h_first + ir + 1, H, W, cp, x0, nxt);
    __half2 px[3][4];
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&cur[dx]);
#pragma unroll
      for (int q = 0; q < 4; ++q) px[dx][q] = __float22half2_rn(__bfloat1622float2(h2[q]));
    }
    // input row ir feeds output row o with kernel row dy = ir - o*stride, 0 <= dy < 3
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int num = ir - dy;
      if (num < 0 || (num % kStride) != 0) continue;
      const int o = num / kStride;
      if (o >= kDwRows) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[o][j] = __hfma2(px[dx][j], wt[dy * 3 + dx][j], acc[o][j]);
    }
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) cur[dx] = nxt[dx];
  }
  const __half2 h_zero = __floats2half2_rn(0.f, 0.f), h_six = __floats2half2_rn(6.f, 6.f);
#pragma unroll
  for (int o = 0; o < kDwRows; ++o) {
    if (o < nrows) {
      uint4 pk;
      uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(__hmin2(__hmax2(acc[o][j], h_zero), h_six));
        const __nv_bfloat162 t = __floats2bfloat162_rn(f.x, f.y);
        pw[j] = *reinterpret_cast<const uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(out + ((((int64_t)b * Ho + ho0 + o) * Wo + wo) * cp) + c0) = pk;
    }
  }
}

// Row-sliding depthwise 3x3 for the NARROW late maps (W <= 16: blocks 6-9).  A thread owns 8 channels of a strip
// of 2 output rows and walks the output columns, keeping a 3-column window of the input rows in registers: every
// input pixel is loaded ONCE (the strip kernel above loads it from three threads -- in three different CTAs once
// the channel count exceeds the CTA size -- and those re-reads went to L2).  Packed fp16 arithmetic as above.
constexpr int kDwRowR = 2;

template <int kStride>
__global__ void __launch_bounds__(kDwThreads, 3)
depthwise_row_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int cp, int Ho, int Wo,
                     const float* __restrict__ w /* [9, cp] */, const float* __restrict__ bias,
                     __nv_bfloat16* __restrict__ out) {
  constexpr int IR = (kDwRowR - 1) * kStride + 3;  // input rows feeding the strip
  const int groups = cp >> 3;
  const int strips = (Ho + kDwRowR - 1) / kDwRowR;
  const int64_t total = (int64_t)B * strips * groups;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const int64_t r = idx / groups;
  const int strip = (int)(r % strips);
  const int b = (int)(r / strips);
  const int c0 = g * 8;
  const int ho0 = strip * kDwRowR;
  __half2 wt[9][4], bv[4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + t * cp + c0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + t * cp + c0 + 4));
    wt[t][0] = __floats2half2_rn(w0.x, w0.y);
    wt[t][1] = __floats2half2_rn(w0.z, w0.w);
    wt[t][2] = __floats2half2_rn(w1.x, w1.y);
    wt[t][3] = __floats2half2_rn(w1.z, w1.w);
  }
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    bv[0] = __floats2half2_rn(b0.x, b0.y);
    bv[1] = __floats2half2_rn(b0.z, b0.w);
    bv[2] = __floats2half2_rn(b1.x, b1.y);
    bv[3] = __floats2half2_rn(b1.z, b1.w);
  }
  const int h_first = ho0 * kStride - 1;
  const __nv_bfloat16* base = in + ((int64_t)b * H) * W * cp + c0;
  auto load_raw = [&](int x, uint4 (&raw)[IR]) {
#pragma unroll
    for (int ir = 0; ir < IR; ++ir) {
      const int h = h_first + ir;
      raw[ir] = make_uint4(0u, 0u, 0u, 0u);
      if (h >= 0 && h < H && x >= 0 && x < W) raw[ir] = __ldg(reinterpret_cast<const uint4*>(base + ((int64_t)h * W + x) * cp));
    }
  };
  auto unpack = [&](const uint4 (&raw)[IR], __half2 (&col)[IR][4]) {
#pragma unroll
    for (int ir = 0; ir < IR; ++ir) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[ir]);
#pragma unroll
      for (int q = 0; q < 4; ++q) col[ir][q] = __float22half2_rn(__bfloat1622float2(h2[q]));
    }
  };
  const __half2 h_zero = __floats2half2_rn(0.f, 0.f), h_six = __floats2half2_rn(6.f, 6.f);
  __half2 ca[IR][4], cb[IR][4], cc[IR][4];  // input columns x-1, x, x+1 of the current output column
  uint4 raw[IR], raw2[IR];
  load_raw(-1, raw);
  unpack(raw, ca);
  if (kStride == 1) {
    load_raw(0, raw);
    unpack(raw, cb);
  }
  load_raw(kStride == 1 ? 1 : 0, raw);  // prefetch for the first column
  if (kStride == 2) load_raw(1, raw2);
  for (int wo = 0; wo < Wo; ++wo) {
    if (kStride == 1) {
      unpack(raw, cc);
      if (wo + 1 < Wo) load_raw(wo + 2, raw);  // next column's loads fly under this column's arithmetic
    } else {
      unpack(raw, cb);
      unpack(raw2, cc);
      if (wo + 1 < Wo) {
        load_raw(2 * wo + 2, raw);
        load_raw(2 * wo + 3, raw2);
      }
    }
#pragma unroll
    for (int o = 0; o < kDwRowR; ++o) {
      if (ho0 + o < Ho) {
        __half2 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = bv[j];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int ir = o * kStride + dy;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[j] = __hfma2(ca[ir][j], wt[dy * 3 + 0][j], acc[j]);
            acc[j] = __hfma2(cb[ir][j], wt[dy * 3 + 1][j], acc[j]);
            acc[j] = __hfma2(cc[ir][j], wt[dy * 3 + 2][j], acc[j]);
          }
        }
        uint4 pk;
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(__hmin2(__hmax2(acc[j], h_zero), h_six));
          const __nv_bfloat162 t = __floats2bfloat162_rn(f.x, f.y);
          pw[j] = *reinterpret_cast<const uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(out + ((((int64_t)b * Ho + ho0 + o) * Wo + wo) * cp) + c0) = pk;
      }
    }
    // slide the window
#pragma unroll
    for (int ir = 0; ir < IR; ++ir)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (kStride == 1) {
          ca[ir][j] = cb[ir][j];
          cb[ir][j] = cc[ir][j];
        } else {
          ca[ir][j] = cc[ir][j];
        }
      }
  }
}

// head step 1: mean over the positions a 1x1 stride-s conv visits -> f32 [B, C]
__global__ void strided_mean_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int cp, int C, int stride,
                                    float* __restrict__ out) {
  const int b = blockIdx.x;
  const int hs = (H + stride - 1) / stride, ws = (W + stride - 1) / stride;
  const float inv = 1.0f / (float)(hs * ws);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int h = 0; h < H; h += stride)
      for (int w = 0; w < W; w += stride) acc += __bfloat162float(in[(((int64_t)b * H + h) * W + w) * cp + c]);
    out[(int64_t)b * C + c] = acc * inv;
  }
}

// y[b, n] = bias[n] + sum_k f(x[b, k]) * W[n, k]   (f = identity or exact GELU), fp32.
// Shared-memory tiled SGEMM: kTR (batch rows) x 64 (outputs) per CTA, 16-wide K steps, 256 threads with a
// (kTR / 16) x 4 register micro-tile each.  kTR = 16 keeps the head's small problems (256 rows) on every
// SM: with 64-row tiles linear1 ran on 32 CTAs and took 0.2 ms for 0.3 GFLOP.
template <int kTR>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, int B, int K, const float* __restrict__ W,
                  const float* __restrict__ bias, int N, float* __restrict__ y, int in_act) {
  constexpr int kMR = kTR / 16;
  __shared__ float s_x[16][kTR + 4];  // [k][row]
  __shared__ float s_w[16][64 + 4];   // [k][col]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int row0 = blockIdx.y * kTR, col0 = blockIdx.x * 64;
  float acc[kMR][4];
#pragma unroll
  for (int i = 0; i < kMR; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // k fastest for coalesced global reads (16 consecutive floats per row)
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      const int k = k0 + kk;
      float wv = 0.f;
      if (k < K && col0 + r < N) wv = __ldg(&W[(int64_t)(col0 + r) * K + k]);
      s_w[kk][r] = wv;
      if (r < kTR) {
        float xv = 0.f;
        if (k < K && row0 + r < B) {
          xv = apply_act(in_act, x[(int64_t)(row0 + r) * K + k]);
        }
        s_x[kk][r] = xv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[kMR], bq[4];
#pragma unroll
      for (int i = 0; i < kMR; ++i) a[i] = s_x[kk][ty * kMR + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bq[j] = s_w[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < kMR; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < kMR; ++i) {
    const int r = row0 + ty * kMR + i;
    if (r >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c < N) y[(int64_t)r * N + c] = acc[i][j] + (bias ? bias[c] : 0.f);
    }
  }
}

static int launch_linear(const float* x, int B, int K, const float* W, const float* bias, int N, float* y, int in_act,
                         cudaStream_t st) {
  // 64-row tiles only when they still give every SM two CTAs
  if ((int64_t)ceil_div(N, 64) * ceil_div(B, 64) >= 2 * (int64_t)sm_count()) {
    AM_LAUNCH((linear_f32_kernel<64>), dim3(ceil_div(N, 64), ceil_div(B, 64)), 256, 0, st, x, B, K, W, bias, N, y, in_act);
  } else {
    AM_LAUNCH((linear_f32_kernel<16>), dim3(ceil_div(N, 64), ceil_div(B, 16)), 256, 0, st, x, B, K, W, bias, N, y, in_act);
  }
  return AM_OK;
}

// fp32 -> three bf16 terms so that ONE bf16 tensor-core GEMM over K' = 3*Kp reproduces the fp32 product to ~2^-16:
//   x = hi + lo (+ 2^-17 x),  A' = [hi | hi | lo],  W' = [hi | lo | hi]  =>  A'.W'^T = hi.hi + hi.lo + lo.hi
// (the dropped lo.lo term is 2^-18).  The head's three linears are 0.7 GFLOP in total: on CUDA cores they were
// latency bound at 0.45 ms, as split-bf16 GEMMs on the tcgen05 kernel they are a few microseconds each.
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ x, int B, int K, int Kp, int in_act, __nv_bfloat16* __restrict__ out) {
  const int64_t n = (int64_t)B * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / Kp;
    const int k = (int)(i - b * Kp);
    float v = 0.f;
    if (k < K) {
      v = apply_act(in_act, x[b * K + k]);
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + b * 3 * Kp + k;
    o[0] = hi;
    o[Kp] = hi;
    o[2 * Kp] = lo;
  }
}

// W fp32 [N, K] -> bf16 [N, 3*Kp] = [hi | lo | hi]
static int upload_split3(DevBuf<__nv_bfloat16>& dst, const std::vector<float>& w, int N, int K) {
  const int Kp = (int)round_up((size_t)K, 8);
  std::vector<__nv_bfloat16> h((size_t)N * 3 * Kp, __float2bfloat16_rn(0.f));
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float v = w[(size_t)n * K + k];
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      __nv_bfloat16* o = h.data() + (size_t)n * 3 * Kp + k;
      o[0] = hi;
      o[Kp] = lo;
      o[2 * Kp] = hi;
    }
  AM_TRY(dst.alloc(h.size()));
  AM_CUDA(cudaMemcpy(dst.p, h.data(), h.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  return AM_OK;
}

// head final: z = LayerNorm(e1 + e2) * g + b ; out = z / max(||z||, eps2)   (one CTA per row)
__global__ void __launch_bounds__(256)
head_finalize_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int E, const float* __restrict__ g,
                     const float* __restrict__ bt, float eps, float eps2, float* __restrict__ out) {
  __shared__ float s_red[32];
  __shared__ float s_stat[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* a = e1 + (int64_t)b * E;
  const float* c = e2 + (int64_t)b * E;
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
      t = lane < (blockDim.x >> 5) ? s_red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) s_stat[0] = t;
    }
    __syncthreads();
    const float r = s_stat[0];
    __syncthreads();
    return r;
  };
  float loc = 0.f;
  for (int i = tid; i < E; i += blockDim.x) loc += a[i] + c[i];
  const float mean = block_sum(loc) / (float)E;
  loc = 0.f;
  for (int i = tid; i < E; i += blockDim.x) {
    const float d = a[i] + c[i] - mean;
    loc = fmaf(d, d, loc);
  }
  const float var = block_sum(loc) / (float)E;
  const float rstd = rsqrtf(var + eps);
  loc = 0.f;
  for (int i = tid; i < E; i += blockDim.x) {
    const float z = (a[i] + c[i] - mean) * rstd * g[i] + bt[i];
    loc = fmaf(z, z, loc);
  }
  const float nrm = fmaxf(sqrtf(block_sum(loc)), eps2);
  for (int i = tid; i < E; i += blockDim.x) {
    const float z = (a[i] + c[i] - mean) * rstd * g[i] + bt[i];
    out[(int64_t)b * E + i] = z / nrm;
  }
}

// K3: per-track mean of window embeddings, then / (||.|| + 1e-9)   (clap_analyzer.py:552-562)
__global__ void __launch_bounds__(256)
track_pool_kernel(const float* __restrict__ seg_emb, const int32_t* __restrict__ seg_off, int E,
                  float* __restrict__ out) {
  __shared__ float s_red[8];
  __shared__ float s_norm;
  const int t = blockIdx.x, tid = threadIdx.x;
  const int s0 = seg_off[t], s1 = seg_off[t + 1];
  const int n = s1 - s0;
  float loc = 0.f;
  for (int i = tid; i < E; i += blockDim.x) {
    float acc = 0.f;
    for (int s = s0; s < s1; ++s) acc += seg_emb[(int64_t)s * E + i];
    const float m = n > 0 ? acc / (float)n : 0.f;
    out[(int64_t)t * E + i] = m;
    loc = fmaf(m, m, loc);
  }
  loc = warp_sum(loc);
  if ((tid & 31) == 0) s_red[tid >> 5] = loc;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
    s_norm = sqrtf(tot) + 1e-9f;
  }
  __syncthreads();
  if (n > 0)
    for (int i = tid; i < E; i += blockDim.x) out[(int64_t)t * E + i] /= s_norm;
}

// ---------------------------------------------------------------- blob reader
struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  template <typename T>
  T get() {
    T v{};
    if (p + sizeof(T) > end) {
      ok = false;
      return v;
    }
    std::memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  std::vector<float> floats() {
    const uint64_t n = get<uint64_t>();
    std::vector<float> v;
    if (!ok || n > (uint64_t)(end - p) / 4) {
      ok = false;
      return v;
    }
    v.resize(n);
    std::memcpy(v.data(), p, n * 4);
    p += n * 4;
    return v;
  }
};

template <typename T>
static int upload(DevBuf<T>& dst, const std::vector<T>& src) {
  AM_TRY(dst.alloc(std::max<size_t>(src.size(), 1)));
  if (!src.empty()) AM_CUDA(cudaMemcpy(dst.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return AM_OK;
}

static std::vector<float> padded(const std::vector<float>& v, size_t n) {
  std::vector<float> o(n, 0.f);
  std::copy(v.begin(), v.begin() + std::min(n, v.size()), o.begin());
  return o;
}

}  // namespace am

struct am_model {
  int n_mels = 0, emb = 0;
  std::string source;
  std::vector<std::unique_ptr<am::Layer>> layers;
  std::vector<std::unique_ptr<am::HeadOp>> head;  // row program: pool, linears, ..., the last op writes the embedding
  std::vector<int> reg_dim;
  std::vector<std::unique_ptr<am::DevBuf<float>>> regs;  // head registers f32 [n, reg_dim]
  int head_cin = 0, head_cin_p = 0;                      // channels the trunk hands to the head's pooling
  // workspace (grown on demand, reused across calls; one user thread per model)
  am::DevBuf<__nv_bfloat16> act[3];
  am::DevBuf<float> mel_ws, seg_emb, se_mean, se_gate;
  am::DevBuf<__nv_bfloat16> a3;       // head GEMM operand [n, 3 * Kp] (split3_kernel)
  am::DevBuf<__nv_bfloat16> late_in;  // [n, H, W, C] output of the early (fused) blocks for all windows of a call
  int late_sub = 256;                 // windows per pass of the late phase
  size_t act_elems = 0;
  int max_sub = 256;         // windows per pass of the early (chunked) phase (256 measured 1 % faster than 128)
  am::Stream stream;         // compute stream of the host-pointer entry points
  am::Stream copy_stream;    // H2D staging stream (host API): copies of chunk i+1 overlap compute of chunk i
  am::DevBuf<int16_t> pcm_stage[2];
  am::DevBuf<int32_t> off_stage;
  am::DevBuf<float> out_stage;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  // submitted-but-not-collected host calls (am_clap_embed_tracks_submit / _collect): results land in pinned
  // staging first, so the D2H is truly asynchronous and the NEXT call's H2D + early blocks overlap this call's tail
  struct Ticket {
    am::PinnedBuf<float> stage;
    float* user_out = nullptr;
    size_t count = 0;
    cudaEvent_t ready = nullptr;
    bool open = false;
  } tickets[2];
  unsigned n_submitted = 0, n_collected = 0;
  bool slot_used[2] = {false, false};
  am_mel_plan* host_plan = nullptr;  // mel plan of the host entry point, cached per cfg (building one = host trig
  am_mel_cfg host_plan_cfg{};        // tables + cudaMalloc + upload: ~1 ms, was paid on every call)
  int use_simt_gemm = 0;     // debug: AM_GEMM_IMPL=simt
  struct Block {             // inverted-residual block = [expand] depthwise project
    int first = 0, expand = -1, dw = 0, proj = 0;
  };
  std::vector<Block> blocks;
  unsigned fused_mask = 0xffffffffu;   // AM_FUSED_BLOCKS: bit i = fuse block i when it fits on chip
  ~am_model() {
    for (int i = 0; i < 2; ++i) {
      if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
    }
    if (host_plan) am_mel_plan_free(host_plan);
    for (auto& t : tickets)
      if (t.ready) cudaEventDestroy(t.ready);
  }
};

namespace am {

struct Shape {
  int H, W;
};

// output extent of a spatial layer (first conv / stem / depthwise); 1x1 layers keep the shape
static Shape layer_out(const Layer& l, Shape s) {
  if (l.type == kDepthwise || l.type == kConvFirst || l.type == kStem)
    return {(s.H + l.pad_t + l.pad_b - l.kh) / l.stride + 1, (s.W + l.pad_l + l.pad_r - l.kw) / l.stride + 1};
  return s;
}
static Shape stem_out(const Layer& l, int T, int n_mels) {
  return layer_out(l, l.h_is_time ? Shape{T, n_mels} : Shape{n_mels, T});
}
static Shape dw_out(const Layer& l, Shape s) { return layer_out(l, s); }

// AMW1 blob (audiomuse-ai_b200/weights.py) -> ModelSpec
static int parse_blob(const void* blob, size_t nbytes, ModelSpec* spec) {
  Reader r{(const uint8_t*)blob, (const uint8_t*)blob + nbytes};
  char magic[4];
  for (int i = 0; i < 4; ++i) magic[i] = (char)r.get<uint8_t>();
  if (!r.ok || std::memcmp(magic, "AMW1", 4) != 0) {
    set_error("weights: bad magic (expected AMW1 or an ONNX ModelProto)");
    return AM_ERR_IO;
  }
  const uint32_t version = r.get<uint32_t>();
  spec->n_mels = (int)r.get<uint32_t>();
  spec->emb = (int)r.get<uint32_t>();
  spec->source = "AMW1";
  const uint32_t n_layers = r.get<uint32_t>();
  if (!r.ok || version != 1 || n_layers == 0 || n_layers > 4096) {
    set_error("weights: bad header (version %u, %u layers)", version, n_layers);
    return AM_ERR_IO;
  }
  for (uint32_t li = 0; li < n_layers; ++li) {
    const uint32_t type = r.get<uint32_t>();
    int32_t prm[8];
    for (int i = 0; i < 8; ++i) prm[i] = r.get<int32_t>();
    if (!r.ok) break;
    if (type == kHead) {  // pn_block (1x1 stride s) -> mean -> Projection -> LayerNorm -> L2 as a row program
      const int cin = prm[0], trunk = prm[1], emb = prm[2], stride = prm[3];
      float ln_eps;
      std::memcpy(&ln_eps, &prm[4], 4);
      auto pn_w = r.floats(), pn_b = r.floats(), l1 = r.floats(), l2 = r.floats(), g = r.floats(), b = r.floats();
      if (!r.ok || pn_w.size() != (size_t)trunk * cin || pn_b.size() != (size_t)trunk || l1.size() != (size_t)emb * trunk ||
          l2.size() != (size_t)emb * emb || g.size() != (size_t)emb || b.size() != (size_t)emb) {
        set_error("weights: malformed head record");
        return AM_ERR_IO;
      }
      auto reg = [&](int dim) {
        spec->reg_dim.push_back(dim);
        return spec->n_regs++;
      };
      VecOp pool;
      pool.kind = kVecPool;
      pool.stride = stride;
      pool.N = cin;
      pool.dst = reg(cin);
      VecOp pn;
      pn.kind = kVecLinear;
      pn.a = pool.dst;
      pn.K = cin;
      pn.N = trunk;
      pn.w = pn_w;
      pn.bias = pn_b;
      pn.dst = reg(trunk);
      VecOp e1;
      e1.kind = kVecLinear;
      e1.a = pn.dst;
      e1.K = trunk;
      e1.N = emb;
      e1.w = l1;
      e1.dst = reg(emb);
      VecOp e2;
      e2.kind = kVecLinear;
      e2.a = e1.dst;
      e2.K = emb;
      e2.N = emb;
      e2.act = kActGelu;
      e2.w = l2;
      e2.dst = reg(emb);
      VecOp fin;
      fin.kind = kVecAddLnL2;
      fin.a = e1.dst;
      fin.b = e2.dst;
      fin.N = emb;
      fin.eps = ln_eps;
      fin.eps2 = 1e-12f;
      fin.w = g;
      fin.bias = b;
      fin.dst = reg(emb);
      spec->head = {pool, pn, e1, e2, fin};
      continue;
    }
    LayerSpec L;
    L.type = (int)type;
    if (type == kStem) {
      L.cin = 1;
      L.cout = prm[0];
      L.kh = L.kw = 3;
      L.stride = 2;
      L.act = kActRelu6;
      L.pad_t = prm[1];
      L.pad_b = prm[2];
      L.pad_l = prm[3];
      L.pad_r = prm[4];
      L.aux0 = r.floats();
      L.aux1 = r.floats();
      L.w = r.floats();
      L.aux2 = r.floats();
      L.bias = r.floats();
      if (!r.ok || L.aux0.size() != (size_t)spec->n_mels || L.aux1.size() != (size_t)spec->n_mels || L.w.size() != 9 ||
          L.aux2.size() != (size_t)L.cout || L.bias.size() != (size_t)L.cout) {
        set_error("weights: malformed stem record");
        return AM_ERR_IO;
      }
    } else if (type == kPointwise) {
      L.cin = prm[0];
      L.cout = prm[1];
      L.act = prm[2];
      L.residual = prm[3];
      L.block_start = prm[4];
      L.w = r.floats();
      L.bias = r.floats();
      if (!r.ok || L.w.size() != (size_t)L.cin * L.cout || L.bias.size() != (size_t)L.cout) {
        set_error("weights: malformed pointwise record (layer %u)", li);
        return AM_ERR_IO;
      }
    } else if (type == kDepthwise) {
      L.cin = L.cout = prm[0];
      L.stride = prm[1];
      L.block_start = prm[4];
      L.kh = L.kw = 3;
      L.pad_t = L.pad_b = L.pad_l = L.pad_r = 1;
      L.act = kActRelu6;
      L.w = r.floats();
      L.bias = r.floats();  // w: [c, 9]
      if (!r.ok || L.w.size() != (size_t)L.cin * 9 || L.bias.size() != (size_t)L.cin || (L.stride != 1 && L.stride != 2)) {
        set_error("weights: malformed depthwise record (layer %u)", li);
        return AM_ERR_IO;
      }
    } else {
      set_error("weights: unknown layer type %u", type);
      return AM_ERR_IO;
    }
    spec->layers.push_back(std::move(L));
  }
  if (!r.ok) {
    set_error("weights: truncated blob");
    return AM_ERR_IO;
  }
  return AM_OK;
}

// ModelSpec -> device model: pads channels to 16, converts weights to the kernels' layouts, validates the chain
static int build_model(am_model* m, const ModelSpec& spec) {
  m->n_mels = spec.n_mels;
  m->emb = spec.emb;
  m->source = spec.source;
  if (spec.layers.empty() || (spec.layers[0].type != kStem && spec.layers[0].type != kConvFirst) || spec.head.empty() ||
      spec.emb <= 0 || spec.n_mels <= 0) {
    set_error("weights: model must start with a convolution on the mel spectrogram and end with a head");
    return AM_ERR_IO;
  }
  int c = 0;
  for (size_t li = 0; li < spec.layers.size(); ++li) {
    const LayerSpec& S = spec.layers[li];
    auto L = std::make_unique<Layer>();
    L->type = S.type;
    L->cin = S.cin;
    L->cout = S.cout;
    L->cin_p = pad16(S.cin);
    L->cout_p = pad16(S.cout);
    L->kh = S.kh;
    L->kw = S.kw;
    L->stride = S.stride;
    L->act = S.act;
    L->block_start = S.block_start;
    L->residual = S.residual;
    L->pad_t = S.pad_t;
    L->pad_b = S.pad_b;
    L->pad_l = S.pad_l;
    L->pad_r = S.pad_r;
    L->h_is_time = S.h_is_time;
    L->gate_act = S.gate_act;
    L->cmid = S.cmid;
    if (li > 0 && S.cin != c) {
      set_error("weights: layer %zu expects %d input channels, previous layer produces %d", li, S.cin, c);
      return AM_ERR_IO;
    }
    if (li > 0 && (S.type == kStem || S.type == kConvFirst)) {
      set_error("weights: layer %zu: a first convolution in the middle of the trunk", li);
      return AM_ERR_IO;
    }
    if (S.type == kStem) {
      AM_TRY(upload(L->aux0, S.aux0));
      AM_TRY(upload(L->aux1, S.aux1));
      AM_TRY(upload(L->w_f32, S.w));
      AM_TRY(upload(L->aux2, padded(S.aux2, L->cout_p)));
      AM_TRY(upload(L->bias, padded(S.bias, L->cout_p)));
    } else if (S.type == kConvFirst) {
      const int taps = S.kh * S.kw;
      if (S.w.size() != (size_t)S.cout * taps || S.bias.size() != (size_t)S.cout ||
          (!S.aux0.empty() && (S.aux0.size() != (size_t)spec.n_mels || S.aux1.size() != (size_t)spec.n_mels))) {
        set_error("weights: malformed first convolution");
        return AM_ERR_IO;
      }
      std::vector<float> wt((size_t)taps * L->cout_p, 0.f);
      for (int o = 0; o < S.cout; ++o)
        for (int t = 0; t < taps; ++t) wt[(size_t)t * L->cout_p + o] = S.w[(size_t)o * taps + t];
      AM_TRY(upload(L->w_f32, wt));
      AM_TRY(upload(L->bias, padded(S.bias, L->cout_p)));
      if (!S.aux0.empty()) {
        AM_TRY(upload(L->aux0, S.aux0));
        AM_TRY(upload(L->aux1, S.aux1));
      }
    } else if (S.type == kPointwise) {
      if (S.w.size() != (size_t)S.cin * S.cout || S.bias.size() != (size_t)S.cout) {
        set_error("weights: malformed pointwise layer %zu", li);
        return AM_ERR_IO;
      }
      std::vector<__nv_bfloat16> wb((size_t)L->cout_p * L->cin_p, __float2bfloat16_rn(0.f));
      for (int o = 0; o < S.cout; ++o)
        for (int i = 0; i < S.cin; ++i) wb[(size_t)o * L->cin_p + i] = __float2bfloat16_rn(S.w[(size_t)o * S.cin + i]);
      AM_TRY(upload(L->w_bf16, wb));
      if (S.act == kActNone) {  // fp16 copy, saturated to the fp16 range (folded weights are O(1))
        std::vector<__half> wh((size_t)L->cout_p * L->cin_p, __float2half_rn(0.f));
        for (int o = 0; o < S.cout; ++o)
          for (int i = 0; i < S.cin; ++i)
            wh[(size_t)o * L->cin_p + i] = __float2half_rn(std::min(65504.f, std::max(-65504.f, S.w[(size_t)o * S.cin + i])));
        AM_TRY(upload(L->w_f16, wh));
      }
      AM_TRY(upload(L->bias, padded(S.bias, L->cout_p)));
    } else if (S.type == kDepthwise) {
      const int taps = S.kh * S.kw;
      if (S.w.size() != (size_t)S.cin * taps || S.bias.size() != (size_t)S.cin) {
        set_error("weights: malformed depthwise layer %zu", li);
        return AM_ERR_IO;
      }
      std::vector<float> wt((size_t)taps * L->cin_p, 0.f);
      for (int ch = 0; ch < S.cin; ++ch)
        for (int t = 0; t < taps; ++t) wt[(size_t)t * L->cin_p + ch] = S.w[(size_t)ch * taps + t];
      AM_TRY(upload(L->w_f32, wt));
      AM_TRY(upload(L->bias, padded(S.bias, L->cin_p)));
    } else if (S.type == kSqueezeExcite) {
      if (S.w.size() != (size_t)S.cmid * S.cin || S.bias.size() != (size_t)S.cmid || S.aux0.size() != (size_t)S.cin * S.cmid ||
          S.aux1.size() != (size_t)S.cin) {
        set_error("weights: malformed squeeze-excite layer %zu", li);
        return AM_ERR_IO;
      }
      AM_TRY(upload(L->w_f32, S.w));
      AM_TRY(upload(L->bias, S.bias));
      AM_TRY(upload(L->aux0, S.aux0));
      AM_TRY(upload(L->aux1, S.aux1));
    } else {
      set_error("weights: unknown layer type %d", S.type);
      return AM_ERR_IO;
    }
    c = S.cout;
    m->layers.push_back(std::move(L));
  }
  // ---- head program
  m->reg_dim = spec.reg_dim;
  for (int rdim : spec.reg_dim) {
    (void)rdim;
    m->regs.push_back(std::make_unique<DevBuf<float>>());
  }
  for (size_t q = 0; q < spec.head.size(); ++q) {
    const VecOp& S = spec.head[q];
    auto H = std::make_unique<HeadOp>();
    H->kind = S.kind;
    H->a = S.a;
    H->b = S.b;
    H->dst = S.dst;
    H->K = S.K;
    H->N = S.N;
    H->act = S.act;
    H->stride = S.stride;
    H->eps = S.eps;
    H->eps2 = S.eps2;
    auto reg_ok = [&](int r) { return r >= 0 && r < spec.n_regs; };
    if (!reg_ok(S.dst) || (S.kind != kVecPool && !reg_ok(S.a)) || ((S.kind == kVecAdd || S.kind == kVecAddLnL2) && !reg_ok(S.b))) {
      set_error("weights: head op %zu references an unknown register", q);
      return AM_ERR_IO;
    }
    if (S.kind == kVecPool) {
      if (q != 0 || S.N != c) {
        set_error("weights: the head must start by pooling the trunk's %d channels", c);
        return AM_ERR_IO;
      }
      m->head_cin = c;
      m->head_cin_p = pad16(c);
    } else if (S.kind == kVecLinear) {
      if (S.w.size() != (size_t)S.N * S.K || (!S.bias.empty() && S.bias.size() != (size_t)S.N) || spec.reg_dim[(size_t)S.a] != S.K ||
          spec.reg_dim[(size_t)S.dst] != S.N) {
        set_error("weights: malformed linear in the head (op %zu)", q);
        return AM_ERR_IO;
      }
      AM_TRY(upload(H->w, S.w));
      AM_TRY(upload_split3(H->w3, S.w, S.N, S.K));
      H->has_w = true;
      if (!S.bias.empty()) {
        AM_TRY(upload(H->bias, S.bias));
        H->has_bias = true;
      }
    } else if (S.kind == kVecAffine || S.kind == kVecLayerNorm || S.kind == kVecAddLnL2) {
      const size_t dim = (size_t)spec.reg_dim[(size_t)S.dst];
      if ((!S.w.empty() && S.w.size() != dim) || (!S.bias.empty() && S.bias.size() != dim) ||
          (S.kind != kVecAffine && (S.w.empty() || S.bias.empty()))) {
        set_error("weights: malformed row op %zu in the head", q);
        return AM_ERR_IO;
      }
      if (!S.w.empty()) {
        AM_TRY(upload(H->w, S.w));
        H->has_w = true;
      }
      if (!S.bias.empty()) {
        AM_TRY(upload(H->bias, S.bias));
        H->has_bias = true;
      }
    }
    m->head.push_back(std::move(H));
  }
  if (m->head.empty() || m->head[0]->kind != kVecPool || spec.reg_dim[(size_t)m->head.back()->dst] != m->emb) {
    set_error("weights: the head must start with the spatial pooling and end with the %d-d embedding", m->emb);
    return AM_ERR_IO;
  }
  // group layers into inverted-residual blocks for the fused kernel: [1x1 expand + ReLU6] -> 3x3 dw + ReLU6 -> linear 1x1
  for (size_t i = 1; i < m->layers.size();) {
    const Layer& l = *m->layers[i];
    am_model::Block blk;
    blk.first = (int)i;
    if (l.type == kPointwise && l.block_start && l.act == kActRelu6 && i + 2 < m->layers.size() + 0 &&
        m->layers[i + 1]->dw_fast() && m->layers[i + 2]->type == kPointwise && m->layers[i + 2]->act == kActNone) {
      blk.expand = (int)i;
      blk.dw = (int)i + 1;
      blk.proj = (int)i + 2;
      {  // constants of the channel-per-lane fused kernel (fused_block_t.cu), packed once
        Layer& dwl = *m->layers[i + 1];
        AM_TRY(dwl.tpack.alloc(fusedt::consts_words(dwl.cin_p)));
        AM_TRY(fusedt::pack_consts(dwl.w_f32.p, dwl.bias.p, l.bias.p, dwl.cin_p, dwl.tpack.p, nullptr));
        AM_CUDA(cudaStreamSynchronize(nullptr));
      }
      m->blocks.push_back(blk);
      i += 3;
    } else if (l.dw_fast() && l.block_start && i + 1 < m->layers.size() && m->layers[i + 1]->type == kPointwise &&
               m->layers[i + 1]->act == kActNone) {
      blk.dw = (int)i;
      blk.proj = (int)i + 1;
      if (i == 1 && m->layers[0]->type == kStem) {
        // block 0 behind the rank-1 stem: the stem's per-channel scale / shift become expansion weights over the
        // (v_hi, v_lo, v_hi, 1, 1) rows stem_x0_kernel writes, and the block runs on the channel-per-lane kernel
        Layer& stem = *m->layers[0];
        Layer& dwl = *m->layers[i];
        const auto& S0 = spec.layers[0];
        std::vector<__nv_bfloat16> w1((size_t)stem.cout_p * 16, __float2bfloat16_rn(0.f));
        for (int c = 0; c < S0.cout; ++c) {
          const float sc = S0.aux2[(size_t)c], sh = S0.bias[(size_t)c];
          const __nv_bfloat16 sch = __float2bfloat16_rn(sc), shh = __float2bfloat16_rn(sh);
          __nv_bfloat16* r = &w1[(size_t)c * 16];
          r[0] = sch;
          r[1] = sch;
          r[2] = __float2bfloat16_rn(sc - __bfloat162float(sch));
          r[3] = shh;
          r[4] = __float2bfloat16_rn(sh - __bfloat162float(shh));
        }
        AM_TRY(upload(stem.w1s, w1));
        DevBuf<float> zeros;
        AM_TRY(zeros.alloc((size_t)dwl.cin_p));
        AM_CUDA(cudaMemset(zeros.p, 0, (size_t)dwl.cin_p * 4));
        AM_TRY(dwl.tpack.alloc(fusedt::consts_words(dwl.cin_p)));
        AM_TRY(fusedt::pack_consts(dwl.w_f32.p, dwl.bias.p, zeros.p, dwl.cin_p, dwl.tpack.p, nullptr));
        AM_CUDA(cudaStreamSynchronize(nullptr));
      }
      m->blocks.push_back(blk);
      i += 2;
    } else {
      ++i;
    }
  }
  return AM_OK;
}

static int grid_for(int64_t total_threads) {
  const int64_t blocks = (total_threads + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sm_count() * 16));
}

// fused-kernel description of block `bi` at input shape `s` (plan() decides whether it fits on chip)
static bool block_desc(const am_model* m, int bi, Shape s, bool stem_fp16, fused::BlockDesc* d, fused::Plan* pl) {
  if (m->use_simt_gemm || !((m->fused_mask >> bi) & 1u)) return false;
  const am_model::Block& blk = m->blocks[bi];
  const Layer& dwl = *m->layers[blk.dw];
  const Layer& pj = *m->layers[blk.proj];
  *d = fused::BlockDesc{};
  d->H = s.H;
  d->W = s.W;
  d->cin_p = blk.expand >= 0 ? m->layers[blk.expand]->cin_p : dwl.cin_p;
  d->cmid_p = dwl.cin_p;
  d->cout_p = pj.cout_p;
  d->stride = dwl.stride;
  d->has_expand = blk.expand >= 0 ? 1 : 0;
  d->residual = pj.residual;
  d->x_is_fp16 = (bi == 0 && stem_fp16) ? 1 : 0;
  return fused::plan(*d, pl);
}

static int block_index_at(const am_model* m, size_t layer) {
  for (size_t q = 0; q < m->blocks.size(); ++q)
    if (m->blocks[q].first == (int)layer) return (int)q;
  return -1;
}

// The trunk runs in two phases.  EARLY = stem + the leading run of blocks that execute fused (huge
// spatial extent): processed chunk by chunk, so host->device copies of the next chunk hide under it.
// LATE = everything after (small tensors): processed ONCE for all windows of the call, so the deep,
// narrow layers get full-size grids.  late_start() is the first layer of the late phase.
static size_t late_start(const am_model* m, int T, Shape* s_split, int* c_split) {
  Shape s = stem_out(*m->layers[0], T, m->n_mels);
  int c = m->layers[0]->cout_p;
  size_t i = 1;
  while (i < m->layers.size()) {
    const int bi = block_index_at(m, i);
    if (bi < 0) break;
    fused::BlockDesc d;
    fused::Plan pl;
    // x_is_fp16 does not influence plan(); pass false
    if (!block_desc(m, bi, s, false, &d, &pl)) break;
    s = dw_out(*m->layers[m->blocks[bi].dw], s);
    c = m->layers[m->blocks[bi].proj]->cout_p;
    i = (size_t)m->blocks[bi].proj + 1;
  }
  *s_split = s;
  *c_split = c;
  return i;
}

// Runs layers [lo, hi) on `nb` windows.  lo == 0: starts from the log-mel (stem); else from `in` (shape s_in).
// The last operation writes to `final_out` when given, else to a workspace buffer; *out / *s_out describe it.
static int run_range(am_model* m, const float* mel_dev, const __nv_bfloat16* in, Shape s_in, size_t lo, size_t hi,
                     int nb, int T, __nv_bfloat16* final_out, const __nv_bfloat16** out, Shape* s_out,
                     cudaStream_t st) {
  Shape s = s_in;
  const __nv_bfloat16* cur = in;
  const __nv_bfloat16* block_in = in;
  auto pick_dst = [&](bool is_last) -> __nv_bfloat16* {
    if (is_last && final_out) return final_out;
    for (auto& b : m->act)
      if (b.p != cur && b.p != block_in) return b.p;
    return nullptr;
  };
  size_t i = lo;
  bool stem_fp16 = false;
  bool stem_x0 = false;        // the stem was written as its rank-1 factor: block 0 runs it as an expansion GEMM
  fused::BlockDesc d0{};
  fusedt::Plan plt0;
  if (lo == 0) {
    const Layer& stem = *m->layers[0];
    s = stem_out(stem, T, m->n_mels);
    AM_CHECK(s.H > 0 && s.W > 0, "encoder: input of %d frames x %d mels is too small", T, m->n_mels);
    __nv_bfloat16* dst = pick_dst(hi == 1);
    if (stem.type == kStem) {
      // the stem feeds only the first block; when that block runs fused its depthwise wants fp16 input
      if (!m->blocks.empty() && m->blocks[0].first == 1 && m->blocks[0].expand < 0 && hi > 1) {
        fused::BlockDesc d;
        fused::Plan pl;
        stem_fp16 = block_desc(m, 0, s, false, &d, &pl);
        // Opt-in (AM_STEM_X0=1): measured on B200 the stem-as-expansion-GEMM route is slower for the shipped student
        // (144 channels = one full 128-lane chunk + a 16-channel one that costs as much: 11.0 k cycles per tile against
        // 7.8 k for the pixel-per-lane kernel + the stem kernel's 0.5 ms per 256 windows)
        static const bool use_x0 = std::getenv("AM_STEM_X0") != nullptr;
        const am_model::Block& b0 = m->blocks[0];
        if (stem_fp16 && use_x0 && stem.w1s.p && !m->layers[b0.proj]->residual && (size_t)b0.proj < hi) {
          d0 = d;
          d0.cin_p = 16;
          d0.has_expand = 1;
          d0.x_is_fp16 = 0;
          stem_x0 = fusedt::plan(d0, &plt0);
        }
      }
      const int64_t n_tiles = (int64_t)nb * ((s.H + kStemTH - 1) / kStemTH) * ((s.W + kStemTW - 1) / kStemTW);
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)sm_count() * 8));
      if (stem_x0) {
        AM_LAUNCH(stem_x0_kernel, grid, 256, 0, st, mel_dev, nb, m->n_mels, T, s.H, s.W, stem.pad_t, stem.pad_l, stem.aux0.p,
                  stem.aux1.p, stem.w_f32.p, dst);
      } else if (stem_fp16) {
        AM_LAUNCH(stem_kernel<true>, grid, 256, (size_t)stem.cout_p * 8, st, mel_dev, nb, m->n_mels, T, s.H, s.W,
                  stem.pad_t, stem.pad_l, stem.aux0.p, stem.aux1.p, stem.w_f32.p, stem.aux2.p, stem.bias.p, stem.cout_p,
                  dst);
      } else {
        AM_LAUNCH(stem_kernel<false>, grid, 256, (size_t)stem.cout_p * 8, st, mel_dev, nb, m->n_mels, T, s.H, s.W,
                  stem.pad_t, stem.pad_l, stem.aux0.p, stem.aux1.p, stem.w_f32.p, stem.aux2.p, stem.bias.p, stem.cout_p,
                  dst);
      }
    } else {
      const size_t smem = (size_t)(stem.kh * stem.kw + 1) * stem.cout_p * sizeof(float);
      AM_CHECK(smem <= 48 * 1024, "encoder: first convolution with %d x %d taps x %d channels does not fit in shared memory",
               stem.kh, stem.kw, stem.cout_p);
      const int grid = grid_for((int64_t)nb * s.H * s.W * (stem.cout_p / 8));
      AM_LAUNCH(conv_first_kernel, grid, 256, smem, st, mel_dev, nb, m->n_mels, T, s.H, s.W, stem.kh, stem.kw, stem.stride,
                stem.pad_t, stem.pad_l, stem.h_is_time, stem.aux0.p, stem.aux1.p, stem.w_f32.p, stem.bias.p, stem.cout_p,
                stem.act, dst);
    }
    cur = block_in = dst;
    i = 1;
  }
  for (; i < hi; ++i) {
    const Layer& l = *m->layers[i];
    if (l.block_start) {
      // whole block in one kernel when it fits on chip (fused_block.cu)
      const int bi = block_index_at(m, i);
      fused::BlockDesc d;
      fused::Plan pl;
      if (bi >= 0 && (size_t)m->blocks[bi].proj < hi && block_desc(m, bi, s, stem_fp16, &d, &pl)) {
        const am_model::Block& blk = m->blocks[bi];
        const Layer& dwl = *m->layers[blk.dw];
        const Layer& pj = *m->layers[blk.proj];
        const Layer* ex = blk.expand >= 0 ? m->layers[blk.expand].get() : nullptr;
        block_in = cur;
        __nv_bfloat16* dst = pick_dst((size_t)blk.proj + 1 == hi);
        fusedt::Plan plt;
        if (bi == 0 && stem_x0) {            // stem + block 0: the stem's affine is this block's expansion GEMM
          AM_TRY(fusedt::run(d0, plt0, cur, m->layers[0]->w1s.p, dwl.tpack.p, pj.w_f16.p, pj.bias.p, dst, nb, st));
        } else if (ex && fusedt::plan(d, &plt)) {   // channel-per-lane kernel: depthwise taps straight from TMEM
          AM_TRY(fusedt::run(d, plt, cur, ex->w_bf16.p, dwl.tpack.p, pj.w_f16.p, pj.bias.p, dst, nb, st));
        } else {
          AM_TRY(fused::run(d, pl, cur, ex ? ex->w_bf16.p : nullptr, ex ? ex->bias.p : nullptr, dwl.w_f32.p,
                            dwl.bias.p, pj.w_f16.p, pj.bias.p, dst, nb, st));
        }
        s = dw_out(dwl, s);
        cur = block_in = dst;
        i = (size_t)blk.proj;
        continue;
      }
      block_in = cur;
    }
    __nv_bfloat16* dst = pick_dst(i + 1 == hi);
    if (l.type == kDepthwise && !l.dw_fast()) {
      const Shape o = dw_out(l, s);
      AM_CHECK(o.H > 0 && o.W > 0, "encoder: depthwise layer %zu has an empty output", i);
      const int grid = grid_for((int64_t)nb * o.H * o.W * (l.cout_p / 8));
      AM_LAUNCH(depthwise_generic_kernel, grid, 256, 0, st, cur, nb, s.H, s.W, l.cin_p, o.H, o.W, l.kh, l.stride, l.pad_t,
                l.pad_l, l.w_f32.p, l.bias.p, l.act, dst);
      s = o;
    } else if (l.type == kDepthwise) {
      const Shape o = dw_out(l, s);
      const int strips = (o.H + kDwRows - 1) / kDwRows;
      const int64_t total = (int64_t)nb * strips * o.W * (l.cout_p / 8);
      const unsigned grid = (unsigned)((total + kDwThreads - 1) / kDwThreads);
      static const bool no_row = std::getenv("AM_DW_NO_ROW") != nullptr;
      const int rstrips = (o.H + kDwRowR - 1) / kDwRowR;
      const int64_t rtotal = (int64_t)nb * rstrips * (l.cout_p / 8);
      if (!no_row && s.W <= 16 && rtotal >= (int64_t)sm_count() * 8 * kDwThreads) {
        // narrow late maps with enough (window, strip, channel group) work items to fill the GPU
        const unsigned rgrid = (unsigned)((rtotal + kDwThreads - 1) / kDwThreads);
        if (l.stride == 1) {
          AM_LAUNCH(depthwise_row_kernel<1>, rgrid, kDwThreads, 0, st, cur, nb, s.H, s.W, l.cin_p, o.H, o.W, l.w_f32.p,
                    l.bias.p, dst);
        } else {
          AM_LAUNCH(depthwise_row_kernel<2>, rgrid, kDwThreads, 0, st, cur, nb, s.H, s.W, l.cin_p, o.H, o.W, l.w_f32.p,
                    l.bias.p, dst);
        }
      } else if (l.stride == 1) {
        AM_LAUNCH(depthwise_kernel<1>, grid, kDwThreads, 0, st, cur, nb, s.H, s.W, l.cin_p, o.H, o.W, l.w_f32.p,
                  l.bias.p, dst);
      } else {
        AM_LAUNCH(depthwise_kernel<2>, grid, kDwThreads, 0, st, cur, nb, s.H, s.W, l.cin_p, o.H, o.W, l.w_f32.p,
                  l.bias.p, dst);
      }
      s = o;
    } else if (l.type == kSqueezeExcite) {
      const int HW = s.H * s.W;
      AM_TRY(m->se_mean.ensure((size_t)nb * l.cin));
      AM_TRY(m->se_gate.ensure((size_t)nb * l.cin));
      AM_LAUNCH(channel_mean_kernel, dim3((unsigned)ceil_div(l.cin, 64), (unsigned)nb), 256, 0, st, cur, HW, l.cin_p, l.cin,
                m->se_mean.p);
      const size_t smem = (size_t)(l.cin + l.cmid) * sizeof(float);
      AM_CHECK(smem <= 48 * 1024, "encoder: squeeze-excite over %d channels does not fit in shared memory", l.cin);
      AM_LAUNCH(se_gate_kernel, nb, 256, smem, st, m->se_mean.p, l.cin, l.cmid, l.w_f32.p, l.bias.p, l.aux0.p, l.aux1.p, l.act,
                l.gate_act, m->se_gate.p);
      AM_LAUNCH(se_scale_kernel, grid_for((int64_t)nb * HW * (l.cin_p / 8)), 256, 0, st, cur, (int64_t)HW, l.cin_p, l.cin,
                m->se_gate.p, nb, dst);
    } else if (l.type == kPointwise) {
      const int64_t M = (int64_t)nb * s.H * s.W;
      gemm::Epilogue ep;
      ep.bias = l.bias.p;
      ep.act = l.act;
      if (l.residual) {
        ep.residual = block_in;
        ep.ld_res = l.cout_p;
      }
      if (m->use_simt_gemm) {
        for (int64_t m0 = 0; m0 < M; m0 += 32768) {
          const int64_t mm = std::min<int64_t>(32768, M - m0);
          gemm::Epilogue e2 = ep;
          if (e2.residual) e2.residual += m0 * l.cout_p;
          AM_TRY(gemm::gemm_bf16_simt(cur + m0 * l.cin_p, mm, l.cin_p, l.w_bf16.p, l.cout_p, l.cin_p, l.cin_p,
                                      dst + m0 * l.cout_p, l.cout_p, false, e2, st));
        }
      } else {
        AM_TRY(gemm::gemm_bf16(cur, M, l.cin_p, l.w_bf16.p, l.cout_p, l.cin_p, l.cin_p, dst, l.cout_p, false, ep,
                               /*m_fastest=*/false, st));
      }
    } else {
      set_error("encoder: layer %zu of type %d cannot run here", i, l.type);
      return AM_ERR_INVALID;
    }
    cur = dst;
  }
  *out = cur;
  *s_out = s;
  return AM_OK;
}

// EARLY phase of `nb` windows (log-mel at mel_dev): result appended to m->late_in at window offset b0
static int forward_early(am_model* m, const float* mel_dev, int nb, int b0, int T, cudaStream_t st) {
  Shape ss;
  int cs;
  const size_t split = late_start(m, T, &ss, &cs);
  const size_t per_win = (size_t)ss.H * ss.W * cs;
  const __nv_bfloat16* o;
  Shape so;
  return run_range(m, mel_dev, nullptr, Shape{0, 0}, 0, split, nb, T, m->late_in.p + (size_t)b0 * per_win, &o, &so, st);
}

// LATE phase for all `n` windows + the head program's pooling (for a 1x1 stride-s conv in front of the mean, the
// mean over the positions the conv visits commutes with it: the conv runs on the pooled rows)
static int forward_late(am_model* m, int n, int T, cudaStream_t st) {
  Shape ss;
  int cs;
  const size_t split = late_start(m, T, &ss, &cs);
  const size_t per_win = (size_t)ss.H * ss.W * cs;
  const HeadOp& pool = *m->head[0];
  for (int b0 = 0; b0 < n; b0 += m->late_sub) {
    const int nb = std::min(m->late_sub, n - b0);
    const __nv_bfloat16* o = m->late_in.p + (size_t)b0 * per_win;
    Shape so = ss;
    if (split < m->layers.size())
      AM_TRY(run_range(m, nullptr, o, ss, split, m->layers.size(), nb, T, nullptr, &o, &so, st));
    AM_LAUNCH(strided_mean_kernel, nb, 256, 0, st, o, so.H, so.W, m->head_cin_p, m->head_cin, pool.stride,
              m->regs[(size_t)pool.dst]->p + (size_t)b0 * m->head_cin);
  }
  return AM_OK;
}

// head row program for `n` windows at once; the last op writes out_dev
static int head_forward(am_model* m, int n, float* out_dev, cudaStream_t st) {
  if (n <= 0) return AM_OK;
  const int64_t rows = n;
  for (size_t q = 1; q < m->head.size(); ++q) {
    const HeadOp& h = *m->head[q];
    const bool last = q + 1 == m->head.size();
    float* dst = last ? out_dev : m->regs[(size_t)h.dst]->p;
    const float* a = h.a >= 0 ? m->regs[(size_t)h.a]->p : nullptr;
    const float* b = h.b >= 0 ? m->regs[(size_t)h.b]->p : nullptr;
    const int dim = m->reg_dim[(size_t)h.dst];
    const int egrid = grid_for(rows * dim);
    switch (h.kind) {
      case kVecLinear: {
        if (m->use_simt_gemm) {  // AM_GEMM_IMPL=simt: everything on CUDA cores (debug)
          AM_TRY(launch_linear(a, n, h.K, h.w.p, h.has_bias ? h.bias.p : nullptr, h.N, dst, h.act, st));
        } else {
          const int Kp = (int)round_up((size_t)h.K, 8);
          const int grid = (int)std::min<int64_t>(((int64_t)n * Kp + 255) / 256, (int64_t)sm_count() * 8);
          AM_LAUNCH(split3_kernel, grid, 256, 0, st, a, n, h.K, Kp, h.act, m->a3.p);
          gemm::Epilogue ep;
          ep.bias = h.has_bias ? h.bias.p : nullptr;
          AM_TRY(gemm::gemm_bf16(m->a3.p, n, 3 * Kp, h.w3.p, h.N, 3 * Kp, 3 * Kp, dst, h.N, /*d_is_f32=*/true, ep, false, st));
        }
        break;
      }
      case kVecUnary:
        AM_LAUNCH(vec_unary_kernel, egrid, 256, 0, st, a, rows * dim, h.act, dst);
        break;
      case kVecAdd:
        AM_LAUNCH(vec_add_kernel, egrid, 256, 0, st, a, b, rows * dim, dst);
        break;
      case kVecAffine:
        AM_LAUNCH(vec_affine_kernel, egrid, 256, 0, st, a, rows * dim, dim, h.has_w ? h.w.p : nullptr,
                  h.has_bias ? h.bias.p : nullptr, dst);
        break;
      case kVecLayerNorm:
        AM_LAUNCH(vec_layernorm_kernel, n, 256, 0, st, a, dim, h.w.p, h.bias.p, h.eps, dst);
        break;
      case kVecL2Norm:
        AM_LAUNCH(vec_l2norm_kernel, n, 256, 0, st, a, dim, h.eps2, dst);
        break;
      case kVecAddLnL2:
        AM_LAUNCH(head_finalize_kernel, n, 256, 0, st, a, b, dim, h.w.p, h.bias.p, h.eps, h.eps2, dst);
        break;
      default:
        set_error("encoder: head op %zu of kind %d cannot run here", q, h.kind);
        return AM_ERR_INVALID;
    }
  }
  return AM_OK;
}

// per-window elements of the largest activation written by layers [lo, hi) (lo == 0 includes the stem)
static size_t max_act_range(const am_model* m, int T, size_t lo, size_t hi) {
  Shape s = stem_out(*m->layers[0], T, m->n_mels);
  size_t mx = lo == 0 ? (size_t)s.H * s.W * m->layers[0]->cout_p : 0;
  for (size_t i = 1; i < hi && i < m->layers.size(); ++i) {
    const Layer& l = *m->layers[i];
    s = layer_out(l, s);
    if (i >= lo) mx = std::max(mx, (size_t)std::max(s.H, 0) * std::max(s.W, 0) * l.cout_p);
  }
  return mx;
}

static int ensure_workspace(am_model* m, int T, int nb, int n_total) {
  Shape ss;
  int cs;
  const size_t split = late_start(m, T, &ss, &cs);
  const size_t nt = (size_t)std::max(n_total, 1);
  static const int late_cap = std::getenv("AM_CLAP_LATE_SUB") ? std::max(1, std::atoi(std::getenv("AM_CLAP_LATE_SUB"))) : 256;
  m->late_sub = (int)std::min<size_t>(nt, (size_t)late_cap);
  const size_t need = std::max(max_act_range(m, T, 0, split) * (size_t)nb,
                               max_act_range(m, T, split, m->layers.size()) * (size_t)m->late_sub);
  if (need > m->act_elems) {
    for (auto& b : m->act) b.release();
    for (auto& b : m->act) AM_TRY(b.alloc(std::max<size_t>(need, 16)));
    m->act_elems = need;
  }
  AM_TRY(m->late_in.ensure(nt * (size_t)std::max(ss.H, 1) * std::max(ss.W, 1) * cs));
  size_t kmax = 8;
  for (size_t r = 0; r < m->regs.size(); ++r) {
    AM_TRY(m->regs[r]->ensure(nt * (size_t)m->reg_dim[r]));
    kmax = std::max(kmax, round_up((size_t)m->reg_dim[r], 8));
  }
  AM_TRY(m->a3.ensure(nt * 3 * kmax));
  return AM_OK;
}

}  // namespace am

using namespace am;

static int finish_load(am_model* m, const am::ModelSpec& spec, am_model** out) {
  int s = build_model(m, spec);
  if (s == AM_OK) s = m->stream.create();
  if (s != AM_OK) {
    delete m;
    return s;
  }
  const char* impl = std::getenv("AM_GEMM_IMPL");
  m->use_simt_gemm = (impl && std::strcmp(impl, "simt") == 0) ? 1 : 0;
  if (const char* sb = std::getenv("AM_CLAP_SUB_BATCH")) m->max_sub = std::max(1, std::atoi(sb));
  if (const char* fb = std::getenv("AM_FUSED_BLOCKS")) m->fused_mask = (unsigned)std::strtoul(fb, nullptr, 0);
  if (!m->use_simt_gemm && !gemm::available()) {
    set_error("am_clap_load: the tcgen05 GEMM path is unavailable on this device; no fallback is shipped");
    delete m;
    return AM_ERR_NO_DEVICE;
  }
  *out = m;
  return AM_OK;
}

// `path` (may be NULL) locates external tensor data of an ONNX model (model.onnx.data next to the model file)
static int load_any(const void* blob, size_t nbytes, const char* path, am::ModelSpec* spec) {
  if (looks_like_onnx(blob, nbytes)) return load_onnx_spec(blob, nbytes, path, spec);
  return parse_blob(blob, nbytes, spec);
}

extern "C" int am_clap_load_mem(const void* blob, size_t nbytes, am_model** out) {
  AM_CHECK(out != nullptr, "am_clap_load_mem: out is NULL");
  *out = nullptr;
  AM_CHECK(blob != nullptr && nbytes >= 20, "am_clap_load_mem: empty blob");
  ModelSpec spec;
  AM_TRY(load_any(blob, nbytes, nullptr, &spec));
  AM_TRY(ensure_init());
  return finish_load(new am_model(), spec, out);
}

static int read_file(const char* path, std::vector<uint8_t>* buf) {
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    set_error("am_clap_load: cannot open %s", path);
    return AM_ERR_IO;
  }
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf->resize((size_t)std::max<long>(n, 0));
  const size_t got = n > 0 ? std::fread(buf->data(), 1, (size_t)n, f) : 0;
  std::fclose(f);
  if (n <= 0 || got != (size_t)n) {
    set_error("am_clap_load: short read on %s", path);
    return AM_ERR_IO;
  }
  return AM_OK;
}

extern "C" int am_clap_load(const char* path, am_model** out) {
  AM_CHECK(out != nullptr && path != nullptr, "am_clap_load: NULL argument");
  *out = nullptr;
  std::vector<uint8_t> buf;
  AM_TRY(read_file(path, &buf));
  ModelSpec spec;
  AM_TRY(load_any(buf.data(), buf.size(), path, &spec));
  AM_TRY(ensure_init());
  return finish_load(new am_model(), spec, out);
}

static const char* act_name(int a) {
  static const char* names[] = {"none", "relu6", "relu", "hardswish", "gelu", "sigmoid", "hardsigmoid", "tanh"};
  return a >= 0 && a < 8 ? names[a] : "?";
}

// one line per layer / head op of the lowered program
static std::string describe_spec(const ModelSpec& sp) {
  char line[256];
  std::string o;
  std::snprintf(line, sizeof line, "source %s; n_mels %d; embedding %d; %zu layers; %zu head ops\n", sp.source.c_str(), sp.n_mels,
                sp.emb, sp.layers.size(), sp.head.size());
  o += line;
  for (size_t i = 0; i < sp.layers.size(); ++i) {
    const LayerSpec& l = sp.layers[i];
    static const char* tn[] = {"stem", "pointwise", "depthwise", "head", "conv_first", "squeeze_excite"};
    std::snprintf(line, sizeof line, "L%-3zu %-14s %4d -> %4d  k %dx%d s %d pad %d,%d,%d,%d act %s%s%s%s", i,
                  l.type >= 0 && l.type < 6 ? tn[l.type] : "?", l.cin, l.cout, l.kh, l.kw, l.stride, l.pad_t, l.pad_b, l.pad_l,
                  l.pad_r, act_name(l.act), l.residual ? " +residual" : "", l.block_start ? " [block]" : "",
                  (l.type == kStem || l.type == kConvFirst) ? (l.h_is_time ? " H=time" : " H=mel") : "");
    o += line;
    if (l.type == kSqueezeExcite) {
      std::snprintf(line, sizeof line, " mid %d gate %s", l.cmid, act_name(l.gate_act));
      o += line;
    }
    o += "\n";
  }
  static const char* hn[] = {"pool", "linear", "unary", "add", "affine", "layernorm", "l2norm", "add_layernorm_l2"};
  for (size_t i = 0; i < sp.head.size(); ++i) {
    const VecOp& h = sp.head[i];
    std::snprintf(line, sizeof line, "H%-3zu %-16s r%d%s -> r%d  K %d N %d act %s stride %d bias %d\n", i,
                  h.kind >= 0 && h.kind < 8 ? hn[h.kind] : "?", h.a, h.b >= 0 ? (std::string(",r") + std::to_string(h.b)).c_str() : "",
                  h.dst, h.K, h.N, act_name(h.act), h.stride, h.bias.empty() ? 0 : 1);
    o += line;
  }
  return o;
}

// host-only (no GPU): parse + lower a model file and describe the resulting program
extern "C" int am_clap_describe_file(const char* path, char* buf, int cap) {
  AM_CHECK(path != nullptr, "am_clap_describe_file: NULL path");
  std::vector<uint8_t> data;
  AM_TRY(read_file(path, &data));
  ModelSpec spec;
  AM_TRY(load_any(data.data(), data.size(), path, &spec));
  const std::string d = describe_spec(spec);
  if (buf && cap > 0) {
    const size_t n = std::min((size_t)cap - 1, d.size());
    std::memcpy(buf, d.data(), n);
    buf[n] = 0;
  }
  return (int)d.size() + 1;
}

// frees every workspace buffer (activations, staging, head registers); the next call re-allocates what it needs.
// The cleanup step of the reference's OOM retry (tasks/clap_analyzer.py:536-549 -> cleanup_cuda_memory).
extern "C" int am_clap_release_workspace(am_model* m) {
  AM_CHECK(m != nullptr, "am_clap_release_workspace: NULL model");
  AM_CHECK(m->n_submitted == m->n_collected, "am_clap_release_workspace: submitted calls are still in flight");
  AM_CUDA(cudaDeviceSynchronize());
  for (auto& b : m->act) b.release();
  m->act_elems = 0;
  m->late_in.release();
  m->a3.release();
  m->mel_ws.release();
  m->seg_emb.release();
  m->se_mean.release();
  m->se_gate.release();
  for (auto& r : m->regs) r->release();
  for (auto& b : m->pcm_stage) b.release();
  m->off_stage.release();
  m->out_stage.release();
  m->slot_used[0] = m->slot_used[1] = false;
  return AM_OK;
}

extern "C" void am_clap_free(am_model* m) {
  if (m) cudaDeviceSynchronize();
  delete m;
}
extern "C" int am_clap_embedding_dim(const am_model* m) { return m ? m->emb : 0; }
extern "C" int am_clap_n_mels(const am_model* m) { return m ? m->n_mels : 0; }

static double head_macs(const am_model* m) {
  double macs = 0.0;
  for (const auto& h : m->head)
    if (h->kind == kVecLinear) macs += (double)h->K * h->N;
  return macs;
}

extern "C" double am_clap_flops_per_segment(const am_model* m, int T) {
  if (!m || m->layers.empty()) return 0.0;
  Shape s = stem_out(*m->layers[0], T, m->n_mels);
  const Layer& f = *m->layers[0];
  double macs = f.type == kStem ? (double)s.H * s.W * (9.0 + f.cout) : (double)s.H * s.W * f.kh * f.kw * f.cout;
  for (size_t i = 1; i < m->layers.size(); ++i) {
    const Layer& l = *m->layers[i];
    if (l.type == kDepthwise) {
      s = dw_out(l, s);
      macs += (double)s.H * s.W * l.cin * l.kh * l.kw;
    } else if (l.type == kPointwise) {
      macs += (double)s.H * s.W * l.cin * (double)l.cout;
    } else if (l.type == kSqueezeExcite) {
      macs += 2.0 * l.cin * l.cmid;
    }
  }
  return 2.0 * (macs + head_macs(m));
}

// flops (2 x MAC) of one window of T frames executed by the standalone GEMM kernel and by the fused
// block kernel (pointwise + depthwise inside fused blocks), following the same plan forward_sub uses
extern "C" int am_clap_flops_split(const am_model* m, int T, double* gemm_flops, double* fused_flops,
                                   double* fused_bytes) {
  AM_CHECK(m && gemm_flops && fused_flops && fused_bytes && !m->layers.empty(), "am_clap_flops_split: bad argument");
  Shape s = stem_out(*m->layers[0], T, m->n_mels);
  double g = 0.0, f = 0.0, fb = 0.0;
  for (size_t i = 1; i < m->layers.size(); ++i) {
    const Layer& l = *m->layers[i];
    bool fused_here = false;
    if (l.block_start && !m->use_simt_gemm) {
      const int q = block_index_at(m, i);
      fused::BlockDesc d;
      fused::Plan pl;
      if (q >= 0 && block_desc(m, q, s, false, &d, &pl)) {
        const am_model::Block& blk = m->blocks[q];
        const Layer& dwl = *m->layers[blk.dw];
        const Layer& pj = *m->layers[blk.proj];
        const Shape o = dw_out(dwl, s);
        double macs = (double)o.H * o.W * dwl.cin * 9.0 + (double)o.H * o.W * pj.cin * (double)pj.cout;
        if (blk.expand >= 0) macs += (double)s.H * s.W * m->layers[blk.expand]->cin * (double)m->layers[blk.expand]->cout;
        f += 2.0 * macs;
        fb += 2.0 * ((double)s.H * s.W * d.cin_p + (double)o.H * o.W * d.cout_p);  // X in + Y out, 16-bit
        s = o;
        i = (size_t)blk.proj;
        fused_here = true;
      }
    }
    if (fused_here) continue;
    if (l.type == kDepthwise) {
      s = dw_out(l, s);
    } else if (l.type == kPointwise) {
      g += 2.0 * (double)s.H * s.W * l.cin * (double)l.cout;
    }
  }
  *gemm_flops = g;
  *fused_flops = f;
  *fused_bytes = fb;
  return AM_OK;
}

extern "C" int am_clap_embed_dev(am_model* m, const float* mel_dev, int B, int T, float* out_dev, void* stream) {
  AM_CHECK(m && mel_dev && out_dev, "am_clap_embed_dev: NULL argument");
  AM_CHECK(B >= 0 && T > 0, "am_clap_embed_dev: bad shape B=%d T=%d", B, T);
  cudaStream_t st = (cudaStream_t)stream;
  const int sub = std::min(std::max(B, 1), m->max_sub);
  AM_TRY(ensure_workspace(m, T, sub, B));
  for (int b0 = 0; b0 < B; b0 += sub) {
    const int nb = std::min(sub, B - b0);
    AM_TRY(forward_early(m, mel_dev + (size_t)b0 * m->n_mels * T, nb, b0, T, st));
  }
  AM_TRY(forward_late(m, B, T, st));
  return head_forward(m, B, out_dev, st);
}

extern "C" int am_clap_embed(am_model* m, const float* mel, int B, int T, float* out) {
  AM_CHECK(m && mel && out, "am_clap_embed: NULL argument");
  AM_CHECK(B >= 0 && T > 0, "am_clap_embed: bad shape B=%d T=%d", B, T);
  if (B == 0) return AM_OK;
  DevBuf<float> d_mel, d_out;
  const size_t mel_elems = (size_t)B * m->n_mels * T;
  AM_TRY(d_mel.alloc(mel_elems));
  AM_TRY(d_out.alloc((size_t)B * m->emb));
  cudaStream_t st = m->stream.s;
  AM_CUDA(cudaMemcpyAsync(d_mel.p, mel, mel_elems * 4, cudaMemcpyHostToDevice, st));
  AM_TRY(am_clap_embed_dev(m, d_mel.p, B, T, d_out.p, st));
  AM_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)B * m->emb * 4, cudaMemcpyDeviceToHost, st));
  AM_CUDA(cudaStreamSynchronize(st));
  return AM_OK;
}

extern "C" int am_clap_embed_tracks_dev(am_model* m, const am_mel_plan* plan, const int16_t* pcm_dev, int n_samples,
                                        const int32_t* seg_offsets_dev, int n_tracks, int n_segments, float* out_dev,
                                        void* stream) {
  AM_CHECK(m && plan && out_dev && seg_offsets_dev, "am_clap_embed_tracks_dev: NULL argument");
  AM_CHECK(n_tracks >= 0 && n_segments >= 0 && (pcm_dev || n_segments == 0), "am_clap_embed_tracks_dev: bad sizes");
  if (n_tracks == 0) return AM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int hop = mel_plan_hop(plan);
  const int T = 1 + n_samples / hop;
  const int sub = std::min(std::max(n_segments, 1), m->max_sub);
  AM_TRY(ensure_workspace(m, T, sub, n_segments));
  AM_TRY(m->mel_ws.ensure((size_t)sub * m->n_mels * T));
  AM_TRY(m->seg_emb.ensure((size_t)std::max(n_segments, 1) * m->emb));
  for (int b0 = 0; b0 < n_segments; b0 += sub) {
    const int nb = std::min(sub, n_segments - b0);
    AM_TRY(am_mel_batch_dev(plan, pcm_dev + (size_t)b0 * n_samples, 1, nb, n_samples, m->mel_ws.p, st));
    AM_TRY(forward_early(m, m->mel_ws.p, nb, b0, T, st));
  }
  AM_TRY(forward_late(m, n_segments, T, st));
  AM_TRY(head_forward(m, n_segments, m->seg_emb.p, st));
  AM_LAUNCH(track_pool_kernel, n_tracks, 256, 0, st, m->seg_emb.p, seg_offsets_dev, m->emb, out_dev);
  return AM_OK;
}

extern "C" int am_clap_embed_tracks_collect(am_model* m) {
  AM_CHECK(m != nullptr, "am_clap_embed_tracks_collect: NULL model");
  AM_CHECK(m->n_collected < m->n_submitted, "am_clap_embed_tracks_collect: nothing was submitted");
  am_model::Ticket& t = m->tickets[m->n_collected & 1];
  ++m->n_collected;
  t.open = false;
  if (t.count == 0) return AM_OK;
  AM_CUDA(cudaEventSynchronize(t.ready));
  std::memcpy(t.user_out, t.stage.p, t.count * sizeof(float));
  return AM_OK;
}

extern "C" int am_clap_embed_tracks_submit(am_model* m, const am_mel_cfg* cfg, const int16_t* pcm, int n_samples,
                                           const int32_t* seg_offsets, int n_tracks, float* out) {
  AM_CHECK(m && cfg && seg_offsets && out, "am_clap_embed_tracks: NULL argument");
  AM_CHECK(n_tracks >= 0, "am_clap_embed_tracks: negative track count");
  AM_CHECK(cfg->n_mels == m->n_mels && cfg->transpose == 0, "am_clap_embed_tracks: mel cfg does not match the model");
  AM_CHECK(m->n_submitted - m->n_collected < 2, "am_clap_embed_tracks_submit: two calls are already in flight; collect one");
  const bool warm = m->n_submitted > m->n_collected;  // an earlier batch is still running
  am_model::Ticket& tk = m->tickets[m->n_submitted & 1];
  if (n_tracks == 0) {
    tk.user_out = out;
    tk.count = 0;
    tk.open = true;
    ++m->n_submitted;
    return AM_OK;
  }
  // Everything that can fail without touching the stream (argument checks, allocations) happens BEFORE the ticket
  // is opened: a failed submit leaves n_submitted == what it was, so the session stays usable (a single OOM in a
  // bulk scan used to leave an uncollectable ticket behind).
  const int n_segments = seg_offsets[n_tracks];
  AM_CHECK(seg_offsets[0] == 0 && n_segments >= 0 && (pcm || n_segments == 0), "am_clap_embed_tracks: bad seg_offsets");
  if (!m->host_plan || std::memcmp(&m->host_plan_cfg, cfg, sizeof(am_mel_cfg)) != 0) {
    if (m->host_plan) am_mel_plan_free(m->host_plan);
    m->host_plan = nullptr;
    AM_TRY(am_mel_plan_create(cfg, &m->host_plan));
    m->host_plan_cfg = *cfg;
  }
  am_mel_plan* plan = m->host_plan;
  cudaStream_t st = m->stream.s;
  AM_TRY(m->copy_stream.create());
  cudaStream_t cs = m->copy_stream.s;
  for (int i = 0; i < 2; ++i) {
    if (!m->ev_copied[i]) AM_CUDA(cudaEventCreateWithFlags(&m->ev_copied[i], cudaEventDisableTiming));
    if (!m->ev_done[i]) AM_CUDA(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
  }
  const int T = 1 + n_samples / cfg->hop;
  const int sub = std::min(std::max(n_segments, 1), m->max_sub);
  // (growing a workspace buffer while a batch is in flight is safe: cudaFree waits for the device)
  AM_TRY(ensure_workspace(m, T, sub, n_segments));
  AM_TRY(m->mel_ws.ensure((size_t)sub * m->n_mels * T));
  AM_TRY(m->seg_emb.ensure((size_t)std::max(n_segments, 1) * m->emb));
  for (auto& b : m->pcm_stage) AM_TRY(b.ensure((size_t)sub * n_samples));
  AM_TRY(m->off_stage.ensure((size_t)n_tracks + 1));
  AM_TRY(m->out_stage.ensure((size_t)n_tracks * m->emb));
  const size_t out_count = (size_t)n_tracks * m->emb;
  AM_TRY(tk.stage.ensure(out_count));
  if (!tk.ready) AM_CUDA(cudaEventCreateWithFlags(&tk.ready, cudaEventDisableTiming));
  if (std::getenv("AM_TEST_FAIL_SUBMIT")) {  // test hook: a submit that fails after its allocations (tests/test_gpu_encoder.py)
    set_error("am_clap_embed_tracks_submit: out of memory (injected by AM_TEST_FAIL_SUBMIT)");
    return AM_ERR_OOM;
  }
  auto enqueue = [&]() -> int {
    AM_CUDA(cudaMemcpyAsync(m->off_stage.p, seg_offsets, (size_t)(n_tracks + 1) * 4, cudaMemcpyHostToDevice, st));
    // double-buffered pipeline: H2D of chunk c+1 (copy stream) overlaps mel + early trunk of chunk c.  The
    // copy runs ~3x faster than the compute it hides under, so only the FIRST chunk's copy is exposed:
    // chunks grow 16, 32, 64, then `sub`: a copy is ~2.2x faster than the compute it hides under, so each chunk may
    // be at most ~2.2x the previous one (8 / 24 / 96 measured slower: the small chunks under-fill the GPU).
    int c = 0;
    for (int b0 = 0; b0 < n_segments; ++c) {
      // a batch submitted while another is still in flight has its copies hidden under that one: two big chunks
      const int want = warm ? (sub + 1) / 2 : (c == 0 ? 16 : (c == 1 ? 32 : (c == 2 ? 64 : sub)));
      const int nb = std::min(std::min(want, sub), n_segments - b0);
      const int slot = c & 1;
      // slot free again (its last reader may belong to the previous, still running, submitted call)
      if (m->slot_used[slot]) AM_CUDA(cudaStreamWaitEvent(cs, m->ev_done[slot], 0));
      m->slot_used[slot] = true;
      AM_CUDA(cudaMemcpyAsync(m->pcm_stage[slot].p, pcm + (size_t)b0 * n_samples, (size_t)nb * n_samples * 2,
                              cudaMemcpyHostToDevice, cs));
      AM_CUDA(cudaEventRecord(m->ev_copied[slot], cs));
      AM_CUDA(cudaStreamWaitEvent(st, m->ev_copied[slot], 0));
      AM_TRY(am_mel_batch_dev(plan, m->pcm_stage[slot].p, 1, nb, n_samples, m->mel_ws.p, st));
      AM_CUDA(cudaEventRecord(m->ev_done[slot], st));  // the mel kernel was the staging slot's only reader
      AM_TRY(forward_early(m, m->mel_ws.p, nb, b0, T, st));
      b0 += nb;
    }
    AM_TRY(forward_late(m, n_segments, T, st));
    AM_TRY(head_forward(m, n_segments, m->seg_emb.p, st));
    AM_LAUNCH(track_pool_kernel, n_tracks, 256, 0, st, m->seg_emb.p, m->off_stage.p, m->emb, m->out_stage.p);
    AM_CUDA(cudaMemcpyAsync(tk.stage.p, m->out_stage.p, out_count * 4, cudaMemcpyDeviceToHost, st));
    AM_CUDA(cudaEventRecord(tk.ready, st));
    return AM_OK;
  };
  const int s = enqueue();
  if (s != AM_OK) {  // a launch / copy failed half way: let the stream drain, keep the ticket closed
    cudaStreamSynchronize(st);
    cudaStreamSynchronize(cs);
    return s;
  }
  tk.user_out = out;
  tk.count = out_count;
  tk.open = true;
  ++m->n_submitted;
  return AM_OK;
}

extern "C" int am_clap_embed_tracks(am_model* m, const am_mel_cfg* cfg, const int16_t* pcm, int n_samples,
                                    const int32_t* seg_offsets, int n_tracks, float* out) {
  AM_CHECK(m != nullptr, "am_clap_embed_tracks: NULL argument");
  AM_CHECK(m->n_submitted == m->n_collected, "am_clap_embed_tracks: submitted calls are still in flight; collect them first");
  AM_TRY(am_clap_embed_tracks_submit(m, cfg, pcm, n_samples, seg_offsets, n_tracks, out));  // transactional: no ticket on failure
  return am_clap_embed_tracks_collect(m);
}
