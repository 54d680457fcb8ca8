// Fused inverted-residual block, channel-per-lane formulation (sm_100a):
//
//     Y = [X +] project( relu6( dw3x3_s( relu6( expand(X) ) ) ) )        all BatchNorms folded
//
// Same contract as fused_block.cu (student_clap/models/student_onnx_model.py:95-148 block shape), other data flow.
// fused_block.cu computes the expansion as D1[pixel, channel] and has to round-trip it through shared memory (TMEM ->
// registers -> E tile -> 9-tap reads) because a thread owns one PIXEL of D1 and the depthwise needs its neighbours.
// Here the expansion GEMM is transposed,
//
//     D1T[channel, pixel] = W1_j[128 ch x Cin] . X^T[Cin x M1 halo pixels]      (A = weights, B = the X halo tile)
//
// so a TMEM lane is a CHANNEL and the columns are the halo pixels in row-major order: a thread reads whole pixel rows
// of its own channel straight from TMEM (tcgen05.ld.32x32b.x16/.x32) and runs the 3x3 depthwise in registers on fp16
// pixel pairs (HFMA2).  No E tile, no shared-memory taps, the 9 depthwise weights are per-thread scalars.  The
// depthwise output of a thread is a run of consecutive pixels of one channel, which is exactly a 16-byte chunk of an
// MN-major (pixels contiguous) SWIZZLE_128B operand: it is stored once and consumed by the projection MMA
//
//     D2[pixel, Cout] += A2[128 px x 128 ch, MN-major] . W2_j^T[128 ch x Cout]   (layout checked by tools/mn_probe.py)
//
// One CTA = 16 compute warps + four single-lane control warps (TMA producer, expansion MMA issuer, projection MMA issuer,
// output store), persistent over (tile, 128-channel chunk) items.
//   TMEM lane quadrant q = warp & 3 -> channels [32q, 32q + 32) of the chunk; the 4 warps sharing a quadrant split the
//   output tile (TH rows x Wo) into 4 row groups, or 2 x 2 (row, column half) when TH == 2.
//   The residual is one more MMA chain (X centre rows x a 16 x 16 identity into D2: exact in fp32); epilogue 2 of a tile
//   runs under the next tile's first chunk and leaves through a staged bulk copy.
// Measured on B200: the taps are bound by the fp16 pipes (HFMA2 / PRMT / F2FP issue at 2 cycles per
// warp-instruction per sub-partition) and the kernel is very sensitive to spills at its 96-register cap.
// The expansion MMA covers exactly the M1 halo pixels (N = M1, or two halves when M1 > 256): nothing is padded to
// 128-row tiles, and the X tile is read from shared memory once per 128 channels instead of once per 64.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fused_block.cuh"
#include "gemm_tcgen05.cuh"
#include "ptx_sm100.cuh"

// 32-column segments: issuing the next row's TMEM load before this row's FMAs keeps 32 more registers live and the kernel
// spills at its 96-register cap (640 threads); measured on B200 the spill-free order wins (stride 1: 3.66 -> 3.53 ms per
// 256 windows over the four blocks; stride 2: no difference), so the early load is kept for 16-column segments only.
#ifndef AM_FUSEDT_PREFETCH32_S1
#define AM_FUSEDT_PREFETCH32_S1 0
#endif
#ifndef AM_FUSEDT_PREFETCH32_S2
#define AM_FUSEDT_PREFETCH32_S2 0
#endif

#ifndef AM_FUSEDT_BACKOFF
#define AM_FUSEDT_BACKOFF 0
#endif

namespace am {
namespace fusedt {

using namespace ptx;

constexpr int kComputeWarps = 16;
constexpr int kProducerWarp = 16;                 // TMA loads (one elected lane)
constexpr int kMmaWarp = 17;                      // expansion tcgen05.mma issue (one elected lane), owns the TMEM allocation
constexpr int kMma2Warp = 18;                     // projection (+ residual) tcgen05.mma issue (one elected lane)
constexpr int kStoreWarp = 19;                    // bulk copy of a finished tile's staged output to Y (one elected lane)
constexpr int kThreads = (kComputeWarps + 4) * 32;
constexpr int kChunk = 128;                       // expanded channels per work item = TMEM lanes
constexpr uint32_t kW1Stage = 128u * 128u;        // one W1 k-block: [128 channels x 64 k] bf16
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kTraceItems = 96;
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B

struct Args {
  int B, H, W, Ho, Wo;
  int cin_p, cmid_p, cout_p;
  int residual;
  int TH, IH, M1, M2;
  int n_parts, n_part;          // expansion MMA: N = n_part columns, n_parts of them (M1 = n_parts * n_part)
  int kb_in, n_chunks;
  int csub;                     // column segments per output row (1 or 2); row groups = 4 / csub
  int tiles_per_window, total_tiles;
  int s1, s2;                   // W1 / W2 ring stages (k-blocks)
  int a2_bufs, d1_bufs;
  uint32_t a2_bytes, a2_lbo;    // one A2 buffer; byte stride between its two 64-pixel atoms (0: second aliases the first)
  uint32_t a2_stride;           // bytes between the two A2 buffers (>= a2_bytes: a buffer also stages one output tile)
  const uint4* cpack;           // [round_up(cmid_p, 128)][2]: per-channel constants, see pack_consts()
  const float* b2;              // [cout_p]
  __nv_bfloat16* Y;             // [B, Ho, Wo, cout_p]
  const __nv_bfloat16* Xg;      // [B, H, W, cin_p]: the residual is re-read from global (L2) in epilogue 2
  uint32_t off_x, off_a2, off_w1, off_w2, off_small, off_bar, off_id;
  int stage_out;                // epilogue 2 stages the output tile in the A2 buffers and writes it with one bulk copy
  uint32_t x_kb_bytes, w2_stage_bytes;
  long long* trace;             // debug (AM_FUSED_TRACE=1, -DAM_FUSED_TRACE_BUILD): [kTraceItems][16] clock64 stamps of CTA 0
};

// barrier slots (uint64_t each)
constexpr int kBarXFull = 0, kBarXFree = 1, kBarW1Full = 2, kBarW1Empty = 10, kBarW2Full = 18, kBarW2Empty = 26,
              kBarMma1 = 34, kBarD1Free = 36, kBarA2Full = 38, kBarMma2 = 40, kBarTile = 42, kBarStage = 43, kBarStaged = 44, kBarCount = 45;

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return (smem_addr >> 4) & 0x3fffu; }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
// (hi half of a, lo half of b): the pixel pair one position to the left of pair b
__device__ __forceinline__ __half2 shift_pair(__half2 a, __half2 b) { return u2h(__byte_perm(h2u(a), h2u(b), 0x5432)); }

// compute-warp wait: back off between polls so spinning warps leave the issue slots to the working ones
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n.reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (AM_FUSEDT_BACKOFF) __nanosleep(32);
  }
}

// Phase timeline for tools/gpu_trace.sh; compiled in only with -DAM_FUSED_TRACE_BUILD.
#ifdef AM_FUSED_TRACE_BUILD
#define AMT_TRACE(slot_)                                                                                   \
  do {                                                                                                     \
    if (a.trace && blockIdx.x == 0 && lane == 0 && w < kTraceItems && (warp == 0 || warp == kMmaWarp))      \
      a.trace[w * 16 + (slot_)] = clock64();                                                               \
  } while (0)
#else
#define AMT_TRACE(slot_) do { } while (0)
#endif

// Barrier protocol.  Item w = (tile, chunk) in flat order; D1 set ds = w % d1_bufs, A2 buffer slot = w % a2_bufs.
//   x_full            TMA: X halo tile of a tile landed                               (1 / tile)
//   x_free            last expansion MMA of a tile retired (epilogue 2 reads the residual from global memory)
//   w1_full/empty[s]  W1 k-block ring: TMA fills, the commit after the MMAs of that k-block frees
//   w2_full/empty[s]  W2 k-block ring, same
//   mma1[ds]          D1T(w) complete                                                (commit)
//   d1free[ds]        16 compute warps finished reading D1T(w)
//   a2_full[slot]     16 compute warps finished writing A2(w)
//   mma2[slot]        projection MMAs of item w retired: A2 slot free; last chunk: D2 ready  (commit)
//   tile              16 compute warps finished reading D2 in epilogue 2: D2 reusable
//   staged            16 compute warps finished writing a tile's output into its staging copy (the A2 buffer of the tile's
//                     last chunk, idle until the chunk after next)
//   stage             the store warp's bulk copy finished reading that staging copy
template <int kStride, int kSeg, int kRows>
__global__ void __launch_bounds__(kThreads, 1)
fused_block_t_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                     const __grid_constant__ CUtensorMap map_w2, const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sm = smem_u32(smem);
  const uint32_t s_x = sm + a.off_x, s_a2 = sm + a.off_a2, s_w1 = sm + a.off_w1, s_w2 = sm + a.off_w2;
  float* s_b2 = reinterpret_cast<float*>(smem + a.off_small);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + kBarCount);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_w1);
    prefetch_tensormap(&map_w2);
    mbar_init(&bars[kBarXFull], 1);
    mbar_init(&bars[kBarXFree], 1);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bars[kBarW1Full + i], 1);
      mbar_init(&bars[kBarW1Empty + i], 1);
      mbar_init(&bars[kBarW2Full + i], 1);
      mbar_init(&bars[kBarW2Empty + i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[kBarMma1 + i], 1);
      mbar_init(&bars[kBarD1Free + i], kComputeWarps);
      mbar_init(&bars[kBarA2Full + i], kComputeWarps);
      mbar_init(&bars[kBarMma2 + i], 1);
    }
    mbar_init(&bars[kBarTile], kComputeWarps);
    mbar_init(&bars[kBarStage], 1);
    mbar_init(&bars[kBarStaged], kComputeWarps);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_ptr, kTmemCols);
  for (int i = tid; i < a.cout_p; i += kThreads) s_b2[i] = a.b2[i];
  if (a.residual) {
    // 16 x 16 bf16 identity, K-major SWIZZLE_128B rows of 128 bytes: the residual add is  D2[:, 16 g .. 16 g + 16) += X_centre[:, same] . I
    for (int i = tid; i < 2048 / 4; i += kThreads) reinterpret_cast<uint32_t*>(smem + a.off_id)[i] = 0u;
    __syncthreads();
    if (tid < 16) *reinterpret_cast<uint16_t*>(smem + a.off_id + sw128_offset((uint32_t)tid, (uint32_t)(tid >> 3)) + (tid & 7) * 2) = 0x3f80u;
    fence_proxy_async();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_d2 = tmem_base + (uint32_t)(a.d1_bufs * a.M1);

  const int n_my_tiles = blockIdx.x < a.total_tiles ? (a.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_items = n_my_tiles * a.n_chunks;

  if (warp == kProducerWarp) {
    // =========================== TMA producer ===========================
    // Loads are issued in the order the MMA warp consumes them -- W1(0), then per item w: [X(next tile)], W1(w+1),
    // W2(w) -- so a full ring never blocks a load that an earlier consumer step is waiting for.
    if (elect_one_sync()) {
      int slot1 = 0, round1 = 0, slot2 = 0, round2 = 0;
      auto load_x = [&](int ti) {
        if (ti > 0) mbar_wait(&bars[kBarXFree], (uint32_t)(ti - 1) & 1u);
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        const int b = tile / a.tiles_per_window;
        const int h0 = (tile - b * a.tiles_per_window) * a.TH * kStride - 1;
        mbar_expect_tx(&bars[kBarXFull], (uint32_t)a.kb_in * (uint32_t)a.M1 * 128u);
        for (int kb = 0; kb < a.kb_in; ++kb)
          tma_load_4d(smem + a.off_x + kb * a.x_kb_bytes, &map_x, &bars[kBarXFull], kb * 64, 0, h0, b);
      };
      auto load_w1 = [&](int j) {
        for (int kb = 0; kb < a.kb_in; ++kb) {
          if (round1 > 0) mbar_wait(&bars[kBarW1Empty + slot1], (uint32_t)(round1 - 1) & 1u);
          mbar_expect_tx(&bars[kBarW1Full + slot1], kW1Stage);
          tma_load_2d(smem + a.off_w1 + slot1 * kW1Stage, &map_w1, &bars[kBarW1Full + slot1], kb * 64, j * kChunk);
          if (++slot1 == a.s1) { slot1 = 0; ++round1; }
        }
      };
      auto load_w2 = [&](int j) {
        const int kpc = min(2, (a.cmid_p - j * kChunk + 63) >> 6);
        for (int kb2 = 0; kb2 < kpc; ++kb2) {
          if (round2 > 0) mbar_wait(&bars[kBarW2Empty + slot2], (uint32_t)(round2 - 1) & 1u);
          mbar_expect_tx(&bars[kBarW2Full + slot2], (uint32_t)a.cout_p * 128u);
          tma_load_2d(smem + a.off_w2 + slot2 * a.w2_stage_bytes, &map_w2, &bars[kBarW2Full + slot2], j * kChunk + kb2 * 64, 0);
          if (++slot2 == a.s2) { slot2 = 0; ++round2; }
        }
      };
      if (n_items > 0) {
        load_x(0);
        load_w1(0);
      }
      int ti = 0, j = 0;
      for (int w = 0; w < n_items; ++w) {
        const bool last = (j == a.n_chunks - 1), nxt = (w + 1 < n_items);
        const int j1 = last ? 0 : j + 1;
        if (nxt && last) load_x(ti + 1);   // X is free once the tile's last expansion MMA retired
        if (nxt) load_w1(j1);
        load_w2(j);
        if (last) { j = 0; ++ti; } else ++j;
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== expansion MMA issuer ===========================
    // Two issuing threads (this one: D1T = W1 . X^T, the next warp: D2 += A2 . W2^T) on two SM sub-partitions: each
    // blocks on its own operands only, and the compute warps keep the HFMA2 pipes so busy that ONE thread got an issue
    // slot too rarely to feed the tensor pipe (measured 160 - 290 cycles per MMA issued, 50 - 100 to execute).
    if (elect_one_sync()) {
      const uint32_t idesc1 = make_idesc(128, a.n_part);                       // bf16 x bf16, both K-major
      const uint32_t part_step = (uint32_t)(a.n_part * 128) >> 4;
      const bool two_parts = a.n_parts == 2;
      const int ksteps_last = (a.cin_p - (a.kb_in - 1) * 64) >> 4;             // k-steps of the last (ragged) k-block
      int slot1 = 0, round1 = 0;
      int ti = 0, j = 0;
      for (int w = 0; w < n_items; ++w) {
        const int ds = (a.d1_bufs == 2) ? (w & 1) : 0;
        const int use = (a.d1_bufs == 2) ? (w >> 1) : w;
        if (a.trace && blockIdx.x == 0 && w < kTraceItems) a.trace[w * 16 + 8] = clock64();
        if (use > 0) mbar_wait(&bars[kBarD1Free + ds], (uint32_t)(use - 1) & 1u);
        if (j == 0) mbar_wait(&bars[kBarXFull], (uint32_t)ti & 1u);
        if (a.trace && blockIdx.x == 0 && w < kTraceItems) a.trace[w * 16 + 9] = clock64();
        const uint32_t d = tmem_base + (uint32_t)(ds * a.M1);
        uint32_t db = desc_lo(s_x);
        const uint32_t db_step = a.x_kb_bytes >> 4;
        for (int kb = 0; kb < a.kb_in; ++kb, db += db_step) {
          mbar_wait(&bars[kBarW1Full + slot1], (uint32_t)round1 & 1u);
          tcgen05_fence_after();
          const int ksteps = (kb + 1 < a.kb_in) ? 4 : ksteps_last;
          const uint32_t da = desc_lo(s_w1 + (uint32_t)slot1 * kW1Stage);
          if (!two_parts) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ksteps) umma_f16_lo(d, da + (uint32_t)(ks * 2), db + (uint32_t)(ks * 2), kDescHi, idesc1, (kb | ks) ? 1u : 0u);
          } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ksteps) {
                umma_f16_lo(d, da + (uint32_t)(ks * 2), db + (uint32_t)(ks * 2), kDescHi, idesc1, (kb | ks) ? 1u : 0u);
                umma_f16_lo(d + (uint32_t)a.n_part, da + (uint32_t)(ks * 2), db + part_step + (uint32_t)(ks * 2), kDescHi, idesc1,
                            (kb | ks) ? 1u : 0u);
              }
          }
          umma_commit(&bars[kBarW1Empty + slot1]);
          if (++slot1 == a.s1) { slot1 = 0; ++round1; }
        }
        umma_commit(&bars[kBarMma1 + ds]);
        if (j == a.n_chunks - 1) umma_commit(&bars[kBarXFree]);
        if (a.trace && blockIdx.x == 0 && w < kTraceItems) a.trace[w * 16 + 10] = clock64();
        if (++j == a.n_chunks) { j = 0; ++ti; }
      }
    }
  } else if (warp == kMma2Warp) {
    // =========================== projection MMA issuer ===========================
    if (elect_one_sync()) {
      const uint32_t idesc2 = make_idesc_f16(128, a.cout_p) | (1u << 15);      // fp16 x fp16, A MN-major
      const uint32_t idesc_res = make_idesc(128, 16);                           // residual: X (bf16) x identity (bf16)
      int slot2 = 0, round2 = 0;
      int ti = 0, j = 0;
      for (int w = 0; w < n_items; ++w) {
        const int slot = (a.a2_bufs == 2) ? (w & 1) : 0;
        const int use = (a.a2_bufs == 2) ? (w >> 1) : w;
        if (a.trace && blockIdx.x == 0 && w < kTraceItems) a.trace[w * 16 + 11] = clock64();
        mbar_wait(&bars[kBarA2Full + slot], (uint32_t)use & 1u);
        if (j == 0 && ti > 0) mbar_wait(&bars[kBarTile], (uint32_t)(ti - 1) & 1u);   // D2 drained by epilogue 2
        if (j == 0 && a.residual) mbar_wait(&bars[kBarXFull], (uint32_t)ti & 1u);    // (complete since chunk 0's expansion: acquire)
        tcgen05_fence_after();
        if (a.trace && blockIdx.x == 0 && w < kTraceItems) a.trace[w * 16 + 12] = clock64();
        const int c_valid = min(kChunk, a.cmid_p - j * kChunk);
        const int kpc = (c_valid + 63) >> 6;
        uint32_t da = desc_lo(s_a2 + (uint32_t)slot * a.a2_stride) | (((a.a2_lbo >> 4) & 0x3fffu) << 16);
        for (int kb2 = 0; kb2 < kpc; ++kb2) {
          mbar_wait(&bars[kBarW2Full + slot2], (uint32_t)round2 & 1u);
          tcgen05_fence_after();
          const int ksteps = min(4, (c_valid - kb2 * 64 + 15) >> 4);
          const uint32_t db = desc_lo(s_w2 + (uint32_t)slot2 * a.w2_stage_bytes);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks, da += 2048u >> 4)   // A2: 16 channels = two 8-channel K groups of 1024 B
            if (ks < ksteps) umma_f16_lo(tmem_d2, da, db + (uint32_t)(ks * 2), kDescHi, idesc2, (j | kb2 | ks) ? 1u : 0u);
          if (a.residual && (j | kb2) == 0) {
            // D2 was just initialised by the tile's first MMA: add the block input (centre pixel of every output = X
            // tile rows shifted by one halo row) through a 16 x 16 identity, 16 channels per MMA -- exact in fp32
            const uint32_t did = desc_lo(sm + a.off_id);
            for (int kb = 0; kb < a.kb_in; ++kb) {
              const int rsteps = min(64, a.cin_p - kb * 64) >> 4;
              const uint32_t dxa = desc_lo(s_x + (uint32_t)kb * a.x_kb_bytes + (uint32_t)a.W * 128u);
              for (int rs = 0; rs < rsteps; ++rs)
                umma_f16_lo(tmem_d2 + (uint32_t)(kb * 64 + rs * 16), dxa + (uint32_t)(rs * 2), did, kDescHi, idesc_res, 1u);
            }
          }
          umma_commit(&bars[kBarW2Empty + slot2]);
          if (++slot2 == a.s2) { slot2 = 0; ++round2; }
        }
        umma_commit(&bars[kBarMma2 + slot]);
        if (a.trace && blockIdx.x == 0 && w < kTraceItems) a.trace[w * 16 + 13] = clock64();
        if (++j == a.n_chunks) { j = 0; ++ti; }
      }
    }
  } else if (warp == kStoreWarp) {
    // =========================== output store ===========================
    if (a.stage_out && elect_one_sync()) {
      for (int ti = 0; ti < n_my_tiles; ++ti) {
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        const int b = tile / a.tiles_per_window;
        const int ho0 = (tile - b * a.tiles_per_window) * a.TH;
        const int valid_px = min(a.M2, (a.Ho - ho0) * a.Wo);
        const int w_last = (ti + 1) * a.n_chunks - 1;                      // the tile's last item: its A2 buffer holds the staged tile
        mbar_wait(&bars[kBarStaged], (uint32_t)ti & 1u);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                         a.Y + ((int64_t)b * a.Ho + ho0) * a.Wo * a.cout_p),
                     "r"(s_a2 + (uint32_t)(w_last & 1) * a.a2_stride), "r"((uint32_t)valid_px * (uint32_t)(a.cout_p * 2))
                     : "memory");
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&bars[kBarStage]);
      }
      tma_store_wait_all();   // the last copy must complete before the CTA's shared memory goes away
    }
  } else {
    // =========================== compute warps ===========================
    constexpr int kIn = (kRows - 1) * kStride + 3;       // halo rows feeding this warp's kRows output rows
    constexpr int kOut = kSeg / kStride;                 // outputs per row of this warp's segment
    constexpr int kOP = kOut / 2;                        // ... as fp16 pairs
    const int q = warp & 3, sub = warp >> 2;
    const int cl = q * 32 + lane;                        // channel inside the chunk == TMEM lane
    const int rg = sub / a.csub, cs = sub - rg * a.csub;
    const int oh0 = rg * kRows;                          // first output row (tile-local) of this warp
    const int ihl0 = oh0 * kStride;                      // first halo row it reads
    const bool has_left = cs > 0;
    const bool has_right = (kStride == 1) && (cs + 1 < a.csub);
    const uint32_t d1_off = ((uint32_t)(q * 32) << 16) + (uint32_t)(ihl0 * a.W + cs * kSeg);
    // A2 byte offsets of this thread's 8-pixel chunks: m = (oh0 + r) * Wo + cs * kOut + 8 c
    const uint32_t a2_row = (uint32_t)(cl >> 3) * 1024u + (uint32_t)(cl & 7) * 128u;
    auto a2_off = [&](int r, int c8) {
      const uint32_t m = (uint32_t)((oh0 + r) * a.Wo + cs * kOut + 8 * c8);
      return (m >> 6) * a.a2_lbo + a2_row + ((((m >> 3) & 7u) ^ ((uint32_t)cl & 7u)) << 4);
    };
    // relu6(x) = 6 sat(x / 6): one saturating HFMA2 instead of add + max + min.  k6 is 1/6 rounded to fp16 and every
    // depthwise tap is pre-multiplied by its exact reciprocal (pack_consts), so the only deviation is
    // the clamp sitting at 6 (1 +- 2.4e-4), below the fp16 resolution at 6.
    const __half k6h = __float2half_rn(1.f / 6.f);
    const __half2 k62 = __half2half2(k6h), zero2 = __floats2half2_rn(0.f, 0.f), six2 = __floats2half2_rn(6.f, 6.f);

    // per-channel constants of chunk j (pack_consts): fp16 pairs (w0,w1) (w2,w3) (w4,w5) (w6,w7) (w8,bd) (b1,0), the
    // taps pre-divided and b1 pre-multiplied by k6
    uint4 cpa;
    uint2 cpb;
    auto load_consts = [&](int j) {
      const uint4* src = a.cpack + (size_t)(j * kChunk + cl) * 2;
      cpa = __ldg(src);
      cpb = __ldg(reinterpret_cast<const uint2*>(src + 1));
    };
    if (n_items > 0) load_consts(0);

    int ti = 0, j = 0, b = 0, ho0 = 0;
    // ---- epilogue 2: D2 (+ residual, added by the MMA warp) -> +b2 -> bf16 -> Y.
    // DEFERRED: the compute warps run it after the depthwise taps of the NEXT item (first chunk of the next tile), so the
    // last projection MMA of the tile (8 MN-major MMAs at ~140 cycles + the slowest warp's arrival, ~2000 cycles)
    // retires under useful work instead of being waited for.  The MMA2 thread does not touch D2 again before `tile`.
    // The tile's outputs are one contiguous block of Y ([M2 pixels][cout_p]): they are staged in that layout in the A2
    // buffer of the tile's last chunk (its MMA retired; the buffer is next written two items later) and leave with ONE
    // bulk copy issued by the store warp once all 16 warps arrived -- nobody waits for anybody here.  (A thread owns a
    // pixel, so direct stores are 16-byte pieces 2 cout_p bytes apart: ~8000 cycles per tile.)
    bool epi_pending = false;
    int ep_b = 0, ep_ho0 = 0, ep_slot = 0, ep_use2 = 0;
    auto do_epilogue = [&](int w) {
      (void)w;
      mbar_wait_relaxed(&bars[kBarMma2 + ep_slot], (uint32_t)ep_use2 & 1u);
      tcgen05_fence_after();
      AMT_TRACE(6);
      const int o = q * 32 + lane;   // output pixel of this thread's TMEM lane
      const int valid_px = min(a.M2, (a.Ho - ep_ho0) * a.Wo);
      const bool valid = o < valid_px;
      __nv_bfloat16* ytile = a.Y + ((int64_t)ep_b * a.Ho + ep_ho0) * a.Wo * a.cout_p;
      const uint32_t st_row = s_a2 + (uint32_t)ep_slot * a.a2_stride + (uint32_t)o * (uint32_t)(a.cout_p * 2);
      if (q * 32 < a.M2) {   // warp-uniform: quadrants beyond the tile's pixels have nothing to write
        for (int c0 = sub * 16; c0 < a.cout_p; c0 += 64) {
          uint32_t v[16];
          tmem_ld_x16(tmem_d2 + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const float4 ba = *reinterpret_cast<const float4*>(s_b2 + c0 + g * 8), bb = *reinterpret_cast<const float4*>(s_b2 + c0 + g * 8 + 4);
              float f[8];
              f[0] = __uint_as_float(v[g * 8 + 0]) + ba.x; f[1] = __uint_as_float(v[g * 8 + 1]) + ba.y;
              f[2] = __uint_as_float(v[g * 8 + 2]) + ba.z; f[3] = __uint_as_float(v[g * 8 + 3]) + ba.w;
              f[4] = __uint_as_float(v[g * 8 + 4]) + bb.x; f[5] = __uint_as_float(v[g * 8 + 5]) + bb.y;
              f[6] = __uint_as_float(v[g * 8 + 6]) + bb.z; f[7] = __uint_as_float(v[g * 8 + 7]) + bb.w;
              const uint32_t p0 = pack2(f[0], f[1]), p1 = pack2(f[2], f[3]), p2 = pack2(f[4], f[5]), p3 = pack2(f[6], f[7]);
              if (a.stage_out) {
                sts128(st_row + (uint32_t)((c0 + g * 8) * 2), p0, p1, p2, p3);
              } else {
                *reinterpret_cast<uint4*>(ytile + (int64_t)o * a.cout_p + c0 + g * 8) = make_uint4(p0, p1, p2, p3);
              }
            }
          }
        }
      }
      AMT_TRACE(7);
      tcgen05_fence_before();
      if (a.stage_out) fence_proxy_async();           // staged bytes -> visible to the bulk copy (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[kBarTile]);                 // D2 drained
        if (a.stage_out) mbar_arrive(&bars[kBarStaged]);
      }
      epi_pending = false;
    };

    uint32_t row_mask = 0;   // bit r: halo row ihl0 + r of this tile is a real image row
    for (int w = 0; w < n_items; ++w) {
      const bool first = (j == 0), last = (j == a.n_chunks - 1);
      if (first) {
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        b = tile / a.tiles_per_window;
        ho0 = (tile - b * a.tiles_per_window) * a.TH;
        const int gh0 = ho0 * kStride - 1 + ihl0;
        row_mask = 0;
#pragma unroll
        for (int r = 0; r < kIn; ++r)
          if ((uint32_t)(gh0 + r) < (uint32_t)a.H) row_mask |= 1u << r;
      }
      __half2 wk[9];
      {
        const __half2 p0 = u2h(cpa.x), p1 = u2h(cpa.y), p2 = u2h(cpa.z), p3 = u2h(cpa.w), p4 = u2h(cpb.x);
        wk[0] = __low2half2(p0); wk[1] = __high2half2(p0); wk[2] = __low2half2(p1); wk[3] = __high2half2(p1);
        wk[4] = __low2half2(p2); wk[5] = __high2half2(p2); wk[6] = __low2half2(p3); wk[7] = __high2half2(p3);
        wk[8] = __low2half2(p4);
      }
      const __half2 bd2 = __high2half2(u2h(cpb.x)), b12 = __low2half2(u2h(cpb.y));
      if (w + 1 < n_items) load_consts(last ? 0 : j + 1);   // next item's constants land under this item's math

      const int ds = (a.d1_bufs == 2) ? (w & 1) : 0;
      const int slot = (a.a2_bufs == 2) ? (w & 1) : 0;
      const int use2 = (a.a2_bufs == 2) ? (w >> 1) : w;
      const bool live = (j * kChunk + q * 32) < a.cmid_p;   // warp-uniform: this quadrant has channels in a ragged last chunk
      AMT_TRACE(0);
      mbar_wait_relaxed(&bars[kBarMma1 + ds], (uint32_t)((a.d1_bufs == 2) ? (w >> 1) : w) & 1u);
      tcgen05_fence_after();
      AMT_TRACE(1);

      __half2 acc[kRows][kOP];
#pragma unroll
      for (int r = 0; r < kRows; ++r)
#pragma unroll
        for (int i = 0; i < kOP; ++i) acc[r][i] = bd2;

      if (live) {
        const uint32_t t_row0 = tmem_base + (uint32_t)(ds * a.M1) + d1_off;
        auto act2 = [&](uint32_t x0, uint32_t x1) {   // relu6(fp16(x) + b1) / 6 on a pair
          return __hfma2_sat(__floats2half2_rn(__uint_as_float(x0), __uint_as_float(x1)), k62, b12);
        };
        uint32_t raw[kSeg], xl = 0u, xr = 0u;
        auto issue_row = [&](int r) {
          const uint32_t taddr = t_row0 + (uint32_t)(r * a.W);
          if constexpr (kSeg == 32) {
            tmem_ld_x32(taddr, raw);
          } else {
            tmem_ld_x16(taddr, raw);
          }
          if (has_left) tmem_ld_x1(taddr - 1u, xl);
          if (has_right) tmem_ld_x1(taddr + (uint32_t)kSeg, xr);
        };
        if (row_mask & 1u) issue_row(0);
#pragma unroll
        for (int r = 0; r < kIn; ++r) {
          const bool next_live = (r + 1 < kIn) && ((row_mask >> (r + 1)) & 1u);
          if (!((row_mask >> r) & 1u)) {   // zero padding row above / below the image (warp-uniform)
            if (next_live) issue_row(r + 1);
            continue;
          }
          tmem_ld_wait();
          const __half2 eL = has_left ? act2(xl, xl) : zero2;    // pixel left of the segment (0 = image edge)
          if constexpr (kStride == 1) {
            const __half2 eR = has_right ? act2(xr, xr) : zero2;
            __half2 h[kSeg / 2], sft[kSeg / 2 + 1];
#pragma unroll
            for (int i = 0; i < kSeg / 2; ++i) h[i] = act2(raw[2 * i], raw[2 * i + 1]);
            sft[0] = shift_pair(eL, h[0]);
#pragma unroll
            for (int i = 1; i < kSeg / 2; ++i) sft[i] = shift_pair(h[i - 1], h[i]);
            sft[kSeg / 2] = shift_pair(h[kSeg / 2 - 1], eR);
            constexpr bool kEarly1 = (kSeg < 32) || (AM_FUSEDT_PREFETCH32_S1 != 0);
            if (kEarly1 && next_live) issue_row(r + 1);   // raw / xl / xr are dead: the next row's TMEM read overlaps the FMAs below
#pragma unroll
            for (int ro = 0; ro < kRows; ++ro) {
              const int kr = r - ro;
              if (kr < 0 || kr > 2) continue;
#pragma unroll
              for (int i = 0; i < kOP; ++i) {
                acc[ro][i] = __hfma2(sft[i], wk[kr * 3 + 0], acc[ro][i]);
                acc[ro][i] = __hfma2(h[i], wk[kr * 3 + 1], acc[ro][i]);
                acc[ro][i] = __hfma2(sft[i + 1], wk[kr * 3 + 2], acc[ro][i]);
              }
            }
            if (!kEarly1 && next_live) issue_row(r + 1);
          } else {
            // stride 2: output pair (2k, 2k+1) <- centre taps (4k, 4k+2), right taps (4k+1, 4k+3), left taps (4k-1, 4k+1)
            __half2 ee[kSeg / 4], eo[kSeg / 4], lf[kSeg / 4];
#pragma unroll
            for (int k = 0; k < kSeg / 4; ++k) {
              ee[k] = act2(raw[4 * k], raw[4 * k + 2]);
              eo[k] = act2(raw[4 * k + 1], raw[4 * k + 3]);
            }
            lf[0] = shift_pair(eL, eo[0]);
#pragma unroll
            for (int k = 1; k < kSeg / 4; ++k) lf[k] = shift_pair(eo[k - 1], eo[k]);
            constexpr bool kEarly2 = (kSeg < 32) || (AM_FUSEDT_PREFETCH32_S2 != 0);
            if (kEarly2 && next_live) issue_row(r + 1);
#pragma unroll
            for (int ro = 0; ro < kRows; ++ro) {
              const int kr = r - 2 * ro;
              if (kr < 0 || kr > 2) continue;
#pragma unroll
              for (int k = 0; k < kOP; ++k) {
                acc[ro][k] = __hfma2(lf[k], wk[kr * 3 + 0], acc[ro][k]);
                acc[ro][k] = __hfma2(ee[k], wk[kr * 3 + 1], acc[ro][k]);
                acc[ro][k] = __hfma2(eo[k], wk[kr * 3 + 2], acc[ro][k]);
              }
            }
            if (!kEarly2 && next_live) issue_row(r + 1);
          }
        }
      }
      AMT_TRACE(2);
      // D1T(w) fully read
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kBarD1Free + ds]);

      // ---- relu6 -> A2 (MN-major: 8 consecutive pixels of this channel per 16-byte store)
      if (use2 > 0) mbar_wait_relaxed(&bars[kBarMma2 + slot], (uint32_t)(use2 - 1) & 1u);   // the slot's previous MMA2 retired
      if (epi_pending) do_epilogue(w);
      // this buffer staged the previous tile's output (written during the previous item): its bulk copy must have read it
      if (a.stage_out && ti > 0 && j == 1) mbar_wait_relaxed(&bars[kBarStage], (uint32_t)(ti - 1) & 1u);
      AMT_TRACE(3);
      if (live) {
        const uint32_t a2_dst = s_a2 + (uint32_t)slot * a.a2_stride;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
#pragma unroll
          for (int c8 = 0; c8 < kOut / 8; ++c8) {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = h2u(__hmin2(__hmax2(acc[r][c8 * 4 + e], zero2), six2));   // relu6 on the ALU pipe
            sts128(a2_dst + a2_off(r, c8), o[0], o[1], o[2], o[3]);
          }
        }
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[kBarA2Full + slot]);
      AMT_TRACE(4);

      if (last) {   // epilogue 2 of this tile is deferred: it runs after the depthwise taps of the next item
        epi_pending = true;
        ep_b = b;
        ep_ho0 = ho0;
        ep_slot = slot;
        ep_use2 = use2;
      }
      AMT_TRACE(5);
      if (last) { j = 0; ++ti; } else ++j;
    }
    if (epi_pending) do_epilogue(n_items);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// per-channel constants in the form the compute warps consume: out[c] = 8 x u32 =
//   h2(w0, w1) h2(w2, w3) h2(w4, w5) h2(w6, w7) h2(w8, bd) h2(b1 k6, 0) 0 0      with w_t = wd[t][c] / k6
__global__ void pack_consts_kernel(const float* __restrict__ wd, const float* __restrict__ bd, const float* __restrict__ b1,
                                   int cmid_p, int cmid128, uint32_t* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cmid128) return;
  const float k6 = __half2float(__float2half_rn(1.f / 6.f)), inv = 1.f / k6;
  float v[12];
#pragma unroll
  for (int t = 0; t < 9; ++t) v[t] = c < cmid_p ? wd[(size_t)t * cmid_p + c] * inv : 0.f;
  v[9] = c < cmid_p ? bd[c] : 0.f;
  v[10] = c < cmid_p ? b1[c] * k6 : 0.f;
  v[11] = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) out[(size_t)c * 8 + i] = h2u(__floats2half2_rn(v[2 * i], v[2 * i + 1]));
  out[(size_t)c * 8 + 6] = 0u;
  out[(size_t)c * 8 + 7] = 0u;
}

size_t consts_words(int cmid_p) { return round_up((size_t)cmid_p, kChunk) * 8; }

int pack_consts(const float* wd, const float* bd, const float* b1, int cmid_p, uint32_t* out, cudaStream_t st) {
  const int cmid128 = (int)round_up((size_t)cmid_p, kChunk);
  AM_LAUNCH(pack_consts_kernel, (cmid128 + 127) / 128, 128, 0, st, wd, bd, b1, cmid_p, cmid128, out);
  return AM_OK;
}

// ---------------------------------------------------------------- host
static size_t layout_smem(Args& a) {
  size_t off = 0;
  a.x_kb_bytes = (uint32_t)round_up((size_t)a.M1 * 128u, 1024);
  a.off_x = (uint32_t)off;
  off += (size_t)a.kb_in * a.x_kb_bytes;
  a.off_a2 = (uint32_t)off;
  a.a2_stride = (uint32_t)std::max<size_t>(a.a2_bytes, round_up((size_t)a.M2 * a.cout_p * 2, 1024));
  off += (size_t)a.a2_bufs * a.a2_stride;
  a.off_w1 = (uint32_t)off;
  off += (size_t)a.s1 * kW1Stage;
  a.off_w2 = (uint32_t)off;
  a.w2_stage_bytes = (uint32_t)round_up((size_t)a.cout_p * 128u, 1024);
  off += (size_t)a.s2 * a.w2_stage_bytes;
  a.off_id = (uint32_t)off;
  if (a.residual) off += 2048;
  a.off_small = (uint32_t)off;
  off += round_up((size_t)a.cout_p * 4, 16);
  a.off_bar = (uint32_t)off;
  off += (size_t)(kBarCount + 2) * 8;
  return off + 1024;  // alignment slack
}

static bool fill_geometry(const fused::BlockDesc& d, int TH, Args* a) {
  const int Ho = (d.H + 2 - 3) / d.stride + 1, Wo = (d.W + 2 - 3) / d.stride + 1;
  int csub, rows;
  if (TH % 4 == 0) { csub = 1; rows = TH / 4; }
  else if (TH == 2) { csub = 2; rows = 1; }
  else if (TH == 1) { csub = 4; rows = 1; }
  else return false;
  const int seg = d.W / csub;
  if (rows > 2 || (seg != 16 && seg != 32) || (d.stride == 2 && rows != 1)) return false;
  a->H = d.H; a->W = d.W; a->Ho = Ho; a->Wo = Wo;
  a->cin_p = d.cin_p; a->cmid_p = d.cmid_p; a->cout_p = d.cout_p;
  a->residual = d.residual;
  a->TH = TH;
  a->IH = (TH - 1) * d.stride + 3;
  a->M1 = a->IH * d.W;
  a->M2 = TH * Wo;
  if (a->M2 > 128 || a->M2 % 8) return false;
  a->n_parts = a->M1 > 256 ? 2 : 1;
  if (a->M1 % (16 * a->n_parts)) return false;
  a->n_part = a->M1 / a->n_parts;
  if (a->n_part > 256) return false;
  a->kb_in = (d.cin_p + 63) / 64;
  a->n_chunks = (d.cmid_p + kChunk - 1) / kChunk;
  a->csub = csub;
  a->d1_bufs = (2 * a->M1 + d.cout_p <= kTmemCols) ? 2 : 1;
  if (a->d1_bufs * a->M1 + d.cout_p > kTmemCols) return false;
  a->a2_bytes = a->M2 > 64 ? 32768u : 16384u;
  a->a2_lbo = a->M2 > 64 ? 16384u : 0u;
  return true;
}

bool plan(const fused::BlockDesc& d, Plan* out) {
  static const bool off = std::getenv("AM_FUSED_V1") != nullptr;   // force the pixel-per-lane kernel (fused_block.cu)
  if (off || !d.has_expand) return false;
  if (d.cout_p > 256 || d.cout_p % 16 || d.cin_p % 16 || d.cmid_p % 16 || d.cin_p > 512) return false;
  if ((d.W != 16 && d.W != 32 && d.W != 64) || (d.stride != 1 && d.stride != 2)) return false;
  if (d.residual && (d.stride != 1 || d.cin_p != d.cout_p)) return false;
  const int Ho = (d.H + 2 - 3) / d.stride + 1, Wo = (d.W + 2 - 3) / d.stride + 1;
  constexpr size_t kSmemLimit = 232448 - 512;  // sm_100 opt-in maximum per CTA
  // Two D1 sets let the expansion MMA of chunk w+1 run under the depthwise of chunk w.  A block whose largest tile
  // leaves room for one set only (stride 2: 2 TH + 1 halo rows) could take a smaller tile instead (pass 0,
  // AM_FUSEDT_SMALL_TILES); measured on B200 the per-tile costs outweigh the overlap, so the largest tile is the default.
  static const bool small_tiles = std::getenv("AM_FUSEDT_SMALL_TILES") != nullptr;
  for (int pass = small_tiles ? 0 : 1; pass < 2; ++pass)
  for (int TH = std::min(Ho, 128 / std::max(Wo, 1)); TH >= 1; --TH) {
    Args a{};
    if (!fill_geometry(d, TH, &a)) continue;
    if (pass == 0 && a.d1_bufs != 2) continue;
    a.a2_bufs = 2;
    a.s1 = std::min(kMaxStages, 2 * a.kb_in);
    a.s2 = 4;
    for (;;) {
      if (layout_smem(a) <= kSmemLimit) break;
      // two A2 buffers matter most (depthwise of chunk w+1 under the projection of chunk w, and room to stage the
      // output tile); the expansion MMA has a whole chunk of slack, so its W1 ring may shrink to one chunk
      if (a.s1 > a.kb_in + 1) --a.s1;
      else if (a.s2 > 2) --a.s2;
      else if (a.s1 > a.kb_in) --a.s1;
      else if (a.a2_bufs == 2) a.a2_bufs = 1;
      else if (a.s1 > 2) --a.s1;
      else { a.s1 = 0; break; }
    }
    if (a.s1 == 0) continue;
    out->TH = TH;
    out->a2_bufs = a.a2_bufs;
    out->d1_bufs = a.d1_bufs;
    out->s1 = a.s1;
    out->s2 = a.s2;
    out->smem_bytes = layout_smem(a);
    return true;
  }
  return false;
}

int run(const fused::BlockDesc& d, const Plan& p, const __nv_bfloat16* X, const __nv_bfloat16* W1, const uint32_t* cpack,
        const __half* W2, const float* b2, __nv_bfloat16* Y, int B, cudaStream_t st) {
  Args a{};
  AM_CHECK(fill_geometry(d, p.TH, &a), "fused block (channel-per-lane): plan / launch geometry mismatch");
  a.B = B;
  a.a2_bufs = p.a2_bufs;
  a.s1 = p.s1;
  a.s2 = p.s2;
  a.tiles_per_window = (a.Ho + p.TH - 1) / p.TH;
  a.total_tiles = a.tiles_per_window * B;
  a.cpack = reinterpret_cast<const uint4*>(cpack);
  a.b2 = b2;
  a.Y = Y;
  a.Xg = X;
  const size_t smem = layout_smem(a);
  AM_CHECK(smem == p.smem_bytes && a.d1_bufs == p.d1_bufs, "fused block (channel-per-lane): plan / launch smem mismatch");
  static const bool no_stage = std::getenv("AM_FUSEDT_NO_STAGE") != nullptr;
  a.stage_out = (!no_stage && a.a2_bufs == 2 && a.n_chunks >= 2) ? 1 : 0;

  CUtensorMap mx, mw1, mw2;
  {
    const uint64_t dims[4] = {(uint64_t)d.cin_p, (uint64_t)d.W, (uint64_t)d.H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)d.cin_p * 2, (uint64_t)d.W * d.cin_p * 2, (uint64_t)d.H * d.W * d.cin_p * 2};
    const uint32_t box[4] = {64, (uint32_t)d.W, (uint32_t)a.IH, 1};
    AM_TRY(gemm::encode_map_bf16(&mx, X, 4, dims, str, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)d.cin_p, (uint64_t)d.cmid_p};
    const uint64_t str[1] = {(uint64_t)d.cin_p * 2};
    const uint32_t box[2] = {64, (uint32_t)kChunk};
    AM_TRY(gemm::encode_map_bf16(&mw1, W1, 2, dims, str, box));
  }
  {
    const uint64_t dims[2] = {(uint64_t)d.cmid_p, (uint64_t)d.cout_p};
    const uint64_t str[1] = {(uint64_t)d.cmid_p * 2};
    const uint32_t box[2] = {64, (uint32_t)d.cout_p};
    // fp16 data through a 16-bit tiled map: TMA only moves the bytes (zero OOB fill is format-agnostic)
    AM_TRY(gemm::encode_map_bf16(&mw2, W2, 2, dims, str, box));
  }
  using KernelFn = void (*)(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap,
                            const __grid_constant__ CUtensorMap, const Args);
  const int seg = d.W / a.csub, rows = (p.TH % 4 == 0) ? p.TH / 4 : 1;
  KernelFn fn = nullptr;
  int variant = -1;
  if (d.stride == 1 && seg == 32 && rows == 1) { fn = fused_block_t_kernel<1, 32, 1>; variant = 0; }
  else if (d.stride == 1 && seg == 32 && rows == 2) { fn = fused_block_t_kernel<1, 32, 2>; variant = 1; }
  else if (d.stride == 1 && seg == 16 && rows == 1) { fn = fused_block_t_kernel<1, 16, 1>; variant = 2; }
  else if (d.stride == 1 && seg == 16 && rows == 2) { fn = fused_block_t_kernel<1, 16, 2>; variant = 3; }
  else if (d.stride == 2 && seg == 32 && rows == 1) { fn = fused_block_t_kernel<2, 32, 1>; variant = 4; }
  else if (d.stride == 2 && seg == 16 && rows == 1) { fn = fused_block_t_kernel<2, 16, 1>; variant = 5; }
  AM_CHECK(fn != nullptr, "fused block (channel-per-lane): no kernel variant for stride=%d segment=%d rows=%d", d.stride, seg, rows);
  static size_t attr[6] = {0, 0, 0, 0, 0, 0};
  if (smem > attr[variant]) {
    AM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[variant] = smem;
  }
  const int grid = std::max(1, std::min(a.total_tiles, sm_count()));
  static const bool trace_on = std::getenv("AM_FUSED_TRACE") != nullptr;
  DevBuf<long long> tr;
  if (trace_on) {
    AM_TRY(tr.alloc((size_t)kTraceItems * 16));
    AM_CUDA(cudaMemsetAsync(tr.p, 0, (size_t)kTraceItems * 16 * 8, st));
    a.trace = tr.p;
  }
  {
    auto fused_block_t_kernel = fn;  // (keeps the profiler's kernel name)
    AM_LAUNCH(fused_block_t_kernel, grid, kThreads, smem, st, mx, mw1, mw2, a);
  }
  if (trace_on) {
    std::vector<long long> h((size_t)kTraceItems * 16);
    AM_CUDA(cudaStreamSynchronize(st));
    AM_CUDA(cudaMemcpy(h.data(), tr.p, h.size() * 8, cudaMemcpyDeviceToHost));
    const int n = std::min(kTraceItems, ((a.total_tiles - 1) / grid + 1) * a.n_chunks);
    std::fprintf(stderr, "[fused-t trace] H=%d W=%d cin=%d cmid=%d cout=%d s=%d TH=%d M1=%d parts=%d chunks=%d d1=%d a2=%d s1=%d s2=%d tiles/CTA=%d\n",
                 a.H, a.W, a.cin_p, a.cmid_p, a.cout_p, d.stride, a.TH, a.M1, a.n_parts, a.n_chunks, a.d1_bufs, a.a2_bufs, a.s1,
                 a.s2, (a.total_tiles - 1) / grid + 1);
    const long long t0 = h[0];
    for (int w = a.n_chunks; w < std::min(n, 3 * a.n_chunks); ++w) {
      const long long* e = &h[(size_t)w * 16];
      std::fprintf(stderr,
                   "  w=%2d @%7lld | compute: waitMMA1 %5lld  taps %5lld  waitA2slot %5lld  store %5lld  epi2 %5lld (wait %5lld, deferred epilogue: d2 loop %5lld, fence+arrive %5lld)"
                   " | mma1 @%7lld wait %5lld issue %5lld | mma2 @%7lld wait %5lld issue %5lld\n",
                   w, e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4], e[6] ? e[6] - e[4] : 0, e[7] ? e[7] - e[6] : 0, e[7] ? e[3] - e[7] : 0, e[8] - t0, e[9] - e[8],
                   e[10] - e[9], e[11] - t0, e[12] - e[11], e[13] - e[12]);
    }
  }
  return AM_OK;
}

}  // namespace fusedt
}  // namespace am
