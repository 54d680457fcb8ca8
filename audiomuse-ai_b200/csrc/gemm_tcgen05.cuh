// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   D[m, n] = epi( alpha * sum_k A[m, k] * B[n, k] )        A: [M, K] row-major (K-major)
//                                                            B: [N, K] row-major (K-major)
//   epi(v)  = act(v + bias[n] - col_sub[n]) + residual[m, n] ; D is bf16 or fp32.
//
// Users: the encoder's 1x1 convolutions / linear layers (encoder.cu; A = NHWC activations,
// B = folded conv weights) and the k-NN bf16 filter pass (knn.cu; A = queries, B = library).
#pragma once

#include "common.cuh"

namespace am {
namespace gemm {

struct Epilogue {
  float alpha = 1.0f;
  const float* bias = nullptr;              // [N] added
  const float* col_sub = nullptr;           // [N] subtracted (k-NN euclidean: ||x||^2)
  int act = 0;                              // 0 none, 1 relu6, 2 relu, 3 hardswish (model_spec.cuh ActKind)
  const __nv_bfloat16* residual = nullptr;  // [M, ld_res] added after act
  int64_t ld_res = 0;
  // k-NN: besides D, write the maximum of every 32-column chunk of epi(...) (columns >= N count as -inf):
  // chunk_max[m * ld_cm + n / 32].  The selection kernel finds its threshold and the few chunks that can hold answers
  // from these (1/32 of the score matrix) and then touches only those chunks of D.
  float* chunk_max = nullptr;
  int64_t ld_cm = 0;
};

// true when the device can run the tcgen05 path (sm_100) and the driver exports
// cuTensorMapEncodeTiled
bool available();

// lda/ldb in elements, multiples of 8 (16-byte TMA row pitch).  ldd in elements of D.
// m_fastest: enumerate tiles with the M index fastest (B tile shared by consecutive CTAs;
// right when A is small, e.g. k-NN queries); otherwise N fastest (A tile shared).
int gemm_bf16(const __nv_bfloat16* A, int64_t M, int64_t lda, const __nv_bfloat16* B, int64_t N,
              int64_t ldb, int K, void* D, int64_t ldd, bool d_is_f32, const Epilogue& ep,
              bool m_fastest, cudaStream_t st);

// bf16 tiled tensor map, rank <= 4, SWIZZLE_128B, zero OOB fill.  map_out: 128-byte CUtensorMap.
int encode_map_bf16(void* map_out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

// reference implementation on CUDA cores (slow; used only by the on-device self test)
int gemm_bf16_simt(const __nv_bfloat16* A, int64_t M, int64_t lda, const __nv_bfloat16* B, int64_t N,
                   int64_t ldb, int K, void* D, int64_t ldd, bool d_is_f32, const Epilogue& ep,
                   cudaStream_t st);

// S[q, j] = Qb[q, :] . Xb[j, :]  (2*dot - xnorm2[j] when xnorm2 != NULL), fp32 out.
// Qb has `qrows` rows (multiple of 128, zero padded).
inline int scores_bf16(const __nv_bfloat16* Qb, int qrows, const __nv_bfloat16* Xb, int64_t N, int dpad,
                       float* S, int64_t ldS, const float* xnorm2, cudaStream_t st) {
  Epilogue ep;
  ep.alpha = xnorm2 ? 2.0f : 1.0f;
  ep.col_sub = xnorm2;
  return gemm_bf16(Qb, qrows, dpad, Xb, N, dpad, dpad, S, ldS, true, ep, /*m_fastest=*/true, st);
}

// same scores + the maximum of every 32 consecutive ones: CM[q, j / 32]
inline int scores_chunkmax_bf16(const __nv_bfloat16* Qb, int qrows, const __nv_bfloat16* Xb, int64_t N, int dpad,
                                float* S, int64_t ldS, float* CM, int64_t ldCM, const float* xnorm2, cudaStream_t st) {
  Epilogue ep;
  ep.alpha = xnorm2 ? 2.0f : 1.0f;
  ep.col_sub = xnorm2;
  ep.chunk_max = CM;
  ep.ld_cm = ldCM;
  return gemm_bf16(Qb, qrows, dpad, Xb, N, dpad, dpad, S, ldS, true, ep, /*m_fastest=*/true, st);
}

}  // namespace gemm
}  // namespace am
