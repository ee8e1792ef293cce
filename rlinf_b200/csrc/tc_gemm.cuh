// Host-side interface of the tcgen05 3xTF32 GEMM (tc_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32;  // BK fp32 = 128 bytes = one swizzle span

enum Epi { EPI_STORE = 0, EPI_BIAS_TANH_SPLIT = 1, EPI_TANHGRAD_SPLIT = 2 };

struct Params {
  int64_t M;
  int K;              // multiple of 32
  const float* bias;  // [256]                         (EPI_BIAS_TANH_SPLIT)
  const float* h_hi;  // [M,256] previous activation   (EPI_TANHGRAD_SPLIT)
  const float* h_lo;
  float* c_hi;        // [M,256] (EPI_STORE: plain fp32 result)
  float* c_lo;        // [M,256]
  float* colsum;      // [256] += column sums of the output (bias gradient of the producing layer), or NULL
  int epi;
};

// C[M,256] = epi( (A_hi+A_lo)[M,K] . (B_hi+B_lo)[256,K]^T ); all operands exact-TF32 fp32 arrays, 16-byte aligned.
int launch(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, const Params& p, cudaStream_t st);
// dW[256, IN] += (Z_hi+Z_lo)[n,256]^T . (H_hi+H_lo)[n,IN]   (IN % 32 == 0, <= 256), fp32 atomics into dW
int wgrad(const float* z_hi, const float* z_lo, const float* h_hi, const float* h_lo, float* dW, int64_t n, int IN,
          cudaStream_t st);
// x -> (hi, lo) exact-TF32 pair, hi + lo ~= x to 2^-21 relative; hi may alias x.
int split(const float* x, float* hi, float* lo, int64_t n, cudaStream_t st);
// out[c][r] = split(x[r][c]) for a small [R,C] matrix
int split_transpose(const float* x, float* hi, float* lo, int R, int C, cudaStream_t st);

}  // namespace tc
}  // namespace rb
