// K1': GRPO advantages.
// Reference: calculate_scores, rlinf/algorithms/utils.py:134-152 (reverse accumulation, CPU only) and
// compute_grpo_advantages, rlinf/algorithms/advantages.py:89-121.
#include "common.cuh"

namespace {

// scores[b] = sum of rewards of the first episode, accumulated in the reference's reverse order:
//   s = (s * !done[t+1]) + r[t]   for t = T-1 .. 0     (two fp32 roundings per step, bit-exact)
__global__ void __launch_bounds__(32) grpo_scores_kernel(const float* __restrict__ rewards,
                                                         const uint8_t* __restrict__ dones,
                                                         float* __restrict__ scores, int T, int B) {
  constexpr int U = 16;
  const int col = blockIdx.x * 32 + threadIdx.x;
  if (col >= B) return;
  float s = 0.0f;
  for (int t_hi = T - 1; t_hi >= 0; t_hi -= U) {
    float r[U];
    uint8_t d[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t_hi - u;
      const bool ok = t >= 0;
      const size_t o = (size_t)(ok ? t : 0) * B + col;
      r[u] = ok ? rewards[o] : 0.0f;
      d[u] = ok ? dones[o + B] : (uint8_t)0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (t_hi - u >= 0) s = __fadd_rn(__fmul_rn(s, d[u] ? 0.0f : 1.0f), r[u]);
    }
  }
  scores[col] = s;
}

// a[b] = (score[b] - mean_group) / (std_group_unbiased + eps), then
// adv[t,b] = (0 + a[b]) * mask[t,b]   (advantages.py:107-119: zeros_like(bool mask) + a, times mask).
// One CTA per 32 envs x all T: the per-env normalised score is computed once into shared memory,
// then 8 warps stride over time with coalesced 128-byte row segments (1 B read + 4 B written per step).
constexpr int kCols = 32;
constexpr int kRowsPar = 8;

__global__ void __launch_bounds__(kCols* kRowsPar) grpo_adv_kernel(const float* __restrict__ scores,
                                                                    const uint8_t* __restrict__ mask,
                                                                    float* __restrict__ adv, int T, int B, int G,
                                                                    float eps) {
  __shared__ float a_sh[kCols];
  const int tx = threadIdx.x & (kCols - 1), ty = threadIdx.x / kCols;
  const int b = blockIdx.x * kCols + tx;
  if (ty == 0 && b < B) {
    const int g0 = (b / G) * G;
    double sum = 0.0;
    for (int k = 0; k < G; ++k) sum += (double)scores[g0 + k];
    const double mean_d = sum / (double)G;
    double m2 = 0.0;
    for (int k = 0; k < G; ++k) {
      const double dlt = (double)scores[g0 + k] - mean_d;
      m2 += dlt * dlt;
    }
    const float mean = (float)mean_d;
    const float stdv = (float)sqrt(m2 / (double)(G - 1));  // G==1 -> NaN like torch.std
    a_sh[tx] = __fadd_rn(0.0f, __fdiv_rn(__fsub_rn(scores[b], mean), __fadd_rn(stdv, eps)));
  }
  __syncthreads();
  if (b >= B) return;
  const float av = a_sh[tx];
  for (int t = ty; t < T; t += kRowsPar) {
    const size_t o = (size_t)t * B + b;
    adv[o] = __fmul_rn(av, mask ? (mask[o] ? 1.0f : 0.0f) : 1.0f);
  }
}

}  // namespace

extern "C" int rb200_grpo_scores(const float* rewards, const uint8_t* dones, float* scores, int T, int B,
                                 rb200_stream_t stream) {
  if (!rewards || !dones || !scores) return RB200_E_NULL;
  if (T <= 0 || B <= 0) return RB200_E_SHAPE;
  grpo_scores_kernel<<<(B + 31) / 32, 32, 0, rb::as_stream(stream)>>>(rewards, dones, scores, T, B); rb::count_launch();
  RB_RETURN_LAUNCH();
}

extern "C" int rb200_grpo_advantages(const float* scores, const uint8_t* loss_mask, float* adv, int T, int B, int G,
                                     float eps, rb200_stream_t stream) {
  if (!scores || !adv) return RB200_E_NULL;
  if (T <= 0 || B <= 0 || G <= 0 || B % G != 0) return RB200_E_SHAPE;
  grpo_adv_kernel<<<(B + kCols - 1) / kCols, kCols * kRowsPar, 0, rb::as_stream(stream)>>>(scores, loss_mask, adv, T,
                                                                                           B, G, eps); rb::count_launch();
  RB_RETURN_LAUNCH();
}
