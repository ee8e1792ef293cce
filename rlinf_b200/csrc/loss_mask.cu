// K0: loss mask from done flags (integer/byte work, bit-exact).
// Reference: compute_loss_mask, rlinf/utils/metric_utils.py:516-537.
//
// dones [(nc+1), B, C] bool.  Step-major flat index j in [0, nc*C] maps to full row
// f = (C-1) + j -> (chunk f/C, slot f%C).  first_done[b] = min j with dones set;
// mask step s (0..nc*C-1) is valid iff s < first_done[b]  (cumsum over rows [0..s] == 0).
// mask_sum[b] = min(first_done[b], nc*C).
//
// One CTA per 32 envs, 8 warps striding over time: every byte is read once and written once
// (2 B/step algorithmic), coalesced along the env axis.
#include "common.cuh"

namespace {

constexpr int kCols = 32;
constexpr int kRowsPar = 8;

__global__ void __launch_bounds__(kCols* kRowsPar) loss_mask_kernel(const uint8_t* __restrict__ dones,
                                                                     uint8_t* __restrict__ mask,
                                                                     int64_t* __restrict__ mask_sum, int nc, int B,
                                                                     int C) {
  __shared__ int first_done[kCols];
  const int tx = threadIdx.x & (kCols - 1), ty = threadIdx.x / kCols;
  const int b = blockIdx.x * kCols + tx;
  const int n_steps = nc * C;
  if (ty == 0) first_done[tx] = n_steps + 1;  // sentinel: no done in rows [0..n_steps]
  __syncthreads();
  if (b < B) {
    int local = n_steps + 1;
    for (int j = ty; j <= n_steps; j += kRowsPar) {
      const int f = (C - 1) + j;
      const int chunk = f / C, slot = f - chunk * C;
      if (dones[((size_t)chunk * B + b) * C + slot]) {
        local = j;
        break;  // j only grows: first hit of this thread's stride is its minimum
      }
    }
    if (local <= n_steps) atomicMin(&first_done[tx], local);
  }
  __syncthreads();
  if (b < B) {
    const int fd = first_done[tx];
    for (int s = ty; s < n_steps; s += kRowsPar) {
      const int chunk = s / C, slot = s - chunk * C;
      mask[((size_t)chunk * B + b) * C + slot] = (uint8_t)(s < fd);
    }
    if (ty == 0) mask_sum[b] = (int64_t)(fd < n_steps ? fd : n_steps);
  }
}

}  // namespace

extern "C" int rb200_loss_mask(const uint8_t* dones, uint8_t* mask, int64_t* mask_sum, int nc, int B, int C,
                               rb200_stream_t stream) {
  if (!dones || !mask || !mask_sum) return RB200_E_NULL;
  if (nc <= 0 || B <= 0 || C <= 0) return RB200_E_SHAPE;
  const int grid = (B + kCols - 1) / kCols;
  loss_mask_kernel<<<grid, kCols * kRowsPar, 0, rb::as_stream(stream)>>>(dones, mask, mask_sum, nc, B, C); rb::count_launch();
  RB_RETURN_LAUNCH();
}
