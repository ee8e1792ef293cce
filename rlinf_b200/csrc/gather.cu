// a12: trajectory-row gather  dst[i,:] = src[idx[i],:]   (bit-exact index/byte work)
// Reference: process_nested_dict_for_train, rlinf/utils/nested_dict_process.py:272-285
// (`value.reshape(-1, *value.shape[2:])[shuffle_id]`, flat row index t*B + b).
#include "common.cuh"

namespace {

// rows that are multiples of 16 bytes: a group of (row_bytes/16) lanes copies one row with 16-byte accesses
__global__ void __launch_bounds__(256) gather_rows_vec16(const uint4* __restrict__ src,
                                                         const int64_t* __restrict__ idx, uint4* __restrict__ dst,
                                                         int64_t n_out, int64_t n_src, int vec_per_row) {
  const int64_t total = n_out * vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / vec_per_row;
    const int v = (int)(i - row * vec_per_row);
    int64_t s = idx[row];
    if (s < 0) s += n_src;  // torch-style negative index
    dst[i] = src[s * vec_per_row + v];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) gather_rows_small(const T* __restrict__ src, const int64_t* __restrict__ idx,
                                                         T* __restrict__ dst, int64_t n_out, int64_t n_src,
                                                         int elems_per_row) {
  const int64_t total = n_out * elems_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / elems_per_row;
    const int e = (int)(i - row * elems_per_row);
    int64_t s = idx[row];
    if (s < 0) s += n_src;
    dst[i] = src[s * elems_per_row + e];
  }
}

}  // namespace

extern "C" int rb200_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n_rows_out,
                                 int64_t n_rows_src, int64_t row_bytes, rb200_stream_t stream) {
  if (!src || !idx || !dst) return RB200_E_NULL;
  if (n_rows_out <= 0 || n_rows_src <= 0 || row_bytes <= 0) return RB200_E_SHAPE;
  cudaStream_t st = rb::as_stream(stream);
  const int64_t cap = (int64_t)rb::sm_count() * 16;
  auto grid_for = [&](int64_t total) {
    int64_t b = (total + 255) / 256;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
  };
  const uintptr_t a = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
  if (row_bytes % 16 == 0 && (a & 15) == 0) {
    const int vpr = (int)(row_bytes / 16);
    gather_rows_vec16<<<grid_for(n_rows_out * vpr), 256, 0, st>>>(static_cast<const uint4*>(src), idx,
                                                                   static_cast<uint4*>(dst), n_rows_out, n_rows_src,
                                                                   vpr); rb::count_launch();
  } else if (row_bytes % 4 == 0 && (a & 3) == 0) {
    const int epr = (int)(row_bytes / 4);
    gather_rows_small<uint32_t><<<grid_for(n_rows_out * epr), 256, 0, st>>>(
        static_cast<const uint32_t*>(src), idx, static_cast<uint32_t*>(dst), n_rows_out, n_rows_src, epr); rb::count_launch();
  } else {
    gather_rows_small<uint8_t><<<grid_for(n_rows_out * row_bytes), 256, 0, st>>>(
        static_cast<const uint8_t*>(src), idx, static_cast<uint8_t*>(dst), n_rows_out, n_rows_src, (int)row_bytes); rb::count_launch();
  }
  RB_RETURN_LAUNCH();
}
