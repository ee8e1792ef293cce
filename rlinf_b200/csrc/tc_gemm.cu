// K3 tensor-core path: C[M,256] = epilogue( A[M,K] . B[256,K]^T ) on tcgen05 (5th-gen tensor cores), fp32-accurate
// through 3xTF32 error compensation:   a = a_hi + a_lo, b = b_hi + b_lo (each an exact TF32 number)
//     a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi        (dropped term a_lo.b_lo ~ 2^-22 relative)
// Plain TF32/bf16 misses the 1e-4 parity bar of the policy loss.  Activations and gradients live in HBM as plain
// fp32 (one copy): the streamed operand is split into its (hi, lo) pair INSIDE the kernel by transform warps that
// rewrite the TMA-landed shared-memory tile in place (hi) and into a sibling tile (lo) - the split is elementwise,
// so it is oblivious to the swizzled tile layout.  (Round-1 v1 stored every activation pre-split: twice the HBM
// bytes on kernels that were HBM-bound.)  Weights are small and reused by every tile: they stay pre-split.
//
// Structure (one CTA per SM, persistent over 128-row tiles, 320 threads):
//   warp 0   : TMA producer  - cp.async.bulk.tensor (SWIZZLE_128B boxes) of A [128x32] and B_hi/B_lo [256x32]
//                              per k-block into a 2-stage shared-memory ring (96 KB / stage), mbarrier full/empty.
//   warp 1   : MMA issuer    - one elected thread, tcgen05.mma.cta_group::1.kind::tf32 M=128 N=256 K=8, accumulators
//                              in TMEM (2 x 256 columns, double-buffered across tiles), tcgen05.commit -> mbarriers.
//   warps 2-5: transform     - A tile -> A_lo tile (the tensor core ignores the low 13 mantissa bits, so the landed
//                              fp32 tile IS A_hi), fence.proxy.async, arrive on the stage's xf barrier.
//   warps 6-9: epilogue      - tcgen05.ld (32 lanes x 32 columns per warp), bias+tanh or tanh' scaling, staged in a
//                              SWIZZLE_128B shared-memory block and written with one TMA tile store per [32 x 32]
//                              block (per-row global stores cost ~60 issue cycles each, see gae.cu); overlaps the
//                              next tile's MMAs.
// Reference op chains replaced: nn.Linear + tanh of MLPPolicy.backbone / ValueHead.mlp
// (rlinf/models/embodiment/mlp_policy/mlp_policy.py:91-98, modules/value_head.py:37-45) and autograd's dgrad.
#include "common.cuh"
#include "tc_gemm.cuh"
#include "tma.cuh"

namespace rb {
namespace tc {

constexpr int kStages = 2;
constexpr int kATile = BM * BK * 4;          // 16 KB
constexpr int kBTile = BN * BK * 4;          // 32 KB
constexpr int kStageBytes = 2 * kATile + 2 * kBTile;  // A(hi) A_lo B_hi B_lo = 96 KB
constexpr int kXfWarps = 4;                  // transform warps of the forward/dgrad kernel
constexpr int kThreads = 32 * (2 + kXfWarps + 4);  // producer, MMA, transform, 4 epilogue warps = 320
constexpr int kStagingBytes = 4 * 2 * 4096;  // epilogue: 4 warps x 2 buffers x [32 rows x 128 B]
constexpr int kTmemCols = 512;
constexpr uint32_t kTf32Mask = 0xffffe000u;

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major; 1) | SBO>>4 [32,46) = 1024 B between 8-row groups |
// version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format TF32 [7,10),[10,13)=2,
// K-major A and B (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// tanh(x) = sign(x) * (1 - t) / (1 + t), t = exp(-2|x|): ~3e-7 absolute error (MUFU ex2 + MUFU rcp),
// an order of magnitude fewer instructions than libdevice tanhf in the 32k-element-per-tile epilogue.
__device__ __forceinline__ float tanh_fast(float x) {
  const float t = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & kTf32Mask);
  lo = __uint_as_float((__float_as_uint(__fsub_rn(x, hi)) + 0x1000u) & kTf32Mask);  // round-to-nearest TF32 (no bias)
}

struct __align__(16) Barriers {
  uint64_t full[kStages];
  uint64_t xf[kStages];     // transform warps: the stage's streamed operand has been split into (hi, lo)
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  // forward epilogue: the layer bias (broadcast LDS instead of 64 dependent LDGs / tile);
  // dgrad epilogue: per-CTA column sums of the output tiles (bias gradient), flushed once at kernel end
  float bias[BN];
};

__device__ __forceinline__ float4 split_hi4(const float4 x) {
  return make_float4(__uint_as_float(__float_as_uint(x.x) & kTf32Mask), __uint_as_float(__float_as_uint(x.y) & kTf32Mask),
                     __uint_as_float(__float_as_uint(x.z) & kTf32Mask), __uint_as_float(__float_as_uint(x.w) & kTf32Mask));
}
__device__ __forceinline__ float lo1(float x, float hi) {
  return __uint_as_float((__float_as_uint(__fsub_rn(x, hi)) + 0x1000u) & kTf32Mask);  // round-to-nearest TF32
}
// split of `count` float4 of a TMA-landed tile: dst <- lo (same element positions); src <- hi only when mask_hi
// (kind::tf32 reads the top 19 bits, so the unmasked fp32 value already acts as hi - tools/tc_flags_probe.py)
template <int NT>
__device__ __forceinline__ void transform_tile(float4* __restrict__ src, float4* __restrict__ dst, int count, int t,
                                               bool mask_hi) {
#pragma unroll 4
  for (int i = t; i < count; i += NT) {
    const float4 x = src[i];
    const float4 h = split_hi4(x);
    if (mask_hi) src[i] = h;
    dst[i] = make_float4(lo1(x.x, h.x), lo1(x.y, h.y), lo1(x.z, h.z), lo1(x.w, h.w));
  }
}

__global__ void __launch_bounds__(kThreads, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b_hi,
                   const __grid_constant__ CUtensorMap tm_b_lo, const __grid_constant__ CUtensorMap tm_c, Params p,
                   int flags) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; pointer arithmetic (no integer round trip) keeps the shared
  // address space visible to the compiler (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + kStages * kStageBytes;  // 1024-byte aligned: 8 x 4 KB SWIZZLE_128B blocks
  Barriers* bars = reinterpret_cast<Barriers*>(staging + kStagingBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.M + BM - 1) / BM;
  const int n_kb = p.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(&bars->full[s], 1);
      tma::mbar_init(&bars->xf[s], kXfWarps);
      tma::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tma::mbar_init(&bars->tmem_full[b], 1);
      tma::mbar_init(&bars->tmem_empty[b], 4);  // one arrive per epilogue warp
    }
    tma::fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation is warp-collective
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  if (p.epi == EPI_BIAS_TANH)
    for (int i = threadIdx.x; i < BN; i += kThreads) bars->bias[i] = p.bias[i];
  else
    for (int i = threadIdx.x; i < BN; i += kThreads) bars->bias[i] = 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      tma::prefetch_desc(&tm_a);
      tma::prefetch_desc(&tm_b_hi);
      tma::prefetch_desc(&tm_b_lo);
      // flat loop over this CTA's (tile, k-block) sequence; A boxes are pulled into L2 `pf` k-blocks ahead so that the
      // real load sees L2 latency: with only two 96 KB stages the DRAM latency would otherwise sit on the critical
      // path load -> transform -> MMA (measured: 3.2k cycles per k-block against 1.5k of MMA work)
      const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
      const int64_t total = my_tiles * n_kb;
      const int pf = (flags >> 8) ? (flags >> 8) & 0xff : 3;
      for (int64_t j = 0; j < pf && j < total; ++j)
        tma::prefetch_2d(&tm_a, (int)(j % n_kb) * BK, (int)((blockIdx.x + (j / n_kb) * gridDim.x) * BM));
      for (int64_t j = 0; j < total; ++j) {
        const uint32_t it = (uint32_t)j;
        const int kb = (int)(j % n_kb);
        const int m0 = (int)((blockIdx.x + (j / n_kb) * gridDim.x) * BM);
        const int64_t jp = j + pf;
        if (jp < total && pf < 255)
          tma::prefetch_2d(&tm_a, (int)(jp % n_kb) * BK, (int)((blockIdx.x + (jp / n_kb) * gridDim.x) * BM));
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        tma::mbar_wait(&bars->empty[s], ph ^ 1u);
        uint8_t* st = smem + s * kStageBytes;
        tma::mbar_arrive_expect_tx(&bars->full[s], kATile + 2 * kBTile);
        tma::load_2d(st, &tm_a, kb * BK, m0, &bars->full[s]);  // rows >= M are zero-filled
        tma::load_2d(st + 2 * kATile, &tm_b_hi, kb * BK, 0, &bars->full[s]);
        tma::load_2d(st + 2 * kATile + kBTile, &tm_b_lo, kb * BK, 0, &bars->full[s]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (single thread) =================
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1u;
        const uint32_t bph = (tcount >> 1) & 1u;
        tma::mbar_wait(&bars->tmem_empty[buf], bph ^ 1u);  // epilogue has drained this accumulator
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * BN;  // column offset
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          tma::mbar_wait(&bars->full[s], ph);  // B_hi / B_lo have landed
          tma::mbar_wait(&bars->xf[s], ph);    // A has been split (generic-proxy writes fenced by the writers)
          fence_after_sync();
          const uint32_t sa = tma::smem_u32(smem + s * kStageBytes);
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + kATile);
          const uint64_t b_hi = make_desc(sa + 2 * kATile), b_lo = make_desc(sa + 2 * kATile + kBTile);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per UMMA_K=8 step, in 16-byte units
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            mma_tf32(d_tmem, a_lo + koff, b_hi + koff, kIdesc, acc);   // small terms first
            mma_tf32(d_tmem, a_hi + koff, b_lo + koff, kIdesc, 1u);
            mma_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdesc, 1u);
          }
          mma_commit(&bars->empty[s]);  // smem stage free once these MMAs have read it
        }
        mma_commit(&bars->tmem_full[buf]);  // accumulator complete
      }
    }
  } else if (warp < 2 + kXfWarps) {
    // ================= transform warps: A -> (A_hi, A_lo) =================
    const int t = threadIdx.x - 64;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < n_kb; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        tma::mbar_wait(&bars->full[s], ph);
        float4* a = reinterpret_cast<float4*>(smem + s * kStageBytes);
        transform_tile<32 * kXfWarps>(a, a + kATile / 16, kATile / 16, t, (flags & 1) != 0);
        tma::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&bars->xf[s]);
      }
    }
  } else {
    // ================= epilogue warps 6..9 =================
    const int q = warp & 3;               // TMEM lane quarter this warp may access: lanes [32q, 32q+32)
    const int ew = warp - (2 + kXfWarps);  // staging slot
    float4 (*stg)[32][8] = reinterpret_cast<float4 (*)[32][8]>(staging + ew * 8192);  // [2][32 rows][8 float4]
    if (lane == 0) tma::prefetch_desc(&tm_c);
    const int rs = lane >> 3, c4 = lane & 7;  // transposed phase: 4 rows x 8 float4 per instruction
    // EPI_TANHGRAD: the previous activation h of the NEXT [32 x 32] block is fetched into registers while the
    // current block is processed (ncu round 1: with the loads issued per block the dgrad epilogue exposed one DRAM
    // round trip per block and ran at 248 us against 173 us for the forward GEMM of the same shape)
    float4 hp[8];
    auto load_h = [&](int64_t r0, int cc) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t gr = r0 + it * 4 + rs;
        hp[it] = gr < p.M ? __ldg(reinterpret_cast<const float4*>(p.h + gr * BN + cc) + c4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (p.epi == EPI_TANHGRAD && (int64_t)blockIdx.x < n_tiles) load_h((int64_t)blockIdx.x * BM + q * 32, 0);
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1u;
      const uint32_t bph = (tcount >> 1) & 1u;
      tma::mbar_wait(&bars->tmem_full[buf], bph);
      fence_after_sync();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
      const int64_t row0 = tile * BM + q * 32;   // first row of this warp's 32-row slab
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float4 (*sh)[8] = stg[(c0 >> 5) & 1];
        // the TMA store issued from this buffer two chunks ago must have finished reading it
        if (lane == 0) tma::store_wait_read1();
        __syncwarp();
        uint32_t r[32];
        tmem_ld32(taddr0 + c0, r);  // asynchronous until tmem_ld_wait
        if (p.epi == EPI_TANHGRAD) {
          // h block (coalesced 128-byte row segments, 4 rows per instruction) -> shared memory, one row per thread
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rs;
            sh[rr][c4 ^ (rr & 7)] = hp[it];
          }
          __syncwarp();
          if (c0 + 32 < BN) load_h(row0, c0 + 32);
          else if (tile + gridDim.x < n_tiles) load_h((tile + gridDim.x) * BM + q * 32, 0);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 v;
          if (p.epi == EPI_STORE) {
            v = make_float4(__uint_as_float(r[j4 * 4 + 0]), __uint_as_float(r[j4 * 4 + 1]),
                            __uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3]));
          } else if (p.epi == EPI_BIAS_TANH) {
            const float4 b = *reinterpret_cast<const float4*>(&bars->bias[c0 + j4 * 4]);
            v.x = tanh_fast(__uint_as_float(r[j4 * 4 + 0]) + b.x);
            v.y = tanh_fast(__uint_as_float(r[j4 * 4 + 1]) + b.y);
            v.z = tanh_fast(__uint_as_float(r[j4 * 4 + 2]) + b.z);
            v.w = tanh_fast(__uint_as_float(r[j4 * 4 + 3]) + b.w);
          } else {
            const float4 h = sh[lane][j4 ^ (lane & 7)];  // row `lane` is private to thread `lane` in this phase
            v.x = __uint_as_float(r[j4 * 4 + 0]) * (1.0f - h.x * h.x);
            v.y = __uint_as_float(r[j4 * 4 + 1]) * (1.0f - h.y * h.y);
            v.z = __uint_as_float(r[j4 * 4 + 2]) * (1.0f - h.z * h.z);
            v.w = __uint_as_float(r[j4 * 4 + 3]) * (1.0f - h.w * h.w);
          }
          sh[lane][j4 ^ (lane & 7)] = v;  // == SWIZZLE_128B position of (row lane, 16-byte chunk j4)
        }
        tma::fence_proxy_async();  // staged tile -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
          tma::store_2d(&tm_c, &sh[0][0], c0, (int)row0);  // rows >= M are clipped
          tma::store_commit();
        }
        if (p.colsum != nullptr && p.epi == EPI_TANHGRAD) {
          // bias gradient of the layer that produced this tile: column sums over the warp's 32 rows (rows beyond M
          // hold exact zeros: zero-filled A rows and h = 0)
          const int ch = lane >> 2, el = lane & 3;
          float cs = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) cs += reinterpret_cast<const float*>(&sh[rr][ch ^ (rr & 7)])[el];
          atomicAdd(&bars->bias[c0 + lane], cs);  // shared-memory atomic (4 warps per address)
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(&bars->tmem_empty[buf]);
    }
    if (lane == 0) tma::store_wait_all();
  }

  // ---- teardown ----
  fence_before_sync();
  __syncthreads();
  if (p.colsum != nullptr && p.epi == EPI_TANHGRAD)
    for (int i = threadIdx.x; i < BN; i += kThreads)
      if (bars->bias[i] != 0.f) atomicAdd(p.colsum + i, bars->bias[i]);
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight-gradient GEMM on the tensor cores:  dW[256, IN] += sum_m dZ[m, :]^T . H[m, :]      (IN = 32..256, % 32)
// Both operands are "MN-major" for the MMA (the reduction index m = sample is the strided dimension):
//   A(out, m) = dZ[m][out],  B(in, m) = H[m][in].
// Shared-memory tile of one operand for a k-block of 32 samples: G groups of [32 samples][32 floats] (TMA box
// {32 floats, 32 rows}, SWIZZLE_128B_ATOM_32B) = the canonical MN-major layout for 32-bit operands
// (UMMA LayoutType::SWIZZLE_128B_BASE32B: atoms of 4 samples x 128 B, Swizzle<2,5,2>) with LBO = 4096 B between
// 32-column groups and SBO = 512 B between 4-sample atoms; one K=8 MMA step consumes two atoms (1024 B) per group.
// Grid = 2 output tiles (128 rows of dW each) x C sample chunks; each CTA accumulates its chunk in TMEM and adds the
// tile to dW with vector atomics.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}

constexpr int kWgXfWarps = 8;                              // transform warps (two operands to split per k-block)
constexpr int kWgradThreads = 32 * (2 + 4 + kWgXfWarps);  // producer, MMA, 4 epilogue, 8 transform = 448
__global__ void __launch_bounds__(kWgradThreads, 1)
    tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_h,
                    float* __restrict__ dW, int64_t n, int IN, int kb_per_chunk, int flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int out_tile = blockIdx.x & 1, chunk = blockIdx.x >> 1;
  const int n_kb_total = (int)((n + BK - 1) / BK);
  const int kb0 = chunk * kb_per_chunk;
  const int kb1 = (kb0 + kb_per_chunk < n_kb_total) ? kb0 + kb_per_chunk : n_kb_total;
  const int n_kb = kb1 - kb0;  // may be <= 0 for trailing chunks
  const int gB = IN / 32;      // 32-column groups of the B operand
  const uint32_t b_bytes = (uint32_t)IN * BK * 4;
  const uint32_t tmem_cols = IN <= 128 ? 128u : 256u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(&bars->full[s], 1);
      tma::mbar_init(&bars->xf[s], kWgXfWarps);
      tma::mbar_init(&bars->empty[s], 1);
    }
    tma::mbar_init(&bars->tmem_full[0], 1);
    tma::fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (n_kb > 0) {
    if (warp == 0) {
      // ---- TMA producer: lane 0 arms the barrier, lanes 0..(4+gB-1) each issue one 4 KB box ----
      const int n_box = 4 + gB;
      const int pf = (flags >> 8) ? (flags >> 8) & 0xff : 3;  // L2 prefetch distance (k-blocks), see tc_gemm_kernel
      if (lane < n_box)
        for (int j = 0; j < pf && j < n_kb; ++j) {
          if (lane < 4) tma::prefetch_2d(&tm_z, out_tile * 128 + lane * 32, (kb0 + j) * BK);
          else tma::prefetch_2d(&tm_h, (lane - 4) * 32, (kb0 + j) * BK);
        }
      for (int it = 0; it < n_kb; ++it) {
        if (lane < n_box && it + pf < n_kb && pf < 255) {
          if (lane < 4) tma::prefetch_2d(&tm_z, out_tile * 128 + lane * 32, (kb0 + it + pf) * BK);
          else tma::prefetch_2d(&tm_h, (lane - 4) * 32, (kb0 + it + pf) * BK);
        }
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        if (lane == 0) {
          tma::mbar_wait(&bars->empty[s], ph ^ 1u);
          tma::mbar_arrive_expect_tx(&bars->full[s], kATile + b_bytes);
        }
        __syncwarp();
        if (lane < n_box) {
          uint8_t* st = smem + s * kStageBytes;
          const int m0 = (kb0 + it) * BK;  // samples >= n are zero-filled
          if (lane < 4)  // A: dZ, 32-column group `lane` of this CTA's 128 output rows
            tma::load_2d(st + lane * 4096, &tm_z, out_tile * 128 + lane * 32, m0, &bars->full[s]);
          else           // B: H, 32-column group lane-4
            tma::load_2d(st + 2 * kATile + (lane - 4) * 4096, &tm_h, (lane - 4) * 32, m0, &bars->full[s]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(IN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int it = 0; it < n_kb; ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          tma::mbar_wait(&bars->xf[s], ph);  // implies full[s]: the transform warps waited for it
          fence_after_sync();
          const uint32_t sa = tma::smem_u32(smem + s * kStageBytes);
          const uint64_t a_hi = make_desc_mn(sa), a_lo = make_desc_mn(sa + kATile);
          const uint64_t b_hi = make_desc_mn(sa + 2 * kATile), b_lo = make_desc_mn(sa + 2 * kATile + b_bytes);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t koff = (uint64_t)((k * 1024) >> 4);  // next 8 samples (two 4-sample atoms)
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            mma_tf32(tmem_base, a_lo + koff, b_hi + koff, idesc, acc);
            mma_tf32(tmem_base, a_hi + koff, b_lo + koff, idesc, 1u);
            mma_tf32(tmem_base, a_hi + koff, b_hi + koff, idesc, 1u);
          }
          mma_commit(&bars->empty[s]);
        }
        mma_commit(&bars->tmem_full[0]);
      }
    } else if (warp < 6) {
      const int q = warp & 3;
      tma::mbar_wait(&bars->tmem_full[0], 0);
      fence_after_sync();
      const int row = out_tile * 128 + q * 32 + lane;  // output row of dW (always < 256)
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < IN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr0 + c0, r);
        tmem_ld_wait();
        float* dst = dW + (size_t)row * IN + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          atomicAdd(reinterpret_cast<float4*>(dst + j),
                    make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                __uint_as_float(r[j + 3])));
      }
      fence_before_sync();
    } else {
      // ---- transform warps: dZ tile and H tile -> (hi in place, lo) ----
      const int t = threadIdx.x - 6 * 32;
      const int b_count = (int)(b_bytes / 16);
      for (int it = 0; it < n_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        tma::mbar_wait(&bars->full[s], ph);
        float4* a = reinterpret_cast<float4*>(smem + s * kStageBytes);
        float4* bsrc = a + 2 * kATile / 16;
        transform_tile<32 * kWgXfWarps>(a, a + kATile / 16, kATile / 16, t, (flags & 1) != 0);
        transform_tile<32 * kWgXfWarps>(bsrc, bsrc + b_count, b_count, t, (flags & 1) != 0);
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&bars->xf[s]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// x -> (hi, lo) exact-TF32 pair; optional transpose for [R,C] -> [C,R] (weights for the dgrad GEMM)
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                    float* __restrict__ lo, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float h, l;
    split_tf32(x[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                              float* __restrict__ lo, int R, int C) {
  // out[c][r] = split(x[r][c]); small matrices (<= 256x256): simple smem-tiled transpose
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? x[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) {
      float h, l;
      split_tf32(tile[threadIdx.x][i], h, l);
      hi[(size_t)c * R + r] = h;
      lo[(size_t)c * R + r] = l;
    }
  }
}

int encode_sw128(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
int g_debug_flags = 0;

int launch(const float* a, const float* b_hi, const float* b_lo, const Params& p, cudaStream_t st) {
  if (p.K % BK != 0 || p.K <= 0 || p.M <= 0) return RB200_E_SHAPE;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b_hi) |
                       reinterpret_cast<uintptr_t>(b_lo) | reinterpret_cast<uintptr_t>(p.c) |
                       reinterpret_cast<uintptr_t>(p.h);
  if (al & 15) return RB200_E_ALIGN;
  CUtensorMap ta, tb_hi, tb_lo, tc;
  int e = encode_sw128(&ta, a, (uint64_t)p.M, (uint64_t)p.K, BM);
  if (!e) e = encode_sw128(&tb_hi, b_hi, BN, (uint64_t)p.K, BN);
  if (!e) e = encode_sw128(&tb_lo, b_lo, BN, (uint64_t)p.K, BN);
  if (!e) e = encode_sw128(&tc, p.c, (uint64_t)p.M, BN, 32);  // epilogue store boxes: [32 rows x 32 floats]
  if (e) return RB200_E_UNSUPPORTED;
  static bool attr_done = false;
  constexpr int kSmem = kStages * kStageBytes + kStagingBytes + 1024 + (int)sizeof(Barriers);
  static_assert(kSmem <= 232448, "tc_gemm_kernel shared memory");
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (ce != cudaSuccess) return (int)ce;
    attr_done = true;
  }
  const int64_t n_tiles = (p.M + BM - 1) / BM;
  const int sms = rb::sm_count();
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  tc_gemm_kernel<<<grid, kThreads, kSmem, st>>>(ta, tb_hi, tb_lo, tc, p, g_debug_flags);
  rb::count_launch();
  cudaError_t ce = cudaPeekAtLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int encode_sw128_box32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols);

int wgrad(const float* z, const float* h, float* dW, int64_t n, int IN, cudaStream_t st) {
  if (n <= 0 || IN <= 0 || IN > 256 || IN % 32 != 0) return RB200_E_SHAPE;
  const uintptr_t al =
      reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(dW);
  if (al & 15) return RB200_E_ALIGN;
  CUtensorMap tz, th;
  int e = encode_sw128_box32(&tz, z, (uint64_t)n, 256);
  if (!e) e = encode_sw128_box32(&th, h, (uint64_t)n, (uint64_t)IN);
  if (e) return RB200_E_UNSUPPORTED;
  static bool attr_done = false;
  constexpr int kSmem = kStages * kStageBytes + 1024 + (int)sizeof(Barriers);
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (ce != cudaSuccess) return (int)ce;
    attr_done = true;
  }
  const int n_kb = (int)((n + BK - 1) / BK);
  int chunks = rb::sm_count() / 2;
  if (chunks < 1) chunks = 1;
  if (chunks > n_kb) chunks = n_kb;
  const int kb_per_chunk = (n_kb + chunks - 1) / chunks;
  tc_wgrad_kernel<<<2 * chunks, kWgradThreads, kSmem, st>>>(tz, th, dW, n, IN, kb_per_chunk, g_debug_flags);
  rb::count_launch();
  cudaError_t ce = cudaPeekAtLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int split(const float* x, float* hi, float* lo, int64_t n, cudaStream_t st) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)rb::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  split_kernel<<<(int)blocks, 256, 0, st>>>(x, hi, lo, n);
  rb::count_launch();
  cudaError_t ce = cudaPeekAtLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int split_transpose(const float* x, float* hi, float* lo, int R, int C, cudaStream_t st) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  split_transpose_kernel<<<grid, block, 0, st>>>(x, hi, lo, R, C);
  rb::count_launch();
  cudaError_t ce = cudaPeekAtLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

}  // namespace tc
}  // namespace rb

// Experiment switches of the tensor-core kernels (see tc_gemm.cuh); not part of the product surface.
extern "C" int rb200_debug_set_flags(int flags) {
  rb::tc::g_debug_flags = flags;
  return RB200_OK;
}

// Debug / unit-test entry: C[M,256] (fp32) = A[M,K] . B[256,K]^T through the 3xTF32 tensor-core path.
// `work` holds the split weight operand: 2*256*K floats.
extern "C" int rb200_tc_gemm(const float* A, const float* B, float* C, int64_t M, int K, float* work,
                             rb200_stream_t stream) {
  if (!A || !B || !C || !work) return RB200_E_NULL;
  if (M <= 0 || K <= 0 || K % rb::tc::BK != 0) return RB200_E_SHAPE;
  cudaStream_t st = rb::as_stream(stream);
  float* b_hi = work;
  float* b_lo = b_hi + (int64_t)rb::tc::BN * K;
  int e;
  if ((e = rb::tc::split(B, b_hi, b_lo, (int64_t)rb::tc::BN * K, st))) return e;
  rb::tc::Params p{};
  p.M = M; p.K = K; p.c = C; p.epi = rb::tc::EPI_STORE;
  return rb::tc::launch(A, b_hi, b_lo, p, st);
}

// Unit-test entry for the weight-gradient GEMM: dW[256, IN] += Z[n,256]^T . H[n,IN]; `work` is unused (kept for
// ABI stability with the pre-split version).
extern "C" int rb200_tc_wgrad(const float* Z, const float* H, float* dW, int64_t n, int IN, float* work,
                              rb200_stream_t stream) {
  (void)work;
  if (!Z || !H || !dW) return RB200_E_NULL;
  if (n <= 0 || IN <= 0 || IN > 256 || IN % 32 != 0) return RB200_E_SHAPE;
  return rb::tc::wgrad(Z, H, dW, n, IN, rb::as_stream(stream));
}
