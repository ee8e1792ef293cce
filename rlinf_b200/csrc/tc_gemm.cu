// K3 tensor-core path: C[M,256] = epilogue( A[M,K] . B[256,K]^T ) on tcgen05 (5th-gen tensor cores), fp32-accurate
// through 3xTF32 error compensation:   a = a_hi + a_lo, b = b_hi + b_lo (each an exact TF32 number)
//     a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi        (dropped term a_lo.b_lo ~ 2^-22 relative)
// Plain TF32/bf16 misses the 1e-4 parity bar of the policy loss, so every operand is stored pre-split (hi, lo) by
// the producing kernel's epilogue and each k-step issues three kind::tf32 MMAs into the same TMEM accumulator.
//
// Structure (one CTA per SM, persistent over 128-row tiles, 192 threads):
//   warp 0   : TMA producer  - cp.async.bulk.tensor (SWIZZLE_128B boxes) of A_hi/A_lo [128x32] and B_hi/B_lo [256x32]
//                              per k-block into a 2-stage shared-memory ring (96 KB / stage), mbarrier full/empty.
//   warp 1   : MMA issuer    - one elected thread, tcgen05.mma.cta_group::1.kind::tf32 M=128 N=256 K=8, accumulators
//                              in TMEM (2 x 256 columns, double-buffered across tiles), tcgen05.commit -> mbarriers.
//   warps 2-5: epilogue      - tcgen05.ld (32 lanes x 32 columns per warp), bias+tanh or tanh' scaling, hi/lo split,
//                              transposed through a swizzled shared-memory block so that every global store / load
//                              instruction covers complete 128-byte row segments; overlaps the next tile's MMAs.
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
constexpr int kStageBytes = 2 * kATile + 2 * kBTile;  // A_hi A_lo B_hi B_lo = 96 KB
constexpr int kThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
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
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  // forward epilogue: the layer bias (broadcast LDS instead of 64 dependent LDGs / tile);
  // dgrad epilogue: per-CTA column sums of the output tiles (bias gradient), flushed once at kernel end
  float bias[BN];
  // per-epilogue-warp transpose staging [warp][hi|lo][32 rows][8 float4], 16-byte chunks XOR-swizzled by row & 7
  float4 stage[4][2][32][8];
};

__global__ void __launch_bounds__(kThreads, 1)
    tc_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo, Params p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; pointer arithmetic (no integer round trip) keeps the shared
  // address space visible to the compiler (LDS/STS instead of generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (p.M + BM - 1) / BM;
  const int n_kb = p.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      tma::mbar_init(&bars->full[s], 1);
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
  if (p.epi == EPI_BIAS_TANH_SPLIT)
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
      tma::prefetch_desc(&tm_a_hi);
      tma::prefetch_desc(&tm_a_lo);
      tma::prefetch_desc(&tm_b_hi);
      tma::prefetch_desc(&tm_b_lo);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (int)(tile * BM);
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          tma::mbar_wait(&bars->empty[s], ph ^ 1u);
          uint8_t* st = smem + s * kStageBytes;
          tma::mbar_arrive_expect_tx(&bars->full[s], kStageBytes);
          tma::load_2d(st, &tm_a_hi, kb * BK, m0, &bars->full[s]);
          tma::load_2d(st + kATile, &tm_a_lo, kb * BK, m0, &bars->full[s]);
          tma::load_2d(st + 2 * kATile, &tm_b_hi, kb * BK, 0, &bars->full[s]);
          tma::load_2d(st + 2 * kATile + kBTile, &tm_b_lo, kb * BK, 0, &bars->full[s]);
        }
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
          tma::mbar_wait(&bars->full[s], ph);
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
  } else {
    // ================= epilogue warps 2..5 =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access: lanes [32q, 32q+32)
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1u;
      const uint32_t bph = (tcount >> 1) & 1u;
      tma::mbar_wait(&bars->tmem_full[buf], bph);
      fence_after_sync();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
      float4 (*sh)[8] = bars->stage[warp - 2][0];
      float4 (*sl)[8] = bars->stage[warp - 2][1];
      const int64_t row0 = tile * BM + q * 32;   // first row of this warp's 32-row slab
      const int rs = lane >> 3, c4 = lane & 7;     // transposed phase: 4 rows x 8 float4 per instruction
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr0 + c0, r);  // asynchronous until tmem_ld_wait
        if (p.epi == EPI_TANHGRAD_SPLIT) {
          // previous activation h = hi + lo of this warp's [32 x 32] block: coalesced 128-byte row segments
          // (4 rows per instruction), transposed through shared memory to one row per thread
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rs;
            const int64_t gr = row0 + rr;
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < p.M) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p.h_hi + gr * BN + c0) + c4);
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.h_lo + gr * BN + c0) + c4);
              h = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            }
            sh[rr][c4 ^ (rr & 7)] = h;
          }
          __syncwarp();
        }
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float v[4], hi[4], lo[4];
          if (p.epi == EPI_STORE) {
#pragma unroll
            for (int e = 0; e < 4; ++e) hi[e] = __uint_as_float(r[j4 * 4 + e]);
          } else {
            if (p.epi == EPI_BIAS_TANH_SPLIT) {
              const float4 b = *reinterpret_cast<const float4*>(&bars->bias[c0 + j4 * 4]);
              v[0] = tanh_fast(__uint_as_float(r[j4 * 4 + 0]) + b.x);
              v[1] = tanh_fast(__uint_as_float(r[j4 * 4 + 1]) + b.y);
              v[2] = tanh_fast(__uint_as_float(r[j4 * 4 + 2]) + b.z);
              v[3] = tanh_fast(__uint_as_float(r[j4 * 4 + 3]) + b.w);
            } else {
              const float4 h = sh[lane][j4 ^ (lane & 7)];  // row `lane` is private to thread `lane` in this phase
              v[0] = __uint_as_float(r[j4 * 4 + 0]) * (1.0f - h.x * h.x);
              v[1] = __uint_as_float(r[j4 * 4 + 1]) * (1.0f - h.y * h.y);
              v[2] = __uint_as_float(r[j4 * 4 + 2]) * (1.0f - h.z * h.z);
              v[3] = __uint_as_float(r[j4 * 4 + 3]) * (1.0f - h.w * h.w);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(v[e], hi[e], lo[e]);
            sl[lane][j4 ^ (lane & 7)] = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
          sh[lane][j4 ^ (lane & 7)] = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
        __syncwarp();
        if (p.colsum != nullptr && p.epi == EPI_TANHGRAD_SPLIT) {
          // bias gradient of the layer that produced this tile: column sums over the warp's 32 rows (rows beyond M
          // hold exact zeros), one vector-free atomic per column per chunk
          const int ch = lane >> 2, el = lane & 3;
          float cs = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const float* ph = reinterpret_cast<const float*>(&sh[rr][ch ^ (rr & 7)]);
            const float* pl = reinterpret_cast<const float*>(&sl[rr][ch ^ (rr & 7)]);
            cs += ph[el] + pl[el];
          }
          atomicAdd(&bars->bias[c0 + lane], cs);  // shared-memory atomic (4 warps per address)
        }
        // transposed stores: every instruction writes 4 complete 128-byte row segments (no partial sectors)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rs;
          const int64_t gr = row0 + rr;
          if (gr < p.M) {
            reinterpret_cast<float4*>(p.c_hi + gr * BN + c0)[c4] = sh[rr][c4 ^ (rr & 7)];
            if (p.epi != EPI_STORE) reinterpret_cast<float4*>(p.c_lo + gr * BN + c0)[c4] = sl[rr][c4 ^ (rr & 7)];
          }
        }
        __syncwarp();
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) tma::mbar_arrive(&bars->tmem_empty[buf]);
    }
  }

  // ---- teardown ----
  fence_before_sync();
  __syncthreads();
  if (p.colsum != nullptr && p.epi == EPI_TANHGRAD_SPLIT)
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

constexpr int kWgradThreads = 192;
__global__ void __launch_bounds__(kWgradThreads, 1)
    tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_z_hi, const __grid_constant__ CUtensorMap tm_z_lo,
                    const __grid_constant__ CUtensorMap tm_h_hi, const __grid_constant__ CUtensorMap tm_h_lo,
                    float* __restrict__ dW, int64_t n, int IN, int kb_per_chunk) {
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
      // ---- TMA producer: lane 0 arms the barrier, lanes 0..(2*4+2*gB-1) each issue one 4 KB box ----
      const int n_box = 8 + 2 * gB;
      for (int it = 0; it < n_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        if (lane == 0) {
          tma::mbar_wait(&bars->empty[s], ph ^ 1u);
          tma::mbar_arrive_expect_tx(&bars->full[s], 2 * kATile + 2 * b_bytes);
        }
        __syncwarp();
        if (lane < n_box) {
          uint8_t* st = smem + s * kStageBytes;
          const int m0 = (kb0 + it) * BK;
          if (lane < 8) {  // A: dZ hi (boxes 0-3), lo (4-7); 32-column group g of this CTA's 128 output rows
            const int g = lane & 3;
            tma::load_2d(st + (lane < 4 ? 0 : kATile) + g * 4096, lane < 4 ? &tm_z_hi : &tm_z_lo,
                         out_tile * 128 + g * 32, m0, &bars->full[s]);
          } else {         // B: H hi then lo
            const int j = lane - 8;
            const bool lo = j >= gB;
            const int g = lo ? j - gB : j;
            tma::load_2d(st + 2 * kATile + (lo ? b_bytes : 0) + g * 4096, lo ? &tm_h_lo : &tm_h_hi, g * 32, m0,
                         &bars->full[s]);
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(IN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        for (int it = 0; it < n_kb; ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          tma::mbar_wait(&bars->full[s], ph);
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
    } else {
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

int launch(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, const Params& p,
           cudaStream_t st) {
  if (p.K % BK != 0 || p.K <= 0 || p.M <= 0) return RB200_E_SHAPE;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(a_lo) |
                       reinterpret_cast<uintptr_t>(b_hi) | reinterpret_cast<uintptr_t>(b_lo) |
                       reinterpret_cast<uintptr_t>(p.c_hi) | reinterpret_cast<uintptr_t>(p.c_lo);
  if (al & 15) return RB200_E_ALIGN;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int e = encode_sw128(&ta_hi, a_hi, (uint64_t)p.M, (uint64_t)p.K, BM);
  if (!e) e = encode_sw128(&ta_lo, a_lo, (uint64_t)p.M, (uint64_t)p.K, BM);
  if (!e) e = encode_sw128(&tb_hi, b_hi, BN, (uint64_t)p.K, BN);
  if (!e) e = encode_sw128(&tb_lo, b_lo, BN, (uint64_t)p.K, BN);
  if (e) return RB200_E_UNSUPPORTED;
  static bool attr_done = false;
  constexpr int kSmem = kStages * kStageBytes + 1024 + (int)sizeof(Barriers);
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (ce != cudaSuccess) return (int)ce;
    attr_done = true;
  }
  const int64_t n_tiles = (p.M + BM - 1) / BM;
  const int sms = rb::sm_count();
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  tc_gemm_kernel<<<grid, kThreads, kSmem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
  rb::count_launch();
  cudaError_t ce = cudaPeekAtLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

int encode_sw128_box32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols);

int wgrad(const float* z_hi, const float* z_lo, const float* h_hi, const float* h_lo, float* dW, int64_t n, int IN,
          cudaStream_t st) {
  if (n <= 0 || IN <= 0 || IN > 256 || IN % 32 != 0) return RB200_E_SHAPE;
  const uintptr_t al = reinterpret_cast<uintptr_t>(z_hi) | reinterpret_cast<uintptr_t>(z_lo) |
                       reinterpret_cast<uintptr_t>(h_hi) | reinterpret_cast<uintptr_t>(h_lo) |
                       reinterpret_cast<uintptr_t>(dW);
  if (al & 15) return RB200_E_ALIGN;
  CUtensorMap tz_hi, tz_lo, th_hi, th_lo;
  int e = encode_sw128_box32(&tz_hi, z_hi, (uint64_t)n, 256);
  if (!e) e = encode_sw128_box32(&tz_lo, z_lo, (uint64_t)n, 256);
  if (!e) e = encode_sw128_box32(&th_hi, h_hi, (uint64_t)n, (uint64_t)IN);
  if (!e) e = encode_sw128_box32(&th_lo, h_lo, (uint64_t)n, (uint64_t)IN);
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
  tc_wgrad_kernel<<<2 * chunks, kWgradThreads, kSmem, st>>>(tz_hi, tz_lo, th_hi, th_lo, dW, n, IN, kb_per_chunk);
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

// Debug / unit-test entry: C[M,256] (fp32) = A[M,K] . B[256,K]^T through the 3xTF32 tensor-core path.
// `work` holds the split operands: 2*M*K + 2*256*K floats.
extern "C" int rb200_tc_gemm(const float* A, const float* B, float* C, int64_t M, int K, float* work,
                             rb200_stream_t stream) {
  if (!A || !B || !C || !work) return RB200_E_NULL;
  if (M <= 0 || K <= 0 || K % rb::tc::BK != 0) return RB200_E_SHAPE;
  cudaStream_t st = rb::as_stream(stream);
  float* a_hi = work;
  float* a_lo = a_hi + M * K;
  float* b_hi = a_lo + M * K;
  float* b_lo = b_hi + (int64_t)rb::tc::BN * K;
  int e;
  if ((e = rb::tc::split(A, a_hi, a_lo, M * K, st))) return e;
  if ((e = rb::tc::split(B, b_hi, b_lo, (int64_t)rb::tc::BN * K, st))) return e;
  rb::tc::Params p{};
  p.M = M; p.K = K; p.c_hi = C; p.c_lo = C; p.epi = rb::tc::EPI_STORE;
  return rb::tc::launch(a_hi, a_lo, b_hi, b_lo, p, st);
}

// Unit-test entry for the weight-gradient GEMM: dW[256, IN] += Z[n,256]^T . H[n,IN]; work = 2*n*(256+IN) floats.
extern "C" int rb200_tc_wgrad(const float* Z, const float* H, float* dW, int64_t n, int IN, float* work,
                              rb200_stream_t stream) {
  if (!Z || !H || !dW || !work) return RB200_E_NULL;
  if (n <= 0 || IN <= 0 || IN > 256 || IN % 32 != 0) return RB200_E_SHAPE;
  cudaStream_t st = rb::as_stream(stream);
  float* z_hi = work;
  float* z_lo = z_hi + n * 256;
  float* h_hi = z_lo + n * 256;
  float* h_lo = h_hi + n * IN;
  int e;
  if ((e = rb::tc::split(Z, z_hi, z_lo, n * 256, st))) return e;
  if ((e = rb::tc::split(H, h_hi, h_lo, n * IN, st))) return e;
  return rb::tc::wgrad(z_hi, z_lo, h_hi, h_lo, dW, n, IN, st);
}
