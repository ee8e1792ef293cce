// EXPERIMENTAL (round-2 groundwork, not yet run on a GPU; selected only by rb200_debug_set_flags bit 2):
// cta_group::2 variant of the 3xTF32 forward/dgrad GEMM of tc_gemm.cu.
//
// Why: tc_gemm_kernel is bound by shared-memory bandwidth, not by the tensor pipe: per
// 128x256x32 k-block the three MMAs re-read both operands (147 KB) and TMA refills 64 KB of weights per tile.
// With a CTA pair (two SMs of one TPC) on a 256-row tile, each CTA holds its own 128 rows of A and only HALF of the
// weight tile (128 of the 256 output columns); the pair's MMA (M = 256) reads A from both CTAs and the two B halves:
// per CTA and k-block the operand reads drop from 147 to 98 KB and the weight fill from 64 to 32 KB, and the freed
// shared memory buys a third pipeline stage (64 KB / stage instead of 96 KB).
//
// Protocol (r = %cluster_ctarank, leader = rank 0; identical shared-memory layout in both CTAs):
//   producer (warp 0, both CTAs): A [128 x 32] of its own rows -> local `afull`; its B half (hi, lo) with
//             cp.async.bulk.tensor .cta_group::2, complete_tx on the LEADER's `bfull` (expects both halves).
//   transform (warps 2-5, both CTAs): wait local `afull`, write A_lo, fence.proxy.async, arrive on the LEADER's `xf`
//             (8 arrivals: 4 local + 4 remote, mbarrier.arrive.release.cluster through mapa).
//   MMA (warp 1 lane 0, leader only): wait `bfull`, `xf`; 12 x tcgen05.mma.cta_group::2.kind::tf32 (M256 N256 K8);
//             tcgen05.commit ... multicast::cluster 0b11 -> `empty[s]` / `tmem_full[b]` of BOTH CTAs.
//   epilogue (warps 6-9, both CTAs): drain the own 128 TMEM lanes exactly as tc_gemm_kernel, then arrive on the
//             LEADER's `tmem_empty` (8 arrivals).
// Results are expected to be bit-identical to tc_gemm_kernel (same products, same k order, same accumulator).
// Operand placement cross-checked against the CUTLASS/CuTe headers vendored in this image (cute/atom/mma_traits_sm100:
// SM100_MMA_TF32_2x1SM_SS has ALayout 2 x (M/2, K), BLayout 2 x (N/2, K), CLayout 2 x (M/2, N); Allocator2Sm: the same
// warp of BOTH CTAs issues tcgen05.alloc.cta_group::2; SM100_TMA_2SM_LOAD_2D: complete_tx on CTA 0's barrier;
// umma_arrive_multicast_2x1SM: the multicast commit used below).
#include "common.cuh"
#include "tc_gemm.cuh"
#include "tma.cuh"

namespace rb {
namespace tc {

int encode_sw128(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

namespace {

constexpr int kStages2 = 3;
constexpr int kATile = BM * BK * 4;        // 16 KB: this CTA's 128 rows
constexpr int kBHalf = (BN / 2) * BK * 4;  // 16 KB: this CTA's 128 of the 256 output columns
constexpr int kStageBytes2 = 2 * kATile + 2 * kBHalf;  // A(hi) A_lo B_hi B_lo = 64 KB
constexpr int kXfWarps = 4;
constexpr int kThreads2 = 32 * (2 + kXfWarps + 4);
constexpr int kStagingBytes = 4 * 2 * 4096;
constexpr int kTmemCols = 512;
constexpr uint32_t kTf32Mask = 0xffffe000u;

// ---- cluster helpers ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(tma::smem_u32(local)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of this CTA's box whose completion is counted on a barrier that may live in the peer CTA
__device__ __forceinline__ void load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                             uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(tma::smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05, cta_group::2 ----
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tma::smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on the barrier at this offset in every CTA of `mask` once all prior MMAs of this thread have completed
__device__ __forceinline__ void mma_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          tma::smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mma2_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t}" ::"r"(d_tmem),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u),
      "r"(0u), "r"(0u)
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

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {  // K-major SWIZZLE_128B, see tc_gemm.cu
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::tf32, fp32 accumulate, K-major A and B, N = 256, M = 256 (the pair's tile)
constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ float tanh_fast(float x) {
  const float t = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}
__device__ __forceinline__ float lo1(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & kTf32Mask);
  return __uint_as_float((__float_as_uint(__fsub_rn(x, hi)) + 0x1000u) & kTf32Mask);
}

struct __align__(16) Barriers2 {
  uint64_t afull[kStages2];   // local: this CTA's A box has landed
  uint64_t bfull[kStages2];   // leader's copy is used: both CTAs' weight halves have landed
  uint64_t xf[kStages2];      // leader's copy is used: 2 x 4 transform warps
  uint64_t empty[kStages2];   // local copies, multicast commit
  uint64_t tmem_full[2];      // local copies, multicast commit
  uint64_t tmem_empty[2];     // leader's copy is used: 2 x 4 epilogue warps
  uint32_t tmem_base;
  uint32_t pad_[3];
  float bias[BN];
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
    tc_gemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b_hi,
                    const __grid_constant__ CUtensorMap tm_b_lo, const __grid_constant__ CUtensorMap tm_c, Params p,
                    int flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + kStages2 * kStageBytes2;
  Barriers2* bars = reinterpret_cast<Barriers2*>(staging + kStagingBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int64_t n_tiles = (p.M + 2 * BM - 1) / (2 * BM);  // 256-row tiles of the pair
  const int n_kb = p.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) {
      tma::mbar_init(&bars->afull[s], 1);
      tma::mbar_init(&bars->bfull[s], 1);
      tma::mbar_init(&bars->xf[s], 2 * kXfWarps);
      tma::mbar_init(&bars->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tma::mbar_init(&bars->tmem_full[b], 1);
      tma::mbar_init(&bars->tmem_empty[b], 2 * 4);
    }
    tma::fence_barrier_init();
  }
  if (warp == 1) {  // one warp of EACH CTA of the pair
    tmem_alloc2(&bars->tmem_base, kTmemCols);
    tmem_relinquish2();
  }
  if (p.epi == EPI_BIAS_TANH)
    for (int i = threadIdx.x; i < BN; i += kThreads2) bars->bias[i] = p.bias[i];
  else
    for (int i = threadIdx.x; i < BN; i += kThreads2) bars->bias[i] = 0.f;
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / complete_tx
  fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    if (lane == 0) {
      tma::prefetch_desc(&tm_a);
      tma::prefetch_desc(&tm_b_hi);
      tma::prefetch_desc(&tm_b_lo);
      const int64_t my_tiles = n_tiles > pair ? (n_tiles - pair + n_pairs - 1) / n_pairs : 0;
      const int64_t total = my_tiles * n_kb;
      const int pf = (flags >> 8) ? (flags >> 8) & 0xff : 3;
      auto row_of = [&](int64_t j) { return (int)((pair + (j / n_kb) * n_pairs) * (2 * BM) + rank * BM); };
      for (int64_t j = 0; j < pf && j < total; ++j) tma::prefetch_2d(&tm_a, (int)(j % n_kb) * BK, row_of(j));
      for (int64_t j = 0; j < total; ++j) {
        const uint32_t it = (uint32_t)j;
        const int kb = (int)(j % n_kb);
        const int64_t jp = j + pf;
        if (jp < total && pf < 255) tma::prefetch_2d(&tm_a, (int)(jp % n_kb) * BK, row_of(jp));
        const int s = it % kStages2;
        const uint32_t ph = (it / kStages2) & 1u;
        tma::mbar_wait(&bars->empty[s], ph ^ 1u);
        uint8_t* st = smem + s * kStageBytes2;
        tma::mbar_arrive_expect_tx(&bars->afull[s], kATile);
        tma::load_2d(st, &tm_a, kb * BK, row_of(j), &bars->afull[s]);
        if (leader) tma::mbar_arrive_expect_tx(&bars->bfull[s], 2 * 2 * kBHalf);  // both CTAs' halves
        const uint32_t bfull_leader = map_to_rank(&bars->bfull[s], 0);
        load_2d_pair(st + 2 * kATile, &tm_b_hi, kb * BK, (int)rank * (BN / 2), bfull_leader);
        load_2d_pair(st + 2 * kATile + kBHalf, &tm_b_lo, kb * BK, (int)rank * (BN / 2), bfull_leader);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA, single thread) =================
    if (leader && lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++tcount) {
        const uint32_t buf = tcount & 1u;
        const uint32_t bph = (tcount >> 1) & 1u;
        tma::mbar_wait(&bars->tmem_empty[buf], bph ^ 1u);
        fence_after_sync();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
          const int s = it % kStages2;
          const uint32_t ph = (it / kStages2) & 1u;
          tma::mbar_wait(&bars->bfull[s], ph);
          tma::mbar_wait(&bars->xf[s], ph);
          fence_after_sync();
          const uint32_t sa = tma::smem_u32(smem + s * kStageBytes2);
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + kATile);
          const uint64_t b_hi = make_desc(sa + 2 * kATile), b_lo = make_desc(sa + 2 * kATile + kBHalf);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);
            const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
            mma2_tf32(d_tmem, a_lo + koff, b_hi + koff, kIdesc2, acc);
            mma2_tf32(d_tmem, a_hi + koff, b_lo + koff, kIdesc2, 1u);
            mma2_tf32(d_tmem, a_hi + koff, b_hi + koff, kIdesc2, 1u);
          }
          mma_commit2(&bars->empty[s], 0b11);
        }
        mma_commit2(&bars->tmem_full[buf], 0b11);
      }
    }
  } else if (warp < 2 + kXfWarps) {
    // ================= transform warps (both CTAs): own A tile -> A_lo =================
    const int t = threadIdx.x - 64;
    uint32_t it = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs) {
      for (int kb = 0; kb < n_kb; ++kb, ++it) {
        const int s = it % kStages2;
        const uint32_t ph = (it / kStages2) & 1u;
        tma::mbar_wait(&bars->afull[s], ph);
        const float4* a = reinterpret_cast<const float4*>(smem + s * kStageBytes2);
        float4* al = reinterpret_cast<float4*>(smem + s * kStageBytes2 + kATile);
#pragma unroll 4
        for (int i = t; i < kATile / 16; i += 32 * kXfWarps) {
          const float4 x = a[i];
          al[i] = make_float4(lo1(x.x), lo1(x.y), lo1(x.z), lo1(x.w));
        }
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (leader) tma::mbar_arrive(&bars->xf[s]);
          else mbar_arrive_cluster(map_to_rank(&bars->xf[s], 0));
        }
      }
    }
  } else {
    // ================= epilogue warps (both CTAs): same as tc_gemm_kernel on the own 128 rows =================
    const int q = warp & 3;
    const int ew = warp - (2 + kXfWarps);
    float4 (*stg)[32][8] = reinterpret_cast<float4 (*)[32][8]>(staging + ew * 8192);
    if (lane == 0) tma::prefetch_desc(&tm_c);
    const int rs = lane >> 3, c4 = lane & 7;
    float4 hp[8];
    auto load_h = [&](int64_t r0, int cc) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t gr = r0 + it * 4 + rs;
        hp[it] = gr < p.M ? __ldg(reinterpret_cast<const float4*>(p.h + gr * BN + cc) + c4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto slab_row = [&](int64_t tile) { return tile * (2 * BM) + (int64_t)rank * BM + q * 32; };
    if (p.epi == EPI_TANHGRAD && (int64_t)pair < n_tiles) load_h(slab_row(pair), 0);
    uint32_t tcount = 0;
    for (int64_t tile = pair; tile < n_tiles; tile += n_pairs, ++tcount) {
      const uint32_t buf = tcount & 1u;
      const uint32_t bph = (tcount >> 1) & 1u;
      tma::mbar_wait(&bars->tmem_full[buf], bph);
      fence_after_sync();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
      const int64_t row0 = slab_row(tile);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float4 (*sh)[8] = stg[(c0 >> 5) & 1];
        if (lane == 0) tma::store_wait_read1();
        __syncwarp();
        uint32_t r[32];
        tmem_ld32(taddr0 + c0, r);
        if (p.epi == EPI_TANHGRAD) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rs;
            sh[rr][c4 ^ (rr & 7)] = hp[it];
          }
          __syncwarp();
          if (c0 + 32 < BN) load_h(row0, c0 + 32);
          else if (tile + n_pairs < n_tiles) load_h(slab_row(tile + n_pairs), 0);
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
            const float4 h = sh[lane][j4 ^ (lane & 7)];
            v.x = __uint_as_float(r[j4 * 4 + 0]) * (1.0f - h.x * h.x);
            v.y = __uint_as_float(r[j4 * 4 + 1]) * (1.0f - h.y * h.y);
            v.z = __uint_as_float(r[j4 * 4 + 2]) * (1.0f - h.z * h.z);
            v.w = __uint_as_float(r[j4 * 4 + 3]) * (1.0f - h.w * h.w);
          }
          sh[lane][j4 ^ (lane & 7)] = v;
        }
        tma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma::store_2d(&tm_c, &sh[0][0], c0, (int)row0);
          tma::store_commit();
        }
        if (p.colsum != nullptr && p.epi == EPI_TANHGRAD) {
          const int ch = lane >> 2, el = lane & 3;
          float cs = 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) cs += reinterpret_cast<const float*>(&sh[rr][ch ^ (rr & 7)])[el];
          atomicAdd(&bars->bias[c0 + lane], cs);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (leader) tma::mbar_arrive(&bars->tmem_empty[buf]);
        else mbar_arrive_cluster(map_to_rank(&bars->tmem_empty[buf], 0));
      }
    }
    if (lane == 0) tma::store_wait_all();
  }

  // ---- teardown: neither CTA may leave while the peer can still touch its shared memory / barriers ----
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (p.colsum != nullptr && p.epi == EPI_TANHGRAD)
    for (int i = threadIdx.x; i < BN; i += kThreads2)
      if (bars->bias[i] != 0.f) atomicAdd(p.colsum + i, bars->bias[i]);
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

}  // namespace

// Launcher of the experimental pair kernel; same contract as rb::tc::launch (tc_gemm.cu).
int launch_pair(const float* a, const float* b_hi, const float* b_lo, const Params& p, cudaStream_t st) {
  if (p.K % BK != 0 || p.K <= 0 || p.M <= 0) return RB200_E_SHAPE;
  CUtensorMap ta, tb_hi, tb_lo, tc;
  int e = encode_sw128(&ta, a, (uint64_t)p.M, (uint64_t)p.K, BM);
  if (!e) e = encode_sw128(&tb_hi, b_hi, BN, (uint64_t)p.K, BN / 2);  // box = one CTA's half of the output columns
  if (!e) e = encode_sw128(&tb_lo, b_lo, BN, (uint64_t)p.K, BN / 2);
  if (!e) e = encode_sw128(&tc, p.c, (uint64_t)p.M, BN, 32);
  if (e) return RB200_E_UNSUPPORTED;
  static bool attr_done = false;
  constexpr int kSmem = kStages2 * kStageBytes2 + kStagingBytes + 1024 + (int)sizeof(Barriers2);
  static_assert(kSmem <= 232448, "tc_gemm2_kernel shared memory");
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(tc_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (ce != cudaSuccess) return (int)ce;
    attr_done = true;
  }
  const int64_t n_tiles = (p.M + 2 * BM - 1) / (2 * BM);
  int pairs = rb::sm_count() / 2;
  if (pairs > n_tiles) pairs = (int)n_tiles;
  if (pairs < 1) pairs = 1;
  tc_gemm2_kernel<<<2 * pairs, kThreads2, kSmem, st>>>(ta, tb_hi, tb_lo, tc, p, g_debug_flags);
  rb::count_launch();
  cudaError_t ce = cudaPeekAtLastError();
  return ce == cudaSuccess ? 0 : (int)ce;
}

}  // namespace tc
}  // namespace rb
