// Two-CTA ("cta_group::2") variant of the gather GEMM (gemm_fused_tc.cu) for N = 256, K <= 256 -- the edge-gate and
// data-gradient GEMMs of the conv path.
//
// Why: the one-CTA kernel is bound by the SM's shared-memory / L1 data pipe (128 B/clk), not by HBM or the tensor
// pipe: per 128-row tile the tensor core alone fetches 576 KB of operands from shared memory (3 MMAs per K step x
// (4 KB A + 8 KB B)), the weight chunks are re-written into the ring (256 KB), and the A conversion and the epilogue
// transposition add ~900 KB.  A CTA PAIR runs M = 256 MMAs: each CTA supplies its own 128 rows
// of A and HALF of W (128 of the 256 output columns), so per CTA the operand fetch is 8 KB per MMA instead of 12, and
// the half-W image for the whole K (128 KB of bf16 hi/lo planes) stays RESIDENT in shared memory -- no weight ring,
// no re-streaming from L2.
//
// Cluster of 2 CTAs (adjacent SMs of one TPC).  Each CTA keeps its own A pipeline (TMA boxes -> converter warps ->
// bf16 planes), epilogue and TMEM accumulator (its 128 rows x 256 columns).  Only CTA 0 issues tcgen05.mma
// (.cta_group::2); the converter warps of BOTH CTAs arrive on CTA 0's "planes full" barrier (remote mbarrier arrive
// over the cluster), tcgen05.commit multicasts "planes free" / "accumulator full" to both CTAs, and the epilogue warps
// of both CTAs arrive on CTA 0's "accumulator drained" barrier.
#include <atomic>
#include <cuda.h>

#include "tc_common.cuh"
#include "api_common.h"
#include "alignn_b200.h"

namespace alignn {
namespace gemm2 {
// shared with gemm_fused_tc.cu
struct Params;
int make_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);
}  // namespace gemm2

namespace gemmp {

constexpr int BM = 128;         // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 32;
constexpr int NSA = 2;          // fp32 A staging ring (TMA destination)
constexpr int NSP = 2;          // bf16 A plane ring
constexpr int EPI_WARPS = 8;          // two per TMEM lane quarter, alternating 32-column chunks
constexpr int CONV_WARPS = 4;
constexpr int THREADS = 32 * (4 + EPI_WARPS + CONV_WARPS);   // 384
constexpr uint32_t LBO = 128;
constexpr uint32_t SBO = (BK / 8) * 128;
constexpr int A_STAGE = BM * BK * 4;
constexpr int A_PLANE = BM * BK * 2;
constexpr int W_PLANE = (BN / 2) * BK * 2;    // half of the output columns: 8 KB per plane and K chunk
constexpr int MAX_NK = 8;                     // K <= 256
constexpr int EPI_BOX = 32 * 32 * 4;
constexpr bool kTmaStore = true;
constexpr int EPI_BUFS = 1;
constexpr int OFF_ASTG = 0;
constexpr int OFF_APL = OFF_ASTG + NSA * A_STAGE;
constexpr int OFF_W = OFF_APL + NSP * 2 * A_PLANE;
constexpr int OFF_EPI = OFF_W + MAX_NK * 2 * W_PLANE;
constexpr int OFF_BAR = OFF_EPI + EPI_WARPS * EPI_BUFS * EPI_BOX;     // (column statistics accumulate in global memory)
constexpr int SMEM = OFF_BAR + 256;
constexpr int TMEM_COLS = 2 * BN;
static_assert(SMEM <= 232448, "shared memory budget of one sm_100 CTA");
static_assert(OFF_EPI % 1024 == 0, "swizzled boxes need 1024-byte alignment");

struct Params {
  int M, N, K;
  const uint8_t* w_image;
  const float* bias;
  const float* add0; int64_t ld0; const int32_t* idx0;
  const float* add1; int64_t ld1; const int32_t* idx1;
  float* C; int64_t ldc;
  float* stats;
  const float* bn_scale; const float* bn_shift; const float* bn_mean;     // (unused here: the BatchNorm mode stays on the one-CTA kernel)
};

__host__ __device__ constexpr int plane_off(int r, int k) { return (r >> 3) * (int)SBO + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(tc::smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(tc::smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ float dsilu_(float u) { float e, sg;                                             // 4-instruction sigmoid, as common.cuh's sigmoidf_
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + e)); return sg * (1.f + u * (1.f - sg)); }

// ---- cluster plumbing -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// arrive (release at cluster scope) on a barrier given by its shared::cluster address -- possibly in the other CTA
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// bounded wait with cluster-scope acquire (the arrivals may come from the other CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = tc::smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot_in_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[own 128 rows each] * B[half of the N rows each]^T ; issued by ONE thread of CTA 0
__device__ __forceinline__ void mma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <bool BNMODE_UNUSED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_gather_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapC, const Params p) {
  constexpr bool BNMODE = false;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* astg_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* astg_empty = astg_full + NSA;
  uint64_t* apl_full = astg_empty + NSA;      // used in CTA 0 only: 2 * CONV_WARPS arrivals
  uint64_t* apl_empty = apl_full + NSP;       // per CTA, multicast commit
  uint64_t* w_full = apl_empty + NSP;
  uint64_t* tfull = w_full + 1;               // per CTA, multicast commit
  uint64_t* tempty = tfull + 2;               // used in CTA 0 only: 2 * EPI_WARPS arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int nk = p.K / BK;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total = (p.M + 2 * BM - 1) / (2 * BM);                 // 256-row pair tiles
  const int my_tiles = (total - pair + npairs - 1) / npairs;

  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { tc::mbar_init(&astg_full[s], 1); tc::mbar_init(&astg_empty[s], CONV_WARPS); }
    for (int s = 0; s < NSP; ++s) { tc::mbar_init(&apl_full[s], 2 * CONV_WARPS); tc::mbar_init(&apl_empty[s], 1); }
    tc::mbar_init(w_full, 1);
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 2 * EPI_WARPS); }
    tc::mbar_fence_init();
  }
  if (warp == 0) tmem_alloc2(tmem_slot, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  // W: this CTA's half (output columns 128 * rank ...) of every K chunk, loaded ONCE and kept for the whole kernel
  if (warp == 2 && lane == 0) {
    tc::mbar_arrive_expect_tx(w_full, (uint32_t)(nk * 2 * W_PLANE));
    for (int kc = 0; kc < nk; ++kc) {
      const uint8_t* chunk = p.w_image + (int64_t)kc * 2 * (2 * W_PLANE);              // image chunk: [hi 16 KB | lo 16 KB]
      tc::bulk_g2s(smem + OFF_W + kc * 2 * W_PLANE, chunk + rank * W_PLANE, W_PLANE, w_full);
      tc::bulk_g2s(smem + OFF_W + kc * 2 * W_PLANE + W_PLANE, chunk + 2 * W_PLANE + rank * W_PLANE, W_PLANE, w_full);
    }
    tc::mbar_wait(w_full, 0);
  }
  __syncthreads();
  cluster_sync_all();                         // both CTAs: barriers initialised and weights resident before any remote arrive / MMA
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t apl_full_leader = mapa_rank(tc::smem_u32(apl_full), 0);
  const uint32_t tempty_leader = mapa_rank(tc::smem_u32(tempty), 0);

  if (warp == 1) {
    // ================= A producer: TMA boxes of [128 rows x 32 floats], this CTA's half of the pair tile =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      int c = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int ptile = pair + lt * npairs;
        const int m0 = ptile * 2 * BM + (int)rank * BM;
        if (lt + 1 < my_tiles)
          for (int kc = 0; kc < nk; ++kc) tma_prefetch_2d(&mapA, kc * BK, m0 + npairs * 2 * BM);
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % NSA;
          if (c >= NSA) tc::mbar_wait(&astg_empty[s], ((c / NSA) - 1) & 1);
          tc::mbar_arrive_expect_tx(&astg_full[s], A_STAGE);
          tma_load_2d(smem + OFF_ASTG + s * A_STAGE, &mapA, kc * BK, m0, &astg_full[s]);
        }
      }
    }
  } else if (warp >= 4 + EPI_WARPS) {
    // ================= converters: staged fp32 box -> bf16 hi/lo planes (as in gemm_fused_tc.cu) =================
    const int cw = warp - (4 + EPI_WARPS);
    int ld_off[8], st_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int u = (i * CONV_WARPS + cw) * 2 + (lane >> 4);
      const int row = (u >> 2) * 8 + (lane & 7);
      const int kq = (u & 3) * 2 + ((lane >> 3) & 1);
      ld_off[i] = row * 128 + ((kq ^ (row & 7)) << 4);
      st_off[i] = plane_off(row, kq * 4);
    }
    const int nchunks = my_tiles * nk;
    for (int c = 0; c < nchunks; ++c) {
      const int sa = c % NSA, sp = c % NSP;
      tc::mbar_wait(&astg_full[sa], (c / NSA) & 1);
      const uint8_t* src = smem + OFF_ASTG + sa * A_STAGE;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(src + ld_off[i]);
      if (c >= NSP) tc::mbar_wait(&apl_empty[sp], ((c / NSP) - 1) & 1);      // multicast commit from CTA 0's MMA thread
      uint8_t* dst = smem + OFF_APL + sp * 2 * A_PLANE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint2 hi, lo;
        tc::split4(v[i], hi, lo);
        *reinterpret_cast<uint2*>(dst + st_off[i]) = hi;
        *reinterpret_cast<uint2*>(dst + A_PLANE + st_off[i]) = lo;
      }
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(apl_full_leader + 8u * (uint32_t)sp);     // CTA 0 counts the warps of both CTAs
        tc::mbar_arrive(&astg_empty[sa]);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int ew = warp - 4;
    const int q = warp & 3;                                  // TMEM lane quarter = rows 32q .. 32q+31 of the tile
    uint8_t* stg0 = smem + OFF_EPI + ew * EPI_BUFS * EPI_BOX;
    // staged element (row r, 16-byte chunk j) lives at r * 128 + ((j ^ (r & 7)) << 4): what a SWIZZLE_128B tensor map
    // expects, and conflict-free for both the row-per-thread writes and the 8-lanes-per-row reads
    const int st_row = lane * 128, st_sw = lane & 7;
    uint32_t chunk_ctr = 0;
    // column statistics: one private row per (CTA, lane quarter) in the caller's partial buffer, read-modify-written by
    // the warp that owns the chunk (there is no shared memory left for accumulators; the row is L2 resident)
    float* stat = p.stats ? p.stats + ((int64_t)blockIdx.x * 4 + q) * 2 * BN : nullptr;
    const int half = ew >> 2;                                // which of the quarter's two warps: chunks half, half + 2, ...
    const bool do_stats = p.stats != nullptr;
    constexpr bool bnmode = BNMODE;
    if (do_stats) {
      for (int ch = half; ch < BN / 32; ch += 2) { __stcg(stat + ch * 32 + lane, 0.f); __stcg(stat + BN + ch * 32 + lane, 0.f); }
      __syncwarp();
    }
    const int rsub = lane >> 3;                              // row within a group of 4
    const int c4 = (lane & 7) * 4;                           // 4 of the chunk's 32 columns
    constexpr int NCH = BN / 32;
    uint32_t lt = 0;
    for (int ptile = pair; ptile < total; ptile += npairs, ++lt) {
      const int acc = lt & 1;
      const int m0 = ptile * 2 * BM + (int)rank * BM, n0 = 0;
      // rows this lane finishes: it * 4 + rsub, it < 8.  i0 / i1 = addend row of each (-1: no addend / row past M)
      int i0[8], i1[8];
      uint32_t rvm = 0;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int gr = m0 + q * 32 + it * 4 + rsub;
        const bool ok = gr < p.M;
        rvm |= (ok ? 1u : 0u) << it;
        i0[it] = (p.add0 && ok) ? (p.idx0 ? __ldg(p.idx0 + gr) : gr) : -1;
        i1[it] = (p.add1 && ok) ? (p.idx1 ? __ldg(p.idx1 + gr) : gr) : -1;
      }
      const float* base0 = p.add0 + n0 + c4 + half * 32;       // first chunk of this warp
      const float* base1 = p.add1 + n0 + c4 + half * 32;
      float4 a0[8], a1[8];
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        a0[it] = (i0[it] >= 0 && half < NCH) ? __ldg(reinterpret_cast<const float4*>(base0 + (int64_t)i0[it] * p.ld0)) : z4;
        a1[it] = (i1[it] >= 0 && half < NCH) ? __ldg(reinterpret_cast<const float4*>(base1 + (int64_t)i1[it] * p.ld1)) : z4;
      }
      tc::mbar_wait(&tfull[acc], (lt >> 1) & 1);
      tc::fence_after_sync();
#pragma unroll 1
      for (int ch = half; ch < NCH; ch += 2) {
        const int c0 = ch * 32;
        uint8_t* stg = stg0;
        if (kTmaStore && chunk_ctr >= 1) {       // the TMA store that read this box must be done with it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        ++chunk_ctr;
        float4 st_old = z4, sq_old = z4;
        if (do_stats && lane < 8) {                      // requested now, needed at the end of the chunk
          st_old = __ldcg(reinterpret_cast<const float4*>(stat + c0 + c4));
          sq_old = __ldcg(reinterpret_cast<const float4*>(stat + BN + c0 + c4));
        }
        {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
    #pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + st_row + ((j ^ st_sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncwarp();
          float4 b4 = z4;
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + c4));
        float4 bsc = z4, bsh = z4, bmu = z4;
        if (bnmode) {
          bsc = __ldg(reinterpret_cast<const float4*>(p.bn_scale + n0 + c0 + c4));
          bsh = __ldg(reinterpret_cast<const float4*>(p.bn_shift + n0 + c0 + c4));
          bmu = __ldg(reinterpret_cast<const float4*>(p.bn_mean + n0 + c0 + c4));
        }
        float4 s4 = z4, q4 = z4;
        const bool more = ch + 2 < NCH;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rsub;
          float4* cell = reinterpret_cast<float4*>(stg + r * 128 + (((lane & 7) ^ (r & 7)) << 4));
          float4 o = *cell;
          const float4 mrow = a1[it];                       // (BatchNorm mode: the pre-norm row, not an addend)
          const float4 ad1 = bnmode ? z4 : mrow;
          o.x = (o.x + b4.x) + (a0[it].x + ad1.x);
          o.y = (o.y + b4.y) + (a0[it].y + ad1.y);
          o.z = (o.z + b4.z) + (a0[it].z + ad1.z);
          o.w = (o.w + b4.w) + (a0[it].w + ad1.w);
          // this row's addends of the NEXT chunk go out now and land while the rest of this chunk is processed
          if (more) {
            if (i0[it] >= 0) a0[it] = __ldg(reinterpret_cast<const float4*>(base0 + (int64_t)i0[it] * p.ld0 + (c0 - half * 32) + 64));
            if (i1[it] >= 0) a1[it] = __ldg(reinterpret_cast<const float4*>(base1 + (int64_t)i1[it] * p.ld1 + (c0 - half * 32) + 64));
          }
          if ((rvm >> it) & 1u) {
            if (bnmode) {
              // gu = o * silu'(m * scale + shift);  partial sums of gu and gu * (m - mean): the two reductions of the
              // train-mode BatchNorm backward (c1, c2) for the layer that consumes this gradient
              float4 gu;
              gu.x = o.x * dsilu_(fmaf(mrow.x, bsc.x, bsh.x)); gu.y = o.y * dsilu_(fmaf(mrow.y, bsc.y, bsh.y));
              gu.z = o.z * dsilu_(fmaf(mrow.z, bsc.z, bsh.z)); gu.w = o.w * dsilu_(fmaf(mrow.w, bsc.w, bsh.w));
              s4.x += gu.x; s4.y += gu.y; s4.z += gu.z; s4.w += gu.w;
              q4.x = fmaf(gu.x, mrow.x - bmu.x, q4.x); q4.y = fmaf(gu.y, mrow.y - bmu.y, q4.y);
              q4.z = fmaf(gu.z, mrow.z - bmu.z, q4.z); q4.w = fmaf(gu.w, mrow.w - bmu.w, q4.w);
            } else {
              s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
              q4.x = fmaf(o.x, o.x, q4.x); q4.y = fmaf(o.y, o.y, q4.y); q4.z = fmaf(o.z, o.z, q4.z); q4.w = fmaf(o.w, o.w, q4.w);
            }
          }
          if constexpr (kTmaStore) *cell = o;
          else if ((rvm >> it) & 1u)
            *reinterpret_cast<float4*>(p.C + (int64_t)(m0 + q * 32 + r) * p.ldc + n0 + c0 + c4) = o;
        }
        if constexpr (kTmaStore) {   // finished [32 x 32] box -> global by TMA (rows past M are clipped by the tensor map)
          tc::fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapC, stg, n0 + c0, m0 + q * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else {
          __syncwarp();               // the transpose box is rewritten by the next chunk
        }
        if (do_stats) {
          // lanes with equal (lane & 7) hold the same columns for different rows: fold them, fixed order
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            s4.x += __shfl_xor_sync(0xffffffffu, s4.x, o); s4.y += __shfl_xor_sync(0xffffffffu, s4.y, o);
            s4.z += __shfl_xor_sync(0xffffffffu, s4.z, o); s4.w += __shfl_xor_sync(0xffffffffu, s4.w, o);
            q4.x += __shfl_xor_sync(0xffffffffu, q4.x, o); q4.y += __shfl_xor_sync(0xffffffffu, q4.y, o);
            q4.z += __shfl_xor_sync(0xffffffffu, q4.z, o); q4.w += __shfl_xor_sync(0xffffffffu, q4.w, o);
          }
          if (lane < 8) {
            float4* ps = reinterpret_cast<float4*>(stat + c0 + c4);
            float4* pq = reinterpret_cast<float4*>(stat + BN + c0 + c4);
            float4 t = st_old; t.x += s4.x; t.y += s4.y; t.z += s4.z; t.w += s4.w; __stcg(ps, t);
            t = sq_old; t.x += q4.x; t.y += q4.y; t.z += q4.z; t.w += q4.w; __stcg(pq, t);
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader + 8u * (uint32_t)acc);      // the leader's MMA thread counts both CTAs
    }
    if (kTmaStore && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // smem must outlive the last TMA stores
  } else if (warp == 0 && lane == 0 && rank == 0) {
    // ================= MMA issuer: one thread of CTA 0 for the pair =================
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(2 * BM, BN);
    const uint64_t desc0 = tc::smem_desc(tc::smem_u32(smem), LBO, SBO);
    int c = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int acc = lt & 1;
      if (lt >= 2) mbar_wait_cluster(&tempty[acc], ((lt >> 1) - 1) & 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)(acc * BN);
      uint32_t accum = 0;
      for (int kc = 0; kc < nk; ++kc, ++c) {
        const int sp = c % NSP;
        mbar_wait_cluster(&apl_full[sp], (c / NSP) & 1);
        tc::fence_after_sync();
        const uint64_t da = desc0 + (uint64_t)((OFF_APL + sp * 2 * A_PLANE) >> 4);
        const uint64_t db = desc0 + (uint64_t)((OFF_W + kc * 2 * W_PLANE) >> 4);
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t a_hi = da + (uint64_t)((j * 2 * LBO) >> 4);
          const uint64_t a_lo = a_hi + (uint64_t)(A_PLANE >> 4);
          const uint64_t b_hi = db + (uint64_t)((j * 2 * LBO) >> 4);
          const uint64_t b_lo = b_hi + (uint64_t)(W_PLANE >> 4);
          mma2_bf16_ss(d_tmem, a_lo, b_hi, IDESC, accum);   // small terms first (same order as the one-CTA kernels)
          mma2_bf16_ss(d_tmem, a_hi, b_lo, IDESC, 1);
          mma2_bf16_ss(d_tmem, a_hi, b_hi, IDESC, 1);
          accum = 1;
        }
        mma2_commit_both(&apl_empty[sp]);
      }
      mma2_commit_both(&tfull[acc]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  cluster_sync_all();                         // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 0) tmem_dealloc2(tmem, TMEM_COLS);
}

static std::atomic<int> g_pair_enabled{0};   // opt-in: correct on B200 but slower than the one-CTA kernel at these shapes

int launch_pair(const CUtensorMap& mapA, const CUtensorMap& mapC, const Params& p, cudaStream_t st) {
  static alignn::DeviceOnce configured; int cfg_dev;
  if (configured.needed(&cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_gather_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return record_cuda_error((int)e);
    configured.done(cfg_dev);
  }
  const int total = (p.M + 2 * BM - 1) / (2 * BM);
  const int pairs = total < 74 ? total : 74;
  gemm_gather_pair_kernel<false><<<2 * pairs, THREADS, SMEM, st>>>(mapA, mapC, p);
  return check_launch();
}

}  // namespace gemmp
}  // namespace alignn

extern "C" {

__attribute__((visibility("hidden"))) int alignn_b200_gemm_pair_enabled() { return alignn::gemmp::g_pair_enabled.load(); }
void alignn_b200_debug_gemm_pair(int enabled) { alignn::gemmp::g_pair_enabled.store(enabled); }   /* A/B switch, not in the public header */

/* called by alignn_b200_gemm_gather (gemm_fused_tc.cu) for the shapes the pair kernel covers; returns 1 if it took the call */
__attribute__((visibility("hidden"))) int alignn_b200_gemm_gather_try_pair(const alignn_b200_gemm_gather_args* a, int* status) {
  using namespace alignn::gemmp;
  if (!g_pair_enabled.load() || a->N != BN || a->K > MAX_NK * BK || a->K % BK || a->bn_scale || a->M < 2 * BM) return 0;
  CUtensorMap mapA, mapC;
  int rc = alignn::gemm2::make_map_f32(&mapA, a->A, a->M, a->K, a->lda, BM);
  if (rc == ALIGNN_OK) rc = alignn::gemm2::make_map_f32(&mapC, a->C, a->M, a->N, a->ldc, 32);
  if (rc != ALIGNN_OK) { *status = rc; return 1; }
  Params p;
  p.M = (int)a->M; p.N = a->N; p.K = a->K;
  p.w_image = reinterpret_cast<const uint8_t*>(a->w_image);
  p.bias = a->bias;
  p.add0 = a->add0; p.ld0 = a->ld0; p.idx0 = a->idx0;
  p.add1 = a->add1; p.ld1 = a->ld1; p.idx1 = a->idx1;
  p.C = a->C; p.ldc = a->ldc; p.stats = a->stats;
  p.bn_scale = p.bn_shift = p.bn_mean = nullptr;
  *status = launch_pair(mapA, mapC, p, (cudaStream_t)a->stream);
  return 1;
}

}  // extern "C"
