// Linear layer with a gather-add epilogue on the 5th-gen tensor cores -- the edge-gate kernel of the conv path:
//
//   C[r, :] = A[r, :] * W^T (+ bias) (+ add0[i0(r), off0:off0+N]) (+ add1[i1(r), off1:off1+N])        r < M
//   stats[0/1][c] = sum_r C[r, c], sum_r C[r, c]^2                                                    (optional)
//
// With A = edge features y, add0 = P[src, 0:d] (e_src), add1 = P[dst, 2d:3d] (e_dst + both biases) this is the
// pre-activation gate  m = e_src[src] + e_dst[dst] + edge_gate(y)  of alignn/models/alignn.py:98-101 in ONE pass over y
// -- the product  edge_gate(y)  never exists in HBM -- plus the per-channel batch statistics BatchNorm1d(m) needs
// (alignn.py:123).  With i0 = identity and add0 = the incoming gradient it is the data-gradient GEMM of the backward
// (residual in the epilogue); with no addends it is a plain Linear (node projections, embedding MLPs).
//
// Data movement (B200, sm_100a):
//   A     fp32 [128 x 32] boxes by TMA tensor tiles (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 4-deep staging ring;
//         the next tile's boxes are pulled into L2 by cp.async.bulk.prefetch.tensor one tile ahead.  Converter warps
//         split the staged tile into bf16 hi/lo planes in the canonical K-major UMMA layout (bf16x3, tc_common.cuh).
//   W     pre-split bf16 hi/lo image (gemm_prepare_weights), one cp.async.bulk per K chunk, 3-deep ring.
//   acc   tcgen05.mma kind::f16 into a double-buffered TMEM accumulator (2 x BN columns).
//   out   epilogue warps: tcgen05.ld -> warp-private smem transpose -> 8 lanes per row, 4 rows per instruction:
//         + bias + gathered addend rows (coalesced 128-byte row pieces by index, software-pipelined one chunk ahead
//         in registers; measured: TMA tile::gather4 sustains only ~3.2 G instructions/s chip-wide = 1.7-3.3 TB/s for
//         128/256-byte pieces, plain coalesced loads from L2 17 TB/s -- tools/microbench.cu), column statistics,
//         128-byte-per-row coalesced stores.
// Warp roles: 0 = TMEM owner + MMA issuer, 1 = A producer (TMA), 2 = W producer, 4-7 = epilogue, 8-11 = converters.
#include <atomic>
#include <cuda.h>

#include "tc_common.cuh"
#include "api_common.h"
#include "alignn_b200.h"

namespace alignn {
namespace gemm2 {

int make_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int NSA = 3;          // fp32 A staging ring (TMA destination); L2 prefetch one tile ahead covers the HBM latency
constexpr int NSP = 2;          // bf16 A plane ring
constexpr int NSB = 3;          // W chunk ring
constexpr int EPI_WARPS = 8;          // two per TMEM lane quarter, alternating 32-column chunks
constexpr int CONV_WARPS = 4;
constexpr int THREADS = 32 * (4 + EPI_WARPS + CONV_WARPS);   // 384
constexpr uint32_t LBO = 128;
constexpr uint32_t SBO = (BK / 8) * 128;
constexpr int A_STAGE = BM * BK * 4;          // 16 KB fp32 box
constexpr int A_PLANE = BM * BK * 2;          // 8 KB bf16 plane
constexpr int EPI_BOX = 32 * 32 * 4;          // one warp's [32 rows x 32 floats] transpose box (128-byte swizzle)
constexpr bool kTmaStore = true;              // true: finished boxes leave by TMA tile stores (costs 2 more passes over shared
                                              // memory per tile; the kernel is bound by shared-memory bandwidth)
constexpr int EPI_BUFS = 1;

template <int BN>
struct Cfg {
  static constexpr int B_PLANE = BN * BK * 2;
  static constexpr int OFF_ASTG = 0;                                   // 1024-byte aligned boxes
  static constexpr int OFF_APL = OFF_ASTG + NSA * A_STAGE;
  static constexpr int OFF_B = OFF_APL + NSP * 2 * A_PLANE;
  static constexpr int OFF_EPI = (OFF_B + NSB * 2 * B_PLANE + 1023) / 1024 * 1024;
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_BUFS * EPI_BOX;
  static constexpr int OFF_STAT = OFF_EPI + EPI_BYTES;                 // [EPI_WARPS][2][BN] floats
  static constexpr int STAT_BYTES = 4 * 2 * BN * 4;                    // per lane quarter (its two warps own disjoint columns)
  static constexpr int OFF_BAR = OFF_STAT + STAT_BYTES;
  static constexpr int SMEM = OFF_BAR + 256;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static_assert(SMEM <= 232448, "shared memory budget of one sm_100 CTA");
};

struct Params {
  int M, N, K;
  const uint8_t* w_image;
  const float* bias;
  const float* add0; int64_t ld0; const int32_t* idx0;   // addend rows: add0[idx0 ? idx0[r] : r][0 .. N)  (column offset folded in)
  const float* add1; int64_t ld1; const int32_t* idx1;
  float* C; int64_t ldc;
  float* stats;                                          // [stat_quarters * gridDim.x][2][N] or NULL (requires N == BN)
  int stat_quarters;                                     // 1: one partial row per CTA; 4: one per (CTA, lane quarter)
  const float* bn_scale; const float* bn_shift; const float* bn_mean;   // != NULL: add1 rows are NOT added; they are the
                                                         // pre-norm rows m of the BatchNorm+SiLU that produced this GEMM's
                                                         // input gradient, and stats = sum gu, sum gu (m - mean)
  long long* trace;                                      // development aid: per-role event clocks of CTA 0 ([6][512]) or NULL
};

// role r of CTA 0 appends the SM clock to its row of the trace buffer (tools/trace_gemm.py prints the timeline)
#define GEMM2_TRACE(role, ctr) do { if (p.trace && blockIdx.x == 0 && (ctr) < 512) p.trace[(role) * 512 + (ctr)++] = clock64(); } while (0)

__host__ __device__ constexpr int plane_off(int r, int k) { return (r >> 3) * (int)SBO + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ float dsilu_(float u) {          // d/du [u * sigmoid(u)], same formula as egc_kernels.cu
  float e, sg;                                             // 4-instruction sigmoid, as common.cuh's sigmoidf_
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + e));
  return sg * (1.f + u * (1.f - sg));
}
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

// MODE 0: bias / gathered addends / column statistics in the epilogue's row pass; 1: BatchNorm-backward sums in that pass;
// 2: nothing row-wise (bias only, added in registers): accumulator -> swizzled box -> TMA store, no row pass;
// 3: mode 0 with the first addend only and no statistics (data gradient + residual): half the addend registers.
template <int BN, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_gather_bf16x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapC, const Params p) {
  using F = Cfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* astg_full = reinterpret_cast<uint64_t*>(smem + F::OFF_BAR);
  uint64_t* astg_empty = astg_full + NSA;
  uint64_t* apl_full = astg_empty + NSA;
  uint64_t* apl_empty = apl_full + NSP;
  uint64_t* b_full = apl_empty + NSP;
  uint64_t* b_empty = b_full + NSB;
  uint64_t* tfull = b_empty + NSB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = p.K / BK;
  const int n_tiles = p.N / BN;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int total = m_tiles * n_tiles;
  const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { tc::mbar_init(&astg_full[s], 1); tc::mbar_init(&astg_empty[s], CONV_WARPS); }
    for (int s = 0; s < NSP; ++s) { tc::mbar_init(&apl_full[s], CONV_WARPS); tc::mbar_init(&apl_empty[s], 1); }
    for (int s = 0; s < NSB; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], EPI_WARPS); }
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 1) {
    // ================= A producer: TMA boxes of [128 rows x 32 floats] =================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      int c = 0, tr = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        const int m0 = (tile / n_tiles) * BM;
        // the next tile's A boxes -> L2, one tile ahead of the loads that will want them
        if (lt + 1 < my_tiles) {
          const int nm0 = ((tile + (int)gridDim.x) / n_tiles) * BM;
          if (nm0 != m0)
            for (int kc = 0; kc < nk; ++kc) tma_prefetch_2d(&mapA, kc * BK, nm0);
        }
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % NSA;
          if (c >= NSA) tc::mbar_wait(&astg_empty[s], ((c / NSA) - 1) & 1);
          GEMM2_TRACE(0, tr);
          tc::mbar_arrive_expect_tx(&astg_full[s], A_STAGE);
          tma_load_2d(smem + F::OFF_ASTG + s * A_STAGE, &mapA, kc * BK, m0, &astg_full[s]);
        }
      }
    }
  } else if (warp == 2) {
    // ================= W producer: one bulk copy per K chunk (both planes) =================
    if (lane == 0) {
      int c = 0, trb = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        const uint8_t* wsrc = p.w_image + (int64_t)(tile % n_tiles) * nk * 2 * F::B_PLANE;
        for (int kc = 0; kc < nk; ++kc, ++c) {
          const int s = c % NSB;
          if (c >= NSB) tc::mbar_wait(&b_empty[s], ((c / NSB) - 1) & 1);
          GEMM2_TRACE(1, trb);
          tc::mbar_arrive_expect_tx(&b_full[s], 2 * F::B_PLANE);
          tc::bulk_g2s(smem + F::OFF_B + s * 2 * F::B_PLANE, wsrc, 2 * F::B_PLANE, &b_full[s]);
          wsrc += 2 * F::B_PLANE;
        }
      }
    }
  } else if (warp >= 4 + EPI_WARPS) {
    // ================= converters: staged fp32 box -> bf16 hi/lo planes =================
    // unit u = (iteration, warp, half-warp) -> 8 rows x 2 adjacent 16-byte chunks.  A quarter-warp reads one chunk
    // column of 8 consecutive rows (distinct banks under the 128-byte swizzle); a half-warp writes one 128-byte core
    // matrix column of a plane (conflict-free 64-bit stores).
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
    int tr2 = 0, tr3 = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int sa = c % NSA, sp = c % NSP;
      tc::mbar_wait(&astg_full[sa], (c / NSA) & 1);
      if (cw == 0 && lane == 0) GEMM2_TRACE(2, tr2);
      const uint8_t* src = smem + F::OFF_ASTG + sa * A_STAGE;
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(src + ld_off[i]);
      if (c >= NSP) tc::mbar_wait(&apl_empty[sp], ((c / NSP) - 1) & 1);
      uint8_t* dst = smem + F::OFF_APL + sp * 2 * A_PLANE;
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
        tc::mbar_arrive(&apl_full[sp]);
        // the staged box is released only now: the plane stores above consumed every loaded value.  (Releasing right
        // after issuing the shared-memory loads let the TMA refill overtake loads still queued in the LSU: stale
        // 8-row groups, seen on hardware.)
        tc::mbar_arrive(&astg_empty[sa]);
      }
      if (cw == 0 && lane == 0) GEMM2_TRACE(3, tr3);
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int ew = warp - 4;
    const int q = warp & 3;                                  // TMEM lane quarter = rows 32q .. 32q+31 of the tile
    uint8_t* stg0 = smem + F::OFF_EPI + ew * EPI_BUFS * EPI_BOX;
    // staged element (row r, 16-byte chunk j) lives at r * 128 + ((j ^ (r & 7)) << 4): what a SWIZZLE_128B tensor map
    // expects, and conflict-free for both the row-per-thread writes and the 8-lanes-per-row reads
    const int st_row = lane * 128, st_sw = lane & 7;
    uint32_t chunk_ctr = 0;
    float* stat = reinterpret_cast<float*>(smem + F::OFF_STAT) + q * 2 * BN;
    const int half = ew >> 2;                                // which of the quarter's two warps: chunks half, half + 2, ...
    constexpr bool bnmode = MODE == 1;
    constexpr bool plain = MODE == 2 && kTmaStore;
    constexpr bool one = MODE == 3;                                       // add0 only, no add1, no statistics
    constexpr bool late_wait = (plain || one) && kTmaStore;               // enough registers to hold a chunk across the wait
    const bool do_stats = !plain && !one && p.stats != nullptr;
    if (do_stats)
      for (int ch = half; ch < BN / 32; ch += 2) { stat[ch * 32 + lane] = 0.f; stat[BN + ch * 32 + lane] = 0.f; }
    const int rsub = lane >> 3;                              // row within a group of 4
    const int c4 = (lane & 7) * 4;                           // 4 of the chunk's 32 columns
    constexpr int NCH = BN / 32;
    uint32_t lt = 0;
    int tr5 = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      // rows this lane finishes: it * 4 + rsub, it < 8.  i0 / i1 = addend row of each (-1: no addend / row past M)
      int i0[8], i1[8];
      uint32_t rvm = 0;
      const float* base0 = p.add0 + n0 + c4 + half * 32;       // first chunk of this warp
      const float* base1 = p.add1 + n0 + c4 + half * 32;
      float4 a0[8], a1[8];
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (!plain) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int gr = m0 + q * 32 + it * 4 + rsub;
          const bool ok = gr < p.M;
          rvm |= (ok ? 1u : 0u) << it;
          i0[it] = (p.add0 && ok) ? (p.idx0 ? __ldg(p.idx0 + gr) : gr) : -1;
          if constexpr (!one) i1[it] = (p.add1 && ok) ? (p.idx1 ? __ldg(p.idx1 + gr) : gr) : -1;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          a0[it] = (i0[it] >= 0 && half < NCH) ? __ldg(reinterpret_cast<const float4*>(base0 + (int64_t)i0[it] * p.ld0)) : z4;
          if constexpr (!one)
            a1[it] = (i1[it] >= 0 && half < NCH) ? __ldg(reinterpret_cast<const float4*>(base1 + (int64_t)i1[it] * p.ld1)) : z4;
        }
      }
      tc::mbar_wait(&tfull[acc], (lt >> 1) & 1);
      tc::fence_after_sync();
      if (ew == 0 && lane == 0) GEMM2_TRACE(5, tr5);
#pragma unroll 1
      for (int ch = half; ch < NCH; ch += 2) {
        const int c0 = ch * 32;
        uint8_t* stg = stg0;
        // the TMA store that read this box must be done with it.  Modes 2 and 3 wait only after the TMEM load (the store's
        // read of the box overlaps it); the others have no registers to keep the 32 values alive across the wait
        if (!late_wait && kTmaStore && chunk_ctr >= 1) {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
        {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
          if (ew == 0 && lane == 0) GEMM2_TRACE(5, tr5);
          if (late_wait && chunk_ctr >= 1) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          ++chunk_ctr;
          if constexpr (plain) {
            // nothing row-wise to add: bias (the same 32 values for every lane) in registers, then straight to the store
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + 4 * j));
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + st_row + ((j ^ st_sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if constexpr (plain) {
          tc::fence_async_smem();
          __syncwarp();
          if (ew == 0 && lane == 0) GEMM2_TRACE(5, tr5);
          if (lane == 0) {
            tma_store_2d(&mapC, stg, n0 + c0, m0 + q * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          __syncwarp();
          continue;
        }
        __syncwarp();
        if (ew == 0 && lane == 0) GEMM2_TRACE(5, tr5);
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
          float4 mrow = z4;                                 // (BatchNorm mode: the pre-norm row, not an addend)
          if constexpr (!one) mrow = a1[it];
          const float4 ad1 = bnmode ? z4 : mrow;
          o.x = (o.x + b4.x) + (a0[it].x + ad1.x);
          o.y = (o.y + b4.y) + (a0[it].y + ad1.y);
          o.z = (o.z + b4.z) + (a0[it].z + ad1.z);
          o.w = (o.w + b4.w) + (a0[it].w + ad1.w);
          // this row's addends of the NEXT chunk go out now and land while the rest of this chunk is processed
          if (more) {
            if (i0[it] >= 0) a0[it] = __ldg(reinterpret_cast<const float4*>(base0 + (int64_t)i0[it] * p.ld0 + (c0 - half * 32) + 64));
            if constexpr (!one)
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
            float4 t = *ps; t.x += s4.x; t.y += s4.y; t.z += s4.z; t.w += s4.w; *ps = t;
            t = *pq; t.x += q4.x; t.y += q4.y; t.z += q4.z; t.w += q4.w; *pq = t;
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (ew == 0 && lane == 0) GEMM2_TRACE(5, tr5);
    }
    if (kTmaStore && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // smem must outlive the last TMA stores
    if (do_stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      const float* all = reinterpret_cast<const float*>(smem + F::OFF_STAT);
      if (p.stat_quarters == 4) {
        // one partial row per (CTA, lane quarter): [4 * gridDim.x][2][BN] (the pair kernel's layout, when it is enabled)
        float* out_rows = p.stats + (int64_t)blockIdx.x * 4 * 2 * BN;
        for (int i = ew * 32 + lane; i < 4 * 2 * BN; i += EPI_WARPS * 32) out_rows[i] = all[i];
      } else {
        // one partial row per CTA: the four lane quarters summed in a fixed order
        float* out_row = p.stats + (int64_t)blockIdx.x * 2 * BN;
        for (int i = ew * 32 + lane; i < 2 * BN; i += EPI_WARPS * 32)
          out_row[i] = (all[i] + all[2 * BN + i]) + (all[4 * BN + i] + all[6 * BN + i]);
      }
    }
  } else if (warp == 0 && lane == 0) {
    // ================= MMA issuer (one thread) =================
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(BM, BN);
    const uint64_t desc0 = tc::smem_desc(tc::smem_u32(smem), LBO, SBO);
    int c = 0, tr4 = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int acc = lt & 1;
      if (lt >= 2) tc::mbar_wait(&tempty[acc], ((lt >> 1) - 1) & 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)(acc * BN);
      uint32_t accum = 0;
      for (int kc = 0; kc < nk; ++kc, ++c) {
        const int sp = c % NSP, sb = c % NSB;
        tc::mbar_wait(&apl_full[sp], (c / NSP) & 1);
        tc::mbar_wait(&b_full[sb], (c / NSB) & 1);
        tc::fence_after_sync();
        GEMM2_TRACE(4, tr4);
        const uint64_t da = desc0 + (uint64_t)((F::OFF_APL + sp * 2 * A_PLANE) >> 4);
        const uint64_t db = desc0 + (uint64_t)((F::OFF_B + sb * 2 * F::B_PLANE) >> 4);
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          const uint64_t a_hi = da + (uint64_t)((j * 2 * LBO) >> 4);
          const uint64_t a_lo = a_hi + (uint64_t)(A_PLANE >> 4);
          const uint64_t b_hi = db + (uint64_t)((j * 2 * LBO) >> 4);
          const uint64_t b_lo = b_hi + (uint64_t)(F::B_PLANE >> 4);
          tc::mma_bf16_ss(d_tmem, a_lo, b_hi, IDESC, accum);   // small terms first (same order as gemm_tc.cu)
          tc::mma_bf16_ss(d_tmem, a_hi, b_lo, IDESC, 1);
          tc::mma_bf16_ss(d_tmem, a_hi, b_hi, IDESC, 1);
          accum = 1;
        }
        tc::mma_commit(&apl_empty[sp]);
        tc::mma_commit(&b_empty[sb]);
      }
      tc::mma_commit(&tfull[acc]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, F::TMEM_COLS);
}

// ---- host side: tensor map for A (driver entry point fetched through the runtime, no link-time libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static std::atomic<EncodeTiledFn> cached{nullptr};
  EncodeTiledFn f = cached.load(std::memory_order_acquire);
  if (f) return f;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return nullptr;
  f = reinterpret_cast<EncodeTiledFn>(sym);
  cached.store(f, std::memory_order_release);
  return f;
}

// row-major fp32 [rows, cols] with row stride ld (floats): boxes of [box_rows x 32 floats], 128-byte swizzle
int make_map_f32(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return record_cuda_error((int)cudaErrorSymbolNotFound);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return record_cuda_error(1000 + (int)r);
  return ALIGNN_OK;
}

static long long* g_trace = nullptr;   // set by alignn_b200_debug_gemm_trace (development aid, not in the public header)

inline int pick_bn(int N) { return (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : (N % 32 == 0) ? 32 : 0; }

template <int BN, int MODE>
int launch2(const CUtensorMap& mapA, const CUtensorMap& mapC, const Params& p, cudaStream_t st) {
  using F = Cfg<BN>;
  static alignn::DeviceOnce configured; int cfg_dev;
  if (configured.needed(&cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_gather_bf16x3_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM);
    if (e != cudaSuccess) return record_cuda_error((int)e);
    configured.done(cfg_dev);
  }
  const int total = ((p.M + BM - 1) / BM) * (p.N / BN);
  int grid = total < 148 ? total : 148;
  if (p.stats) grid = alignn_b200_gemm_gather_stat_rows(p.M, p.N) / p.stat_quarters;   // the caller sized `stats` for this many
                                                                      // CTAs (a CTA without tiles writes zeros)
  gemm_gather_bf16x3_kernel<BN, MODE><<<grid, THREADS, F::SMEM, st>>>(mapA, mapC, p);
  return check_launch();
}

template <int BN>
int launch(const CUtensorMap& mapA, const CUtensorMap& mapC, const Params& p, cudaStream_t st) {
  if (p.bn_scale) return launch2<BN, 1>(mapA, mapC, p, st);
  if (kTmaStore && !p.add0 && !p.add1 && !p.stats) return launch2<BN, 2>(mapA, mapC, p, st);
  if (p.add0 && !p.add1 && !p.stats) return launch2<BN, 3>(mapA, mapC, p, st);
  return launch2<BN, 0>(mapA, mapC, p, st);
}

}  // namespace gemm2
}  // namespace alignn

extern "C" {

void alignn_b200_debug_gemm_trace(long long* device_buffer) { alignn::gemm2::g_trace = device_buffer; }

__attribute__((visibility("hidden"))) int alignn_b200_gemm_pair_enabled();   /* gemm_pair_tc.cu */

int alignn_b200_gemm_gather_stat_rows(int64_t M, int N) {
  const int bn = alignn::gemm2::pick_bn(N);
  if (bn == 0 || M <= 0) return 0;
  // CTAs of the launch (statistics need N == bn: one column tile): the same count for the one-CTA kernel and for the
  // pair kernel (two CTAs per 256-row tile)
  const int64_t total = 2 * ((M + 2 * alignn::gemm2::BM - 1) / (2 * alignn::gemm2::BM));
  return (alignn_b200_gemm_pair_enabled() ? 4 : 1) * (int)(total < 148 ? total : 148);   // pair kernel: four lane quarters per CTA
}

__attribute__((visibility("hidden"))) int alignn_b200_gemm_gather_try_pair(const alignn_b200_gemm_gather_args* a, int* status);   /* gemm_pair_tc.cu */

int alignn_b200_gemm_gather(const alignn_b200_gemm_gather_args* a) {
  using namespace alignn::gemm2;
  if (!a) return ALIGNN_ERR_BAD_ARG;
  if (a->struct_size != sizeof(*a)) return ALIGNN_ERR_STRUCT_SIZE;
  if (a->M < 0 || a->N <= 0 || a->K <= 0 || a->K % BK != 0 || a->lda < a->K || a->ldc < a->N) return ALIGNN_ERR_BAD_ARG;
  if (a->M == 0) return ALIGNN_OK;
  if (!a->A || !a->w_image || !a->C || a->M > 0x7fffffff) return ALIGNN_ERR_BAD_ARG;
  if ((a->lda % 4) || (a->ldc % 4) || ((uintptr_t)a->A & 15) || ((uintptr_t)a->C & 15)) return ALIGNN_ERR_BAD_ARG;
  if (a->add0 && ((a->ld0 % 4) || a->ld0 < a->N || ((uintptr_t)a->add0 & 15))) return ALIGNN_ERR_BAD_ARG;
  if (a->add1 && ((a->ld1 % 4) || a->ld1 < a->N || ((uintptr_t)a->add1 & 15))) return ALIGNN_ERR_BAD_ARG;
  if ((a->idx0 && !a->add0) || (a->idx1 && !a->add1)) return ALIGNN_ERR_BAD_ARG;
  const int bn = pick_bn(a->N);
  if (bn == 0) return ALIGNN_ERR_UNSUPPORTED_D;
  if (a->stats && a->N != bn) return ALIGNN_ERR_BAD_ARG;      // column statistics need the whole row in one tile
  if (a->bn_scale && (!a->bn_shift || !a->bn_mean || !a->add1 || !a->stats)) return ALIGNN_ERR_BAD_ARG;
  {
    int st = ALIGNN_OK;
    if (alignn_b200_gemm_gather_try_pair(a, &st)) return st;     // N = 256, K <= 256: the two-CTA kernel (gemm_pair_tc.cu)
  }
  CUtensorMap mapA, mapC;
  int rc = make_map_f32(&mapA, a->A, a->M, a->K, a->lda, BM);
  if (rc != ALIGNN_OK) return rc;
  rc = make_map_f32(&mapC, a->C, a->M, a->N, a->ldc, 32);
  if (rc != ALIGNN_OK) return rc;
  Params p;
  p.M = (int)a->M; p.N = a->N; p.K = a->K;
  p.w_image = reinterpret_cast<const uint8_t*>(a->w_image);
  p.bias = a->bias;
  p.add0 = a->add0; p.ld0 = a->ld0; p.idx0 = a->idx0;
  p.add1 = a->add1; p.ld1 = a->ld1; p.idx1 = a->idx1;
  p.C = a->C; p.ldc = a->ldc; p.stats = a->stats; p.stat_quarters = alignn_b200_gemm_pair_enabled() ? 4 : 1;
  p.bn_scale = a->bn_scale; p.bn_shift = a->bn_shift; p.bn_mean = a->bn_mean;
  p.trace = g_trace;
  cudaStream_t st = (cudaStream_t)a->stream;
  switch (bn) {
    case 256: return launch<256>(mapA, mapC, p, st);
    case 128: return launch<128>(mapA, mapC, p, st);
    case 64: return launch<64>(mapA, mapC, p, st);
    default: return launch<32>(mapA, mapC, p, st);
  }
}

}  // extern "C"
