// Machine characterisation for the design of the fused kernels (not part of the product):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench tools/microbench.cu
//   gpurun -- ./tools/microbench
// Prints one JSON line per experiment: L2-resident and HBM read bandwidth with plain loads, TMA tile loads from HBM,
// TMA tile::gather4 row gathers from an L2-resident table (several box widths), TMA tile stores.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

// ---- 1/2: plain vector loads, every CTA strides over `n4` float4 (L2-resident if small, HBM if large) ----
__global__ void __launch_bounds__(512) read_kernel(const float4* __restrict__ p, size_t n4, int reps, float* sink) {
  float acc = 0.f;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      const float4 v = __ldcg(p + i);
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456f) *sink = acc;
}

// ---- 3: TMA tile loads [rows x cols] fp32 boxes from a [R x 256] matrix, STAGES-deep ring, nobody reads the data ----
template <int STAGES>
__global__ void __launch_bounds__(128) tma_tile_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int box_cols, int tiles,
                                                       int chunks_per_tile) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  const uint32_t bytes = (uint32_t)box_rows * box_cols * 4;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    int issued = 0, waited = 0;
    const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * chunks_per_tile;
    for (; issued < total && issued < STAGES; ++issued) {
      const int tile = blockIdx.x + (issued / chunks_per_tile) * gridDim.x, kc = issued % chunks_per_tile;
      mbar_expect(&full[issued % STAGES], bytes);
      tma_load_2d(smem + (issued % STAGES) * bytes, &map, kc * box_cols, tile * box_rows, &full[issued % STAGES]);
    }
    for (; waited < total; ++waited) {
      mbar_wait(&full[waited % STAGES], (waited / STAGES) & 1);
      if (issued < total) {
        const int tile = blockIdx.x + (issued / chunks_per_tile) * gridDim.x, kc = issued % chunks_per_tile;
        mbar_expect(&full[issued % STAGES], bytes);
        tma_load_2d(smem + (issued % STAGES) * bytes, &map, kc * box_cols, tile * box_rows, &full[issued % STAGES]);
        ++issued;
      }
    }
  }
}

// ---- 4: gather4: per "chunk" NG gather4 instructions (4 rows x box_cols each) land on one mbarrier; STAGES-deep ----
template <int STAGES>
__global__ void __launch_bounds__(128) tma_gather_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ rows, int box_cols,
                                                         int ng, int chunks, int col_chunks, float* check) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  const uint32_t inst_bytes = 4u * box_cols * 4u, bytes = inst_bytes * ng;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x < 32) {
    // one warp issues: lane l issues instruction l, l+32, ... of the chunk (row indices pre-loaded)
    int issued = 0, waited = 0;
    auto issue = [&](int c) {
      const int s = c % STAGES;
      if (threadIdx.x == 0) mbar_expect(&full[s], bytes);
      __syncwarp();
      const int* rp = rows + ((size_t)blockIdx.x * chunks + c) * ng * 4;
      const int col = (c % col_chunks) * box_cols;
      for (int g = threadIdx.x; g < ng; g += 32) {
        const int4 r = *reinterpret_cast<const int4*>(rp + g * 4);
        tma_gather4(smem + s * bytes + g * inst_bytes, &map, col, r.x, r.y, r.z, r.w, &full[s]);
      }
    };
    for (; issued < chunks && issued < STAGES; ++issued) issue(issued);
    for (; waited < chunks; ++waited) {
      mbar_wait(&full[waited % STAGES], (waited / STAGES) & 1);
      if (check && blockIdx.x == 0 && waited == 0 && threadIdx.x == 0) {   // layout check: first floats of each of the 4 rows of instr 0
        const float* f = reinterpret_cast<const float*>(smem);
        for (int j = 0; j < 4; ++j) { check[2 * j] = f[j * box_cols]; check[2 * j + 1] = f[j * box_cols + box_cols - 1]; }
      }
      __syncwarp();
      if (issued < chunks) { issue(issued); ++issued; }
    }
  }
}

// ---- 5: TMA tile stores of [rows x cols] boxes from smem ----
__global__ void __launch_bounds__(128) tma_store_kernel(const __grid_constant__ CUtensorMap map, int box_rows, int box_cols, int tiles,
                                                        int chunks_per_tile) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < box_rows * box_cols; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int kc = 0; kc < chunks_per_tile; ++kc) {
        tma_store_2d(&map, smem, kc * box_cols, tile * box_rows);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (EncodeFn)fn;
}
static CUtensorMap make_map(EncodeFn enc, void* base, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows,
                            CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 4};
  cuuint32_t box[2] = {box_cols, box_rows}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

template <class F>
static float time_ms(F f, int iters = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  EncodeFn enc = get_encode();
  float* sink; CK(cudaMalloc(&sink, 64));
  const int SMS = 148;
  // 1/2 plain loads
  {
    const size_t big = (size_t)2 << 30, small = (size_t)32 << 20;
    float4* p; CK(cudaMalloc(&p, big)); CK(cudaMemset(p, 0, big));
    float ms = time_ms([&] { read_kernel<<<SMS * 4, 512>>>(p, small / 16, 40, sink); });
    printf("{\"exp\": \"ldg_l2_resident_32MB\", \"GBps\": %.1f}\n", small * 40 / ms / 1e6);
    ms = time_ms([&] { read_kernel<<<SMS * 4, 512>>>(p, big / 16, 1, sink); });
    printf("{\"exp\": \"ldg_hbm_2GB\", \"GBps\": %.1f}\n", big / ms / 1e6);
    CK(cudaFree(p));
  }
  // 3 TMA tile loads from HBM: [276480 x 256] fp32 (283 MB), boxes 128 x 32 (16 KB)
  {
    const uint64_t R = 276480, C = 256;
    float* y; CK(cudaMalloc(&y, R * C * 4)); CK(cudaMemset(y, 0, R * C * 4));
    for (int bc : {32, 64}) {
      CUtensorMap m = make_map(enc, y, R, C, bc, 128);
      const int tiles = R / 128, cpt = C / bc;
      const size_t smem = (size_t)4 * 128 * bc * 4;
      CK(cudaFuncSetAttribute(tma_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_ms([&] { tma_tile_kernel<4><<<SMS, 128, smem>>>(m, 128, bc, tiles, cpt); });
      printf("{\"exp\": \"tma_tile_load_hbm\", \"box\": \"128x%d\", \"stages\": 4, \"GBps\": %.1f, \"us\": %.1f}\n", bc, R * C * 4 / ms / 1e6, ms * 1e3);
      const size_t smem8 = (size_t)8 * 128 * bc * 4;
      if (smem8 <= 227 * 1024) {
        CK(cudaFuncSetAttribute(tma_tile_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
        ms = time_ms([&] { tma_tile_kernel<8><<<SMS, 128, smem8>>>(m, 128, bc, tiles, cpt); });
        printf("{\"exp\": \"tma_tile_load_hbm\", \"box\": \"128x%d\", \"stages\": 8, \"GBps\": %.1f, \"us\": %.1f}\n", bc, R * C * 4 / ms / 1e6, ms * 1e3);
      }
    }
    // 5 TMA stores
    {
      CUtensorMap m = make_map(enc, y, R, C, 32, 128);
      const size_t smem = 128 * 32 * 4;
      float ms = time_ms([&] { tma_store_kernel<<<SMS, 128, smem>>>(m, 128, 32, (int)(R / 128), 8); });
      printf("{\"exp\": \"tma_tile_store\", \"box\": \"128x32\", \"GBps\": %.1f, \"us\": %.1f}\n", R * C * 4 / ms / 1e6, ms * 1e3);
    }
    CK(cudaFree(y));
  }
  // 4 gather4 from an L2-resident table P [23040 x 1024] fp32 (94 MB): 128 random rows per chunk
  {
    const uint64_t R = 23040, C = 1024;
    float* P; CK(cudaMalloc(&P, R * C * 4));
    std::vector<float> h(R * C);
    for (uint64_t r = 0; r < R; ++r) for (uint64_t c = 0; c < C; ++c) h[r * C + c] = (float)r + (float)c * 1e-4f;
    CK(cudaMemcpy(P, h.data(), R * C * 4, cudaMemcpyHostToDevice));
    float* chk; CK(cudaMalloc(&chk, 64));
    for (int bc : {32, 64, 128, 256}) {
      const int ng = 32, chunks = 15 * (256 / bc) * 2, col_chunks = 256 / bc;   // ~15 tiles per SM, e_src and Bh slices
      std::vector<int> rows((size_t)SMS * chunks * ng * 4);
      uint32_t s = 12345;
      for (size_t i = 0; i < rows.size(); ++i) {
        s = s * 1664525u + 1013904223u;
        const size_t cta = i / ((size_t)chunks * ng * 4);
        rows[i] = (int)((cta * 150 + (s >> 8) % 360) % R);     // rows of "one crystal": locality like L(g)
      }
      int* d_rows; CK(cudaMalloc(&d_rows, rows.size() * 4));
      CK(cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
      CUtensorMap m = make_map(enc, P, R, C, bc, 1);
      const size_t smem = (size_t)4 * ng * 4 * bc * 4;
      if (smem > 227 * 1024) { CK(cudaFree(d_rows)); continue; }
      CK(cudaFuncSetAttribute(tma_gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      float ms = time_ms([&] { tma_gather_kernel<4><<<SMS, 128, smem>>>(m, d_rows, bc, ng, chunks, col_chunks, chk); });
      float hc[8]; CK(cudaMemcpy(hc, chk, 32, cudaMemcpyDeviceToHost));
      const double bytes = (double)SMS * chunks * ng * 4 * bc * 4;
      printf("{\"exp\": \"tma_gather4_l2\", \"box_cols\": %d, \"GBps\": %.1f, \"us\": %.1f, \"Minstr_per_s\": %.1f, \"rows0\": [%d,%d,%d,%d], "
             "\"smem_first_last\": [%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f]}\n",
             bc, bytes / ms / 1e6, ms * 1e3, (double)SMS * chunks * ng / ms / 1e3, rows[0], rows[1], rows[2], rows[3], hc[0], hc[1], hc[2], hc[3],
             hc[4], hc[5], hc[6], hc[7]);
      CK(cudaFree(d_rows));
    }
    CK(cudaFree(P));
  }
  return 0;
}
