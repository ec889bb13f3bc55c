// L2 -> SM operand supply: every CTA (one per SM) streams the SAME 2 MiB "weight" matrix through a shared-memory ring with TMA,
// (a) unicast, (b) in clusters of C CTAs where each CTA loads 1/C of every tile and multicasts it to all C.  No compute.
// Prints delivered bytes per clock per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mcast_bench tools/mcast_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int TILE_ROWS = 256, TILE_BYTES = TILE_ROWS * 128, STAGES = 4;   // 32 KB tiles (256 rows x 64 fp16), 128 KB ring

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(bar), "r"(ph) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// arrive on the barrier at the same offset in CTA `dst` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t dst) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(bar), "r"(dst));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}

template <int C>
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int n_tiles, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem), bars = base + STAGES * TILE_BYTES;
  const uint32_t rank = C > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 64 + 8 * s, C); }   // full, empty (C consumers)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (C > 1) cluster_sync();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {                                   // producer
    int s = 0; uint32_t ph = 1;
    for (int it = 0; it < iters; ++it)
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(bars + 64 + 8 * s, ph);
        mbar_expect(bars + 8 * s, TILE_BYTES);
        if (C == 1) {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(base + s * TILE_BYTES), "l"(&tm), "r"(bars + 8 * s), "r"(0), "r"(t * TILE_ROWS) : "memory");
        } else {
          const int rows = TILE_ROWS / C;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                       ::"r"(base + s * TILE_BYTES + rank * rows * 128), "l"(&tm), "r"(bars + 8 * s), "r"(0), "r"(t * TILE_ROWS + (int)rank * rows),
                         "h"((uint16_t)((1 << C) - 1)) : "memory");
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {                           // consumer: releases the stage in every CTA of the cluster
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it)
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(bars + 8 * s, ph);
        if (C == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 64 + 8 * s) : "memory");
        else for (uint32_t d = 0; d < C; ++d) mbar_arrive_remote(bars + 64 + 8 * s, d);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
  }
  __syncthreads();
  if (C > 1) cluster_sync();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int C>
void run(EncFn enc, void* w, int rows_total, int grid, long long* d_out) {
  CUtensorMap tm;
  const cuuint64_t dims[2] = {64, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)(TILE_ROWS / C)};
  const cuuint32_t es[2] = {1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); exit(1); }
  const int smem = STAGES * TILE_BYTES + 256, n_tiles = rows_total / TILE_ROWS, iters = 8;
  CK(cudaFuncSetAttribute(stream_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaLaunchKernelEx(&cfg, stream_kernel<C>, tm, n_tiles, iters, d_out));
    CK(cudaDeviceSynchronize());
  }
  long long h[256];
  CK(cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  long long mx = 0; double avg = 0;
  for (int i = 0; i < grid; ++i) { if (h[i] > mx) mx = h[i]; avg += (double)h[i] / grid; }
  const double bytes = (double)n_tiles * iters * TILE_BYTES;
  printf("cluster %d, grid %3d: %.1f B/clk/SM delivered (slowest CTA), %.1f (mean)\n", C, grid, bytes / mx, bytes / avg);
}

int main() {
  CK(cudaSetDevice(0));
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncFn enc = (EncFn)sym;
  const int rows_total = 16384;                              // 2 MiB: L2 resident
  void* w; CK(cudaMalloc(&w, (size_t)rows_total * 128)); CK(cudaMemset(w, 1, (size_t)rows_total * 128));
  long long* d_out; CK(cudaMalloc(&d_out, 256 * sizeof(long long)));
  run<1>(enc, w, rows_total, 148, d_out);
  run<1>(enc, w, rows_total, 32, d_out);
  run<2>(enc, w, rows_total, 148, d_out);
  run<4>(enc, w, rows_total, 148, d_out);
  run<8>(enc, w, rows_total, 144, d_out);
  return 0;
}
