// Scratch microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, SS mode, operands resident in shared memory) as a
// function of N and of how many independent accumulators the issue order rotates over.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vn_celeb_face_recognition_b200/csrc/tc_common.cuh"
using namespace tc;
void vnfr_set_error(const char*, int, const char*) {}

__global__ void __launch_bounds__(128, 1) k(int n, int nacc, int iters, int kind_tf32, int row_bytes, int shift_rows, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + 16384u, bar = sb + 32768u, slot = bar + 8u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x < 32) {
    uint32_t idesc = make_idesc_f16(n, 1);
    if (kind_tf32) idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a0 = make_sw_desc(sa + (uint32_t)(shift_rows * row_bytes), row_bytes, 0), b0 = make_sw_desc(sb, row_bytes, 0);
    const int kmax = row_bytes / 32;
    const int stride = n < 32 ? 32 : n;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int kk4 = 0; kk4 < 4; ++kk4)
          for (int u = 0; u < nacc; ++u) {
            const int kk = kk4 % kmax;
            if (kind_tf32)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                           ::"r"(tmem + (uint32_t)(u * stride)), "l"(a0 + (uint64_t)(2 * kk)), "l"(b0 + (uint64_t)(2 * kk)), "r"(idesc), "r"(1u) : "memory");
            else
              umma_bf16(tmem + (uint32_t)(u * stride), a0 + (uint64_t)(2 * kk), b0 + (uint64_t)(2 * kk), idesc, 1u);
          }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int iters = 2000;
  for (int tf = 0; tf < 1; ++tf)
    for (int rb : {128, 64, 32})
     for (int shift : {0, 1, 3, 8})
      for (int n : {32, 64, 128})
        for (int nacc : {1, 4}) {
          const int grid = 148;
          if (nacc * (n < 32 ? 32 : n) > 512) continue;
          k<<<grid, 128, 60000>>>(n, nacc, iters, tf, rb, shift, d);
          long long c = 0;
          cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("row %3d B shift %d rows N %3d accumulators %d : %.1f cycles / mma\n", rb, shift, n, nacc, (double)c / (iters * 4.0 * nacc));
        }
  return 0;
}
