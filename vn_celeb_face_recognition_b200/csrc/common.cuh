// Shared device/host helpers for the sm_100a kernels of the detect -> align -> embed -> classify path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vnfr_b200.h"

#define VNFR_CHECK_LAUNCH()                                                  \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      vnfr_set_error(__FILE__, __LINE__, cudaGetErrorString(e__));           \
      return VNFR_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

#define VNFR_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      vnfr_set_error(__FILE__, __LINE__, cudaGetErrorString(e__));           \
      return VNFR_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

#define VNFR_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) {                                                           \
      vnfr_set_error(__FILE__, __LINE__, msg);                               \
      return VNFR_ERR_ARG;                                                   \
    }                                                                        \
  } while (0)

void vnfr_set_error(const char* file, int line, const char* msg);

// Individually rounded fp32 ops: the reference computes box arithmetic as separate tensor ops, so no FMA contraction.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One-time setup that is per DEVICE (cudaFuncSetAttribute, __device__ tables): true the first time it is asked on the
// current device.  (A process-wide `static bool` would leave a second GPU of the same process unconfigured.)
struct VnfrPerDevice { bool done[64]; };
inline bool vnfr_first_on_device(VnfrPerDevice& s) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
  if (s.done[d]) return false;
  s.done[d] = true;
  return true;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
