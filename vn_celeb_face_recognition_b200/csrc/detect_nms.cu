// Generic segmented NMS entry point (vnfr_nms_segments): one CTA per segment, sort + greedy NMS in shared memory.
// The detector's fused stage kernels (detect_stages.cu) use the same device functions; this entry point exists so the
// NMS arithmetic and keep ORDER can be checked bit-for-bit against torchvision.ops.nms / nms_numpy on arbitrary inputs.
#include "nms.cuh"

extern long long g_vnfr_launches;

namespace {

template <int MODE>
__global__ void __launch_bounds__(512) nms_segments_kernel(int cap, const int* __restrict__ count, const float4* __restrict__ boxes,
                                                           const float* __restrict__ scores, float thr, int* __restrict__ keep_count,
                                                           int* __restrict__ keep) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int seg = blockIdx.x;
  const int n = min(count[seg], cap);
  const int np2 = next_pow2(max(n, 2));
  unsigned long long* key = reinterpret_cast<unsigned long long*>(smem);
  float4* sb = reinterpret_cast<float4*>(key + np2);
  float* sa = reinterpret_cast<float*>(sb + np2);
  uint32_t* val = reinterpret_cast<uint32_t*>(sa + np2);
  int* kept = reinterpret_cast<int*>(val + np2);
  __shared__ NmsScratch sc;
  const float4* gb = boxes + (size_t)seg * cap;
  const float* gs = scores + (size_t)seg * cap;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    if (i < n) {
      // mode 0: ties visit the lower index first (stable descending sort); mode 1: ascending stable sort read from
      // the end, i.e. ties visit the HIGHER index first.
      key[i] = nms_key(gs[i], MODE == 0 ? (uint32_t)i : 0xFFFFFFFFu - (uint32_t)i);
      val[i] = (uint32_t)i;
    } else {
      key[i] = ~0ull;
      val[i] = 0xFFFFFFFFu;
    }
  }
  __syncthreads();
  block_bitonic_sort(key, val, np2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 b = gb[val[i]];
    sb[i] = b;
    sa[i] = nms_area<MODE>(b);
  }
  __syncthreads();
  const int nk = block_nms_sorted<MODE>(sb, sa, n, thr, kept, &sc);
  for (int i = threadIdx.x; i < nk; i += blockDim.x) keep[(size_t)seg * cap + i] = (int)val[kept[i]];
  if (threadIdx.x == 0) keep_count[seg] = nk;
}

}  // namespace

extern "C" int vnfr_nms_segments(int n_segments, int cap, const int32_t* count, const float* boxes, const float* scores,
                                 float threshold, int mode, int32_t* keep_count, int32_t* keep, void* stream) {
  VNFR_REQUIRE(cap >= 1 && cap <= 4096, "cap must be in [1, 4096]");
  VNFR_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (IoU) or 1 (Min)");
  if (n_segments == 0) return VNFR_OK;
  const int np2 = next_pow2(cap < 2 ? 2 : cap);
  const size_t smem = (size_t)np2 * (8 + 16 + 4 + 4 + 4);
  auto kern = mode == 0 ? nms_segments_kernel<0> : nms_segments_kernel<1>;
  VNFR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<n_segments, 512, smem, (cudaStream_t)stream>>>(cap, count, reinterpret_cast<const float4*>(boxes), scores, threshold,
                                                        keep_count, keep);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
