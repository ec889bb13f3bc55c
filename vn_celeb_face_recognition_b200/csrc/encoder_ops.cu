// Memory-bound companions of the tensor-core convolutions of the encoder / classifier: pooling, layout adapter,
// L2 normalisation and log-softmax/argmax.  All are one pass over their input with 16-byte accesses.
#include "common.cuh"
#include <math_constants.h>
#include <cuda_fp16.h>

extern long long g_vnfr_launches;

namespace {

// 16-bit storage helpers: F16 = true -> IEEE half, false -> bfloat16
template <bool F16>
__device__ __forceinline__ float h_lo(uint32_t w) {
  return F16 ? __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))) : __uint_as_float(w << 16);
}
template <bool F16>
__device__ __forceinline__ float h_hi(uint32_t w) {
  return F16 ? __half2float(__ushort_as_half((unsigned short)(w >> 16))) : __uint_as_float(w & 0xFFFF0000u);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  if (F16) {
    a = fminf(fmaxf(a, -65504.f), 65504.f); b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <bool F16>
__device__ __forceinline__ unsigned short pack_h1(float a) {
  if (F16) return __half_as_ushort(__float2half_rn(fminf(fmaxf(a, -65504.f), 65504.f)));
  return __bfloat16_as_ushort(__float2bfloat16_rn(a));
}

// element-wise maximum of two packed 16-bit pairs
template <bool F16>
__device__ __forceinline__ uint32_t hmax2_16(uint32_t a, uint32_t b) {
  if (F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}

// MaxPool2d(3, stride=2), floor mode, no padding (inception_resnet_v1.py:147, :179, :224).  One thread = 8 channels.
template <bool F16>
__global__ void maxpool3s2_kernel(const __nv_bfloat16* __restrict__ in, int n_img, int in_h, int in_w, int c8, int in_pitch,
                                  __nv_bfloat16* __restrict__ out, int out_h, int out_w, int out_pitch) {
  const long long total = (long long)n_img * out_h * out_w * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    long long t = i / c8;
    const int ox = (int)(t % out_w); t /= out_w;
    const int oy = (int)(t % out_h);
    const int img = (int)(t / out_h);
    // packed 16-bit maxima (HMNMX2): all nine 16-byte loads are issued before the first use; converting to fp32 and back cost
    // four times the instructions for the same (exact) result
    uint4 v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const size_t px = ((size_t)img * in_h + (2 * oy + ky)) * in_w + (2 * ox + kx);
        v[ky * 3 + kx] = __ldg(reinterpret_cast<const uint4*>(in + px * in_pitch + cg * 8));
      }
    uint4 m = v[0];
#pragma unroll
    for (int t = 1; t < 9; ++t) {
      m.x = hmax2_16<F16>(m.x, v[t].x); m.y = hmax2_16<F16>(m.y, v[t].y);
      m.z = hmax2_16<F16>(m.z, v[t].z); m.w = hmax2_16<F16>(m.w, v[t].w);
    }
    const size_t opx = ((size_t)img * out_h + oy) * out_w + ox;
    *reinterpret_cast<uint4*>(out + opx * out_pitch + cg * 8) = m;
  }
}

// AdaptiveAvgPool2d(1): mean over hw pixels in fp32 (sum then one division), bf16 out.  One thread = 8 channels.
template <bool F16>
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ in, int n_img, int hw, int c8, int in_pitch,
                               __nv_bfloat16* __restrict__ out) {
  const int total = n_img * c8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cg = i % c8, img = i / c8;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = 0; p < hw; ++p) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)img * hw + p) * in_pitch + cg * 8));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[2 * e] += h_lo<F16>(w[e]); s[2 * e + 1] += h_hi<F16>(w[e]); }
    }
    const float inv = 1.0f / (float)hw;
    *reinterpret_cast<uint4*>(out + (size_t)img * c8 * 8 + cg * 8) =
        make_uint4(pack_h2<F16>(s[0] * inv, s[1] * inv), pack_h2<F16>(s[2] * inv, s[3] * inv), pack_h2<F16>(s[4] * inv, s[5] * inv),
                   pack_h2<F16>(s[6] * inv, s[7] * inv));
  }
}

// fp32 NCHW (3 planes) -> bf16 NHWC, 8 channels per pixel (3 real + 5 zeros) = one 16-byte store per pixel.
template <bool F16>
__global__ void nchw3_to_nhwc8_kernel(const float* __restrict__ in, int n_img, int hw, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)n_img * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw, p = i % hw;
    const float* b = in + img * 3 * hw + p;
    const float r = __ldg(b), g = __ldg(b + hw), bl = __ldg(b + 2 * hw);
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(pack_h2<F16>(r, g), pack_h2<F16>(bl, 0.f), 0u, 0u);
  }
}

// fp32 NCHW (3 planes) -> 16-bit space-to-depth [n][h2][w2][16]: one thread per 2x2 cell = two 16-byte stores.
template <bool F16>
__global__ void nchw3_to_s2d16_kernel(const float* __restrict__ in, int n_img, int h, int w, __nv_bfloat16* __restrict__ out) {
  const int h2 = (h + 1) >> 1, w2 = (w + 1) >> 1;
  const long long total = (long long)n_img * h2 * w2;
  const long long hw = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / (h2 * w2);
    const int rem = (int)(i - img * (h2 * w2));
    const int cy = rem / w2, cx = rem - cy * w2;
    uint32_t wds[8];
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int y = 2 * cy + (sub >> 1), x = 2 * cx + (sub & 1);
      float r = 0.f, g = 0.f, bl = 0.f;
      if (y < h && x < w) {
        const float* b = in + img * 3 * hw + (long long)y * w + x;
        r = __ldg(b); g = __ldg(b + hw); bl = __ldg(b + 2 * hw);
      }
      wds[2 * sub] = pack_h2<F16>(r, g);
      wds[2 * sub + 1] = pack_h2<F16>(bl, 0.f);
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    o[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
  }
}

// u8 HWC faces (what the reference's glue hands to transforms_default, data_loader/__init__.py:27-34, 52-56) -> the
// standardised 16-bit space-to-depth encoder input [n][ceil(h/2)][ceil(w/2)][16]: (x - 127.5) / 128, one thread per 2x2 cell.
template <bool F16>
__global__ void u8hwc_to_s2d16_kernel(const uint8_t* __restrict__ in, int n_img, int h, int w, __nv_bfloat16* __restrict__ out) {
  const int h2 = (h + 1) >> 1, w2 = (w + 1) >> 1;
  const long long total = (long long)n_img * h2 * w2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / (h2 * w2);
    const int rem = (int)(i - img * (h2 * w2));
    const int cy = rem / w2, cx = rem - cy * w2;
    uint32_t wds[8];
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int y = 2 * cy + (sub >> 1), x = 2 * cx + (sub & 1);
      float r = 0.f, g = 0.f, bl = 0.f;
      if (y < h && x < w) {
        const uint8_t* b = in + ((img * h + y) * (long long)w + x) * 3;
        r = ((float)__ldg(b) - 127.5f) * 0.0078125f; g = ((float)__ldg(b + 1) - 127.5f) * 0.0078125f; bl = ((float)__ldg(b + 2) - 127.5f) * 0.0078125f;
      }
      wds[2 * sub] = pack_h2<F16>(r, g);
      wds[2 * sub + 1] = pack_h2<F16>(bl, 0.f);
    }
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    o[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// F.normalize(p=2, dim=1, eps=1e-12): x / max(||x||, eps).  One warp per row.
template <bool F16>
__global__ void l2norm_kernel(const float* __restrict__ x, int n, int d, int x_pitch, float* __restrict__ emb,
                              unsigned short* __restrict__ emb_half) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = x + (size_t)row * x_pitch;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) { const float v = xr[i]; ss += v * v; }
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int i = lane; i < d; i += 32) {
    const float v = xr[i] / denom;
    emb[(size_t)row * d + i] = v;
    if (emb_half != nullptr) emb_half[(size_t)row * d + i] = pack_h1<F16>(v);
  }
}

// log_softmax over the first c columns + argmax (first maximal index, like torch.argmax) + exp(max log-prob).
__global__ void logsoftmax_argmax_kernel(const float* __restrict__ logits, int n, int c, int pitch, float* __restrict__ logp,
                                         long long* __restrict__ label, float* __restrict__ prob) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* xr = logits + (size_t)row * pitch;
  float mx = -CUDART_INF_F;
  int arg = 0x7fffffff;
  for (int i = lane; i < c; i += 32) {
    const float v = xr[i];
    if (v > mx) { mx = v; arg = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.f;
  for (int i = lane; i < c; i += 32) se += expf(xr[i] - mx);
  se = warp_sum(se);
  const float lse = logf(se);
  if (logp != nullptr)
    for (int i = lane; i < c; i += 32) logp[(size_t)row * c + i] = xr[i] - mx - lse;
  if (lane == 0) {
    if (label != nullptr) label[row] = arg;
    if (prob != nullptr) prob[row] = expf(-lse);
  }
}

// Row-wise top-k (k <= 8) of a score matrix, largest first, ties towards the lower index.  One CTA per row: every thread
// keeps a sorted local list over its strided columns, then k rounds of a block-wide arg-max pop the global winners.
// With `accumulate` the row's previous (out_val, out_idx) entries are candidates too, so a gallery larger than one score
// buffer is searched tile by tile.
constexpr int TOPK_MAX = 8, TOPK_THREADS = 256;

__global__ void __launch_bounds__(TOPK_THREADS) topk_rows_kernel(const float* __restrict__ scores, int g, int pitch, int k,
                                                                 int col_offset, int accumulate, float* __restrict__ out_val,
                                                                 int* __restrict__ out_idx) {
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xr = scores + (size_t)row * pitch;
  float v[TOPK_MAX];
  int ix[TOPK_MAX];
#pragma unroll
  for (int j = 0; j < TOPK_MAX; ++j) { v[j] = -CUDART_INF_F; ix[j] = 0x7fffffff; }
  auto insert = [&](float x, int i) {
    // sorted insert (value descending, index ascending)
    if (x > v[TOPK_MAX - 1] || (x == v[TOPK_MAX - 1] && i < ix[TOPK_MAX - 1])) {
      v[TOPK_MAX - 1] = x; ix[TOPK_MAX - 1] = i;
#pragma unroll
      for (int j = TOPK_MAX - 1; j > 0; --j) {
        const bool sw = v[j] > v[j - 1] || (v[j] == v[j - 1] && ix[j] < ix[j - 1]);
        if (sw) { const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv; const int ti = ix[j]; ix[j] = ix[j - 1]; ix[j - 1] = ti; }
      }
    }
  };
  for (int i = tid; i < g; i += TOPK_THREADS) insert(xr[i], col_offset + i);
  if (accumulate && tid < k) insert(out_val[(size_t)row * k + tid], out_idx[(size_t)row * k + tid]);
  __shared__ float s_v[TOPK_THREADS / 32];
  __shared__ int s_i[TOPK_THREADS / 32], s_owner[TOPK_THREADS / 32];
  __shared__ int s_win_owner;
  for (int r = 0; r < k; ++r) {
    // every thread proposes the head of its list
    float bv = v[0];
    int bi = ix[0], bo = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), oo = __shfl_xor_sync(0xffffffffu, bo, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bo = oo; }
    }
    if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; s_owner[warp] = bo; }
    __syncthreads();
    if (tid == 0) {
      float wv = s_v[0]; int wi = s_i[0], wo = s_owner[0];
      for (int w = 1; w < TOPK_THREADS / 32; ++w)
        if (s_v[w] > wv || (s_v[w] == wv && s_i[w] < wi)) { wv = s_v[w]; wi = s_i[w]; wo = s_owner[w]; }
      out_val[(size_t)row * k + r] = wv;
      out_idx[(size_t)row * k + r] = wi;
      s_win_owner = wo;
    }
    __syncthreads();
    if (tid == s_win_owner) {                 // pop the winner's head
#pragma unroll
      for (int j = 0; j < TOPK_MAX - 1; ++j) { v[j] = v[j + 1]; ix[j] = ix[j + 1]; }
      v[TOPK_MAX - 1] = -CUDART_INF_F; ix[TOPK_MAX - 1] = 0x7fffffff;
    }
    __syncthreads();
  }
}

// BGR <-> RGB on packed u8 pixels (cv2.cvtColor(frame, COLOR_BGR2RGB), demo_video.py:107-110): one thread = 4 pixels = three
// 32-bit words, bytes shuffled with PRMT; in == out is allowed.
__global__ void swap_rb_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long n_quads) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_quads; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t w0 = in[3 * i], w1 = in[3 * i + 1], w2 = in[3 * i + 2];
    out[3 * i] = __byte_perm(w0, w1, 0x5012);
    out[3 * i + 1] = __byte_perm(__byte_perm(w1, w0, 0x3070), w2, 0x3410);
    out[3 * i + 2] = __byte_perm(w2, w1, 0x1236);
  }
}

// NV12 (what a hardware video decoder delivers: H x W luma plane + H/2 x W interleaved U,V) -> packed RGB u8, BT.601 limited
// range in OpenCV's fixed point (cv2.cvtColor(..., COLOR_YUV2RGB_NV12), imgproc color_yuv: CY 1220542, CVR 1673527, CVG -852492,
// CUG -409993, CUB 2116026, shift 20): bit-identical to cv2, at half the host -> device bytes of RGB frames.  One thread = 2x2 px.
__global__ void nv12_to_rgb_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n_img, int h, int w) {
  const int h2 = h >> 1, w2 = w >> 1;
  const long long total = (long long)n_img * h2 * w2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / (h2 * w2);
    const int rem = (int)(i - img * (h2 * w2)), cy = rem / w2, cx = rem - cy * w2;
    const uint8_t* base = in + img * ((long long)h * w * 3 / 2);
    const uint8_t* uvp = base + (long long)h * w + (long long)cy * w + 2 * cx;
    const int u = (int)__ldg(uvp) - 128, v = (int)__ldg(uvp + 1) - 128;
    const int ruv = (1 << 19) + 1673527 * v, guv = (1 << 19) - 852492 * v - 409993 * u, buv = (1 << 19) + 2116026 * u;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int y = 2 * cy + (sub >> 1), x = 2 * cx + (sub & 1);
      const int yy = max(0, (int)__ldg(base + (long long)y * w + x) - 16) * 1220542;
      uint8_t* o = out + ((img * h + y) * (long long)w + x) * 3;
      o[0] = (uint8_t)min(max((yy + ruv) >> 20, 0), 255);
      o[1] = (uint8_t)min(max((yy + guv) >> 20, 0), 255);
      o[2] = (uint8_t)min(max((yy + buv) >> 20, 0), 255);
    }
  }
}

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int vnfr_maxpool3s2_nhwc(const void* in, int n_img, int in_h, int in_w, int c, int in_pitch, void* out,
                                    int out_pitch, int dtype, void* stream) {
  VNFR_REQUIRE(c % 8 == 0 && in_pitch % 8 == 0 && out_pitch % 8 == 0, "channels and pitches must be multiples of 8");
  VNFR_REQUIRE(in_h >= 3 && in_w >= 3, "input smaller than the pooling window");
  const int out_h = (in_h - 3) / 2 + 1, out_w = (in_w - 3) / 2 + 1;
  const long long total = (long long)n_img * out_h * out_w * (c / 8);
  if (total == 0) return VNFR_OK;
  auto kern = dtype == 1 ? maxpool3s2_kernel<true> : maxpool3s2_kernel<false>;
  kern<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, n_img, in_h, in_w, c / 8, in_pitch,
                                                             (__nv_bfloat16*)out, out_h, out_w, out_pitch);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_avgpool_nhwc(const void* in, int n_img, int hw, int c, int in_pitch, void* out, int dtype, void* stream) {
  VNFR_REQUIRE(c % 8 == 0 && in_pitch % 8 == 0, "channels and pitch must be multiples of 8");
  if (n_img == 0) return VNFR_OK;
  auto kern = dtype == 1 ? avgpool_kernel<true> : avgpool_kernel<false>;
  kern<<<grid_for((long long)n_img * (c / 8), 128), 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, n_img, hw, c / 8,
                                                                                  in_pitch, (__nv_bfloat16*)out);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_nchw3_to_nhwc8(const float* in, int n_img, int h, int w, void* out, int dtype, void* stream) {
  if (n_img == 0) return VNFR_OK;
  auto kern = dtype == 1 ? nchw3_to_nhwc8_kernel<true> : nchw3_to_nhwc8_kernel<false>;
  kern<<<grid_for((long long)n_img * h * w, 256), 256, 0, (cudaStream_t)stream>>>(in, n_img, h * w, (__nv_bfloat16*)out);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_nchw3_to_s2d16(const float* in, int n_img, int h, int w, void* out, int dtype, void* stream) {
  if (n_img == 0) return VNFR_OK;
  auto kern = dtype == 1 ? nchw3_to_s2d16_kernel<true> : nchw3_to_s2d16_kernel<false>;
  kern<<<grid_for((long long)n_img * ((h + 1) / 2) * ((w + 1) / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, n_img, h, w,
                                                                                                    (__nv_bfloat16*)out);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_l2norm_rows(const float* x, int n, int d, int x_pitch, float* emb, void* emb_half, int dtype, void* stream) {
  if (n == 0) return VNFR_OK;
  auto kern = dtype == 1 ? l2norm_kernel<true> : l2norm_kernel<false>;
  kern<<<ceil_div(n, 4), 128, 0, (cudaStream_t)stream>>>(x, n, d, x_pitch, emb, (unsigned short*)emb_half);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_logsoftmax_argmax(const float* logits, int n, int c, int pitch, float* logp, int64_t* label, float* prob,
                                      void* stream) {
  if (n == 0) return VNFR_OK;
  logsoftmax_argmax_kernel<<<ceil_div(n, 4), 128, 0, (cudaStream_t)stream>>>(logits, n, c, pitch, logp, (long long*)label, prob);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_topk_rows(const float* scores, int n, int g, int pitch, int k, int col_offset, int accumulate, float* out_val,
                              int32_t* out_idx, void* stream) {
  VNFR_REQUIRE(scores && out_val && out_idx, "null pointer");
  VNFR_REQUIRE(k >= 1 && k <= TOPK_MAX && g >= 0 && pitch >= g, "k must be in [1,8] and pitch >= g");
  if (n == 0) return VNFR_OK;
  topk_rows_kernel<<<n, TOPK_THREADS, 0, (cudaStream_t)stream>>>(scores, g, pitch, k, col_offset, accumulate, out_val, out_idx);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_swap_rb_u8(const uint8_t* in, uint8_t* out, long long n_pixels, void* stream) {
  VNFR_REQUIRE(in != nullptr && out != nullptr, "null pointer");
  VNFR_REQUIRE(n_pixels % 4 == 0 && ((uintptr_t)in % 4) == 0 && ((uintptr_t)out % 4) == 0, "pixel count must be a multiple of 4 and the buffers 4-byte aligned");
  if (n_pixels == 0) return VNFR_OK;
  swap_rb_kernel<<<grid_for(n_pixels / 4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(in),
                                                                            reinterpret_cast<uint32_t*>(out), n_pixels / 4);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_u8hwc_to_s2d16(const uint8_t* in, int n_img, int h, int w, void* out, int dtype, void* stream) {
  VNFR_REQUIRE(in != nullptr && out != nullptr, "null pointer");
  if (n_img == 0) return VNFR_OK;
  auto kern = dtype == 1 ? u8hwc_to_s2d16_kernel<true> : u8hwc_to_s2d16_kernel<false>;
  kern<<<grid_for((long long)n_img * ((h + 1) / 2) * ((w + 1) / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, n_img, h, w, (__nv_bfloat16*)out);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_nv12_to_rgb_u8(const uint8_t* in, uint8_t* out, int n_img, int h, int w, void* stream) {
  VNFR_REQUIRE(in != nullptr && out != nullptr, "null pointer");
  VNFR_REQUIRE(h % 2 == 0 && w % 2 == 0 && h > 0 && w > 0, "NV12 frames need even dimensions");
  if (n_img == 0) return VNFR_OK;
  nv12_to_rgb_kernel<<<grid_for((long long)n_img * (h / 2) * (w / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, out, n_img, h, w);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
