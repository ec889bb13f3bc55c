// Scale-pyramid plan + area-resize/normalise of every level in one launch.
//
// Replaces detect_face.py:46 (u8 NHWC -> fp32 NCHW copy of the whole batch, never materialised here), :48-60 (scale
// list), :71-72 + :304-306 (imresample(area) per scale + (x-127.5)*0.0078125).  The u8 frame is the only input: every
// level pixel is the mean of its adaptive-average-pooling window [floor(i*H/oh), ceil((i+1)*H/oh)) computed as an exact
// integer sum followed by sum/kh/kw in fp32 -- bit-identical to torch's CPU adaptive_avg_pool2d on integer pixels.
// One task per output row (see pyramid_rows_kernel); tasks are ordered frame-major so that the 9-14 passes over one
// frame hit L2 and HBM sees each frame once.
#include "common.cuh"
#include <math.h>

extern long long g_vnfr_launches;

extern "C" int vnfr_pyramid_plan(int B, int H, int W, int min_face_size, double factor, VnfrPyramid* out) {
  VNFR_REQUIRE(out != nullptr && B >= 0 && H > 0 && W > 0 && min_face_size > 0 && factor > 0 && factor < 1, "bad arguments");
  memset(out, 0, sizeof(*out));
  out->B = B; out->H = H; out->W = W;
  // detect_face.py:50-60, Python doubles
  const double m = 12.0 / (double)min_face_size;
  double minl = (double)(H < W ? H : W) * m;
  double scale_i = m;
  int n = 0;
  while (minl >= 12) {
    VNFR_REQUIRE(n < VNFR_MAX_LEVELS, "too many pyramid levels");
    out->scale_d[n] = scale_i;
    out->scale[n] = (float)scale_i;
    ++n;
    scale_i = scale_i * factor;
    minl = minl * factor;
  }
  out->n_levels = n;
  int64_t loff = 0, moff = 0, poff = 0;
  int toff = 0;
  for (int l = 0; l < n; ++l) {
    const int lh = (int)((double)H * out->scale_d[l] + 1), lw = (int)((double)W * out->scale_d[l] + 1);   // :71
    out->lh[l] = lh; out->lw[l] = lw;
    // conv3 -> maxpool2 (ceil) -> conv3 -> conv3 (mtcnn.py:38-45)
    const int ph = (lh - 2 + 1) / 2, pw = (lw - 2 + 1) / 2;
    const int oh = ph - 4, ow = pw - 4;
    out->oh[l] = oh > 0 ? oh : 0; out->ow[l] = ow > 0 ? ow : 0;
    out->level_off[l] = loff; loff += (int64_t)B * 3 * lh * lw;
    out->map_off[l] = moff; moff += (int64_t)B * out->oh[l] * out->ow[l];
    out->px_off[l] = poff; poff += (int64_t)lh * lw;
    out->tiles_x[l] = (out->ow[l] + 15) / 16; out->tiles_y[l] = (out->oh[l] + 15) / 16;
    out->tile_off[l] = toff; toff += out->tiles_x[l] * out->tiles_y[l];
  }
  out->level_off[n] = loff; out->map_off[n] = moff; out->px_off[n] = poff; out->tile_off[n] = toff;
  return VNFR_OK;
}

namespace {

constexpr int PYR_THREADS = 256;

struct PyrResizeParams {
  int B, H, W, n_levels;
  int lh[VNFR_MAX_LEVELS], lw[VNFR_MAX_LEVELS];
  long long level_off[VNFR_MAX_LEVELS];
  int row_off[VNFR_MAX_LEVELS + 1];     // prefix sum of lh: task r of a frame is output row r - row_off[l] of level l
};

// Packed accumulation of 4 source bytes: two u32 each holding two u16 lanes (bytes 0,2 and bytes 1,3).
__device__ __forceinline__ void acc_word(uint32_t w, uint32_t& even, uint32_t& odd) {
  even += w & 0x00FF00FFu;
  odd += (w >> 8) & 0x00FF00FFu;
}

// One task = one output row (frame b, level l, row oy).  Its adaptive-average-pooling windows all share the source rows
// [y0, y1), which are ONE contiguous span of the u8 frame: the CTA streams that span with V-byte loads (V = 16 when the
// row pitch and base allow it), each thread keeping exact integer sums of its byte columns (u16 lanes, flushed to u32
// shared memory every 256 rows), then every output pixel adds up its kw column sums.  Arithmetic of the final
// conversion is the reference's: sum / kh / kw in fp32, then (x - 127.5) * 0.0078125 (detect_face.py:72, :305).
template <int V>
__global__ void __launch_bounds__(PYR_THREADS) pyramid_rows_kernel(const __grid_constant__ PyrResizeParams p,
                                                                   const uint8_t* __restrict__ frames, float* __restrict__ levels) {
  extern __shared__ uint32_t colsum[];          // [W*3] exact column sums of the current task
  const int rows_per_frame = p.row_off[p.n_levels];
  const long long n_tasks = (long long)rows_per_frame * p.B;
  const int rowbytes = p.W * 3;
  const int n_chunks = rowbytes / V;
  for (long long task = blockIdx.x; task < n_tasks; task += gridDim.x) {
    const int b = (int)(task / rows_per_frame);
    const int r = (int)(task - (long long)b * rows_per_frame);
    int l = 0;
    while (l + 1 < p.n_levels && r >= p.row_off[l + 1]) ++l;
    const int lw = p.lw[l], lh = p.lh[l];
    const int oy = r - p.row_off[l];
    const int y0 = (int)(((long long)oy * p.H) / lh), y1 = (int)((((long long)oy + 1) * p.H + lh - 1) / lh);
    const uint8_t* src = frames + ((size_t)b * p.H + y0) * rowbytes;
    for (int c = threadIdx.x; c < n_chunks; c += PYR_THREADS) {
      uint32_t tot[V];
#pragma unroll
      for (int j = 0; j < V; ++j) tot[j] = 0;
      for (int yb = y0; yb < y1; yb += 256) {
        const int ye = min(y1, yb + 256);
        const uint8_t* q = src + (size_t)(yb - y0) * rowbytes + (size_t)c * V;
        if (V == 16) {
          uint32_t ev[4] = {0, 0, 0, 0}, od[4] = {0, 0, 0, 0};
          int y = yb;
          for (; y + 4 <= ye; y += 4) {
            uint4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = __ldg(reinterpret_cast<const uint4*>(q + (size_t)u * rowbytes));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc_word(w[u].x, ev[0], od[0]); acc_word(w[u].y, ev[1], od[1]);
              acc_word(w[u].z, ev[2], od[2]); acc_word(w[u].w, ev[3], od[3]);
            }
            q += (size_t)4 * rowbytes;
          }
          for (; y < ye; ++y) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(q));
            acc_word(w.x, ev[0], od[0]); acc_word(w.y, ev[1], od[1]);
            acc_word(w.z, ev[2], od[2]); acc_word(w.w, ev[3], od[3]);
            q += rowbytes;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tot[4 * k + 0] += ev[k] & 0xFFFFu; tot[4 * k + 1] += od[k] & 0xFFFFu;
            tot[4 * k + 2] += ev[k] >> 16;     tot[4 * k + 3] += od[k] >> 16;
          }
        } else {
          for (int y = yb; y < ye; ++y) {
#pragma unroll
            for (int j = 0; j < V; ++j) tot[j] += __ldg(q + j);
            q += rowbytes;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < V; ++j) colsum[c * V + j] = tot[j];
    }
    __syncthreads();
    const float kh = (float)(y1 - y0);
    const size_t plane = (size_t)lh * lw;
    float* orow = levels + p.level_off[l] + (size_t)b * 3 * plane + (size_t)oy * lw;
    for (int i = threadIdx.x; i < 3 * lw; i += PYR_THREADS) {
      const int ch = i / lw, ox = i - ch * lw;
      // 32-bit unsigned division (ox * W < 2^31 is checked on the host): a 64-bit division costs ~8x as much
      const int x0 = (int)(((unsigned)ox * (unsigned)p.W) / (unsigned)lw);
      const int x1 = (int)((((unsigned)ox + 1u) * (unsigned)p.W + (unsigned)lw - 1u) / (unsigned)lw);
      uint32_t s = 0;
      for (int x = x0; x < x1; ++x) s += colsum[3 * x + ch];
      orow[ch * plane + ox] = mul_rn(sub_rn(div_rn(div_rn((float)s, kh), (float)(x1 - x0)), 127.5f), 0.0078125f);
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int vnfr_pyramid_resize_norm(const VnfrPyramid* pyr, const uint8_t* frames, float* levels, void* stream) {
  VNFR_REQUIRE(pyr != nullptr, "pyramid plan is null");
  if (pyr->B == 0 || pyr->n_levels == 0) return VNFR_OK;
  PyrResizeParams p;
  p.B = pyr->B; p.H = pyr->H; p.W = pyr->W; p.n_levels = pyr->n_levels;
  int roff = 0;
  for (int l = 0; l < pyr->n_levels; ++l) {
    p.lh[l] = pyr->lh[l]; p.lw[l] = pyr->lw[l]; p.level_off[l] = pyr->level_off[l];
    p.row_off[l] = roff; roff += pyr->lh[l];
  }
  p.row_off[pyr->n_levels] = roff;
  const long long n_tasks = (long long)roff * p.B;
  const size_t smem = (size_t)p.W * 3 * sizeof(uint32_t);
  VNFR_REQUIRE(smem <= 200 * 1024, "frame too wide for the pyramid kernel (W*3*4 bytes of shared memory needed)");
  VNFR_REQUIRE((long long)p.W * p.W < (1ll << 31) && (long long)p.H * p.H < (1ll << 31), "frame too large");
  const bool vec = ((size_t)p.W * 3) % 16 == 0 && ((uintptr_t)frames % 16) == 0;
  const int per_sm = smem > 0 ? (int)((220 * 1024) / (smem + 1024)) : 8;
  long long grid = 148LL * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  if (grid > n_tasks) grid = n_tasks;
  static bool attr = false;
  if (!attr) {
    VNFR_CUDA(cudaFuncSetAttribute(pyramid_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(pyramid_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  if (vec) pyramid_rows_kernel<16><<<(int)grid, PYR_THREADS, smem, (cudaStream_t)stream>>>(p, frames, levels);
  else pyramid_rows_kernel<1><<<(int)grid, PYR_THREADS, smem, (cudaStream_t)stream>>>(p, frames, levels);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
