// Scale-pyramid plan + area-resize/normalise of every level in one launch.
//
// Replaces detect_face.py:46 (u8 NHWC -> fp32 NCHW copy of the whole batch, never materialised here), :48-60 (scale
// list), :71-72 + :304-306 (imresample(area) per scale + (x-127.5)*0.0078125).  The u8 frame is the only input: every
// level pixel is the mean of its adaptive-average-pooling window [floor(i*H/oh), ceil((i+1)*H/oh)) computed as an exact
// integer sum followed by sum/kh/kw in fp32 -- bit-identical to torch's CPU adaptive_avg_pool2d on integer pixels.
// Work is ordered frame-major so that the 9-14 passes over one frame hit L2, and HBM sees each frame once.
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

struct PyrResizeParams {
  int B, H, W, n_levels;
  int lh[VNFR_MAX_LEVELS], lw[VNFR_MAX_LEVELS];
  long long level_off[VNFR_MAX_LEVELS];
  long long px_off[VNFR_MAX_LEVELS + 1];
};

__global__ void __launch_bounds__(256) pyramid_resize_kernel(const __grid_constant__ PyrResizeParams p,
                                                             const uint8_t* __restrict__ frames, float* __restrict__ levels) {
  const long long per_img = p.px_off[p.n_levels];
  const long long total = per_img * p.B;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per_img);
    const long long q = idx - (long long)b * per_img;
    int l = 0;
    while (l + 1 < p.n_levels && q >= p.px_off[l + 1]) ++l;
    const int lw = p.lw[l], lh = p.lh[l];
    const int r = (int)(q - p.px_off[l]);
    const int oy = r / lw, ox = r - oy * lw;
    const int y0 = (int)(((long long)oy * p.H) / lh), y1 = (int)((((long long)oy + 1) * p.H + lh - 1) / lh);
    const int x0 = (int)(((long long)ox * p.W) / lw), x1 = (int)((((long long)ox + 1) * p.W + lw - 1) / lw);
    unsigned s0 = 0, s1 = 0, s2 = 0;
    const uint8_t* img = frames + (size_t)b * p.H * p.W * 3;
    for (int y = y0; y < y1; ++y) {
      const uint8_t* row = img + ((size_t)y * p.W + x0) * 3;
      for (int x = 0; x < (x1 - x0); ++x) {
        s0 += __ldg(row + 3 * x);
        s1 += __ldg(row + 3 * x + 1);
        s2 += __ldg(row + 3 * x + 2);
      }
    }
    const float kh = (float)(y1 - y0), kw = (float)(x1 - x0);
    const size_t plane = (size_t)lh * lw;
    float* o = levels + p.level_off[l] + (size_t)b * 3 * plane + (size_t)oy * lw + ox;
    o[0] = mul_rn(sub_rn(div_rn(div_rn((float)s0, kh), kw), 127.5f), 0.0078125f);
    o[plane] = mul_rn(sub_rn(div_rn(div_rn((float)s1, kh), kw), 127.5f), 0.0078125f);
    o[2 * plane] = mul_rn(sub_rn(div_rn(div_rn((float)s2, kh), kw), 127.5f), 0.0078125f);
  }
}

}  // namespace

extern "C" int vnfr_pyramid_resize_norm(const VnfrPyramid* pyr, const uint8_t* frames, float* levels, void* stream) {
  VNFR_REQUIRE(pyr != nullptr, "pyramid plan is null");
  if (pyr->B == 0 || pyr->n_levels == 0) return VNFR_OK;
  PyrResizeParams p;
  p.B = pyr->B; p.H = pyr->H; p.W = pyr->W; p.n_levels = pyr->n_levels;
  for (int l = 0; l < pyr->n_levels; ++l) {
    p.lh[l] = pyr->lh[l]; p.lw[l] = pyr->lw[l]; p.level_off[l] = pyr->level_off[l]; p.px_off[l] = pyr->px_off[l];
  }
  p.px_off[pyr->n_levels] = pyr->px_off[pyr->n_levels];
  const long long total = p.px_off[p.n_levels] * p.B;
  long long grid = (total + 255) / 256;
  if (grid > 148LL * 64) grid = 148LL * 64;
  pyramid_resize_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(p, frames, levels);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
