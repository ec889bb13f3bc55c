// Final face extraction: one gather kernel that turns detections into 160x160 (or SxS) encoder inputs.
//
// mode 0 ("extract"): MTCNN.extract / extract_face / crop_resize for torch.Tensor frames + fixed_image_standardization
//   (mtcnn.py:458-518; detect_face.py:317-322, :342-378): margin-expand, int-truncate + clamp the box, adaptive-average
//   resize the crop to SxS, truncate to u8 (`.byte()`), then (x-127.5)/128.
// mode 1 ("align"): the demo_video path (demo_image.py:174-199 get_face_from_boxes, :236-239 move_landmark_to_box;
//   align_face.py:51-57 alignment): integer crop, 5-point least-squares similarity (Umeyama with scale; closed form in
//   2-D) onto the template, cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) in OpenCV's fixed-point arithmetic, then
//   transforms_default (data_loader/__init__.py:27-34, 52-56).
// mode 2 / mode 3: MTCNN.extract for numpy.ndarray / PIL.Image frames (detect_face.py:309-325 crop_resize): the same box
//   arithmetic as mode 0, but the crop is resampled the way the reference's library does it for that input type --
//   mode 2 = cv2.resize(INTER_AREA): fractional-coverage box filter in float when shrinking in both directions (with
//   OpenCV's integer-ratio fast path), 11-bit fixed-point bilinear with the "area" source coordinates otherwise;
//   mode 3 = PIL.Image.resize(BILINEAR): triangle filter whose support scales with the reduction, horizontal pass rounded
//   to u8 before the vertical pass (Pillow's 8-bit resampler, 22-bit fixed-point coefficients).
// All modes write the u8 face (HWC, what the reference's glue returns) and the standardised 16-bit NHWC8 tensor the encoder
// consumes, reading the u8 frame directly -- faces never visit the host.
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>

extern long long g_vnfr_launches;

namespace {

// OpenCV initInterTab2D(INTER_LINEAR, fixpt): weights summing to 32768.  The (0,0) entry is 32768 itself, which does
// not fit a signed short -- kept unsigned here (OpenCV's u8 path behaves as if it were +32768; pinned against cv2).
// It lives in global memory and is staged into shared memory per CTA: every lane indexes it with its own (fy, fx), and
// divergent __constant__ reads serialise (the first version spent 4.9 ms per 768 faces on exactly that).
__device__ unsigned short g_bilin[32 * 32 * 4];
constexpr int MAX_S = 256;                 // largest supported output edge

struct FaceArgs {
  const uint8_t* frames;
  int B, H, W, capf, S, mode, margin, max_faces, f16, s2d;
  int word_loads;          // mode 1: interior pixels read their taps with aligned 32-bit loads (host: geometry allows it)
  const int* count;        // [B]
  const float* box;        // [B][capf][5]
  const float* pts;        // [B][capf][10]  (x0,y0,...,x4,y4)
  const int* offs;         // [B+1]
  float tmpl[10];          // template landmarks (mode 1)
  uint8_t* face_u8;        // [max_faces][S][S][3]
  unsigned short* face_h;  // [max_faces][S][S][8]
  int* face_img;           // [max_faces] image index of each face (nullable)
  int* status;
};

__device__ __forceinline__ void store_px(const FaceArgs& a, size_t px, unsigned r, unsigned g, unsigned b) {
  if (a.face_u8 != nullptr) {
    uint8_t* u = a.face_u8 + px * 3;
    u[0] = (uint8_t)r; u[1] = (uint8_t)g; u[2] = (uint8_t)b;
  }
  // (x - 127.5) / 128 : exact in fp32 (power-of-two divisor)
  const float fr = ((float)r - 127.5f) * 0.0078125f, fg = ((float)g - 127.5f) * 0.0078125f, fb = ((float)b - 127.5f) * 0.0078125f;
  uint32_t w0, w1;
  if (a.f16) {
    const __half2 h0 = __floats2half2_rn(fr, fg), h1 = __floats2half2_rn(fb, 0.f);
    w0 = *reinterpret_cast<const uint32_t*>(&h0); w1 = *reinterpret_cast<const uint32_t*>(&h1);
  } else {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(fr, fg), h1 = __floats2bfloat162_rn(fb, 0.f);
    w0 = *reinterpret_cast<const uint32_t*>(&h0); w1 = *reinterpret_cast<const uint32_t*>(&h1);
  }
  if (a.s2d) {
    // space-to-depth layout [f][ceil(S/2)][ceil(S/2)][16]: channel = ((y&1)*2 + (x&1))*4 + c -- the stride-2 3x3 stem
    // convolution becomes a stride-1 2x2 convolution over 16 channels (see encoder_plan.pack_stem_s2d)
    const int S = a.S, S2 = (S + 1) >> 1;
    const size_t f = px / ((size_t)S * S);
    const int rem = (int)(px - f * (size_t)S * S);
    const int y = rem / S, x = rem - y * S;
    unsigned short* dst = a.face_h + ((f * S2 + (y >> 1)) * S2 + (x >> 1)) * 16 + (((y & 1) << 1) | (x & 1)) * 4;
    *reinterpret_cast<uint2*>(dst) = make_uint2(w0, w1);
  } else {
    *reinterpret_cast<uint4*>(a.face_h + px * 8) = make_uint4(w0, w1, 0u, 0u);
  }
}

// ---- mode 2: cv2.resize(crop, (S,S), INTER_AREA) for one output pixel ------------------------------------------------
// OpenCV resize.cpp: area path (computeResizeAreaTab / ResizeArea_Invoker, float accumulation in table order) when
// scale_x >= 1 and scale_y >= 1, with resizeAreaFast_ for integer ratios; otherwise the bilinear path with
// area_mode source coordinates (HResizeLinear / VResizeLinear<uchar>, INTER_RESIZE_COEF_BITS = 11).
struct AreaTaps { int first, n; float w_first, w_mid, w_last; };     // source cells [first, first+n): weights first | mid.. | last

__device__ __forceinline__ AreaTaps area_taps(int d, double scale, int ssize) {
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  AreaTaps t;
  const bool left = sx1 - fsx1 > 1e-3, right = fsx2 - sx2 > 1e-3;
  t.first = left ? sx1 - 1 : sx1;
  t.n = (left ? 1 : 0) + (sx2 - sx1) + (right ? 1 : 0);
  t.w_mid = (float)(1.0 / cell);
  t.w_first = left ? (float)((sx1 - fsx1) / cell) : t.w_mid;
  t.w_last = right ? (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell) : t.w_mid;
  if (t.n == 1 && left) t.w_last = t.w_first;          // a single (left-partial) cell
  return t;
}
__device__ __forceinline__ float area_w(const AreaTaps& t, int k) { return k == 0 ? t.w_first : (k == t.n - 1 ? t.w_last : t.w_mid); }

__device__ __forceinline__ void linear_area_tap(int d, double scale, double inv_scale, int ssize, int& s, int& c0, int& c1) {
  s = (int)floor(d * scale);
  float f = (float)((d + 1) - (s + 1) * inv_scale);
  f = f <= 0.f ? 0.f : f - floorf(f);
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  c0 = (int)rintf((1.f - f) * 2048.f);                  // saturate_cast<short>(coef * INTER_RESIZE_COEF_SCALE)
  c1 = (int)rintf(f * 2048.f);
}

__device__ void cv_area_pixel(const uint8_t* crop, int pitch_px, int cw, int ch, int S, int ox, int oy, unsigned out[3]) {
  // exactly cv::resize's doubles: inv_scale = dsize / ssize, scale = 1 / inv_scale (NOT ssize / dsize: 1/(160/98) < 98/160,
  // which moves cvFloor(dy * scale) to the previous source row where dy * 98/160 is an integer)
  const double inv_scale_x = (double)S / cw, inv_scale_y = (double)S / ch;
  const double scale_x = 1.0 / inv_scale_x, scale_y = 1.0 / inv_scale_y;
  if (scale_x >= 1.0 && scale_y >= 1.0) {
    if (cw % S == 0 && ch % S == 0) {                   // resizeAreaFast_: integer box sums
      const int ix = cw / S, iy = ch / S;
      int s[3] = {0, 0, 0};
      for (int y = 0; y < iy; ++y) {
        const uint8_t* row = crop + ((size_t)(oy * iy + y) * pitch_px + ox * ix) * 3;
        for (int x = 0; x < ix; ++x) { s[0] += __ldg(row + 3 * x); s[1] += __ldg(row + 3 * x + 1); s[2] += __ldg(row + 3 * x + 2); }
      }
      const float sc = 1.f / (float)(ix * iy);
      for (int c = 0; c < 3; ++c) {
        const int v = (ix == 2 && iy == 2) ? (s[c] + 2) >> 2 : (int)rintf((float)s[c] * sc);
        out[c] = (unsigned)min(max(v, 0), 255);
      }
      return;
    }
    const AreaTaps tx = area_taps(ox, scale_x, cw), ty = area_taps(oy, scale_y, ch);
    float sum[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < ty.n; ++j) {
      const uint8_t* row = crop + ((size_t)(ty.first + j) * pitch_px + tx.first) * 3;
      float buf[3] = {0.f, 0.f, 0.f};
      for (int k = 0; k < tx.n; ++k) {
        const float al = area_w(tx, k);
        buf[0] = __fmaf_rn((float)__ldg(row + 3 * k), al, buf[0]);
        buf[1] = __fmaf_rn((float)__ldg(row + 3 * k + 1), al, buf[1]);
        buf[2] = __fmaf_rn((float)__ldg(row + 3 * k + 2), al, buf[2]);
      }
      const float be = area_w(ty, j);
      for (int c = 0; c < 3; ++c) sum[c] = __fmaf_rn(be, buf[c], sum[c]);
    }
    for (int c = 0; c < 3; ++c) out[c] = (unsigned)min(max((int)rintf(sum[c]), 0), 255);
    return;
  }
  int sx, sy, a0, a1, b0, b1;
  linear_area_tap(ox, scale_x, inv_scale_x, cw, sx, a0, a1);
  linear_area_tap(oy, scale_y, inv_scale_y, ch, sy, b0, b1);
  const int sx1 = min(sx + 1, cw - 1), sy1 = min(sy + 1, ch - 1);
  const uint8_t* r0 = crop + (size_t)sy * pitch_px * 3;
  const uint8_t* r1 = crop + (size_t)sy1 * pitch_px * 3;
  for (int c = 0; c < 3; ++c) {
    const int h0 = (int)__ldg(r0 + 3 * sx + c) * a0 + (int)__ldg(r0 + 3 * sx1 + c) * a1;      // HResizeLinear: x 2^11
    const int h1 = (int)__ldg(r1 + 3 * sx + c) * a0 + (int)__ldg(r1 + 3 * sx1 + c) * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;             // VResizeLinear<uchar>
    out[c] = (unsigned)min(max(v, 0), 255);
  }
}

// ---- mode 3: PIL crop.resize((S,S), Image.BILINEAR) for one output pixel -------------------------------------------
// Pillow Resample.c (8 bits per channel): precompute_coeffs with the triangle filter (support 1 x max(scale, 1)),
// coefficients normalised then rounded to PRECISION_BITS = 22 fixed point; horizontal pass first, rounded and clipped to
// u8, then the vertical pass.
struct PilTaps { int xmin, n; double ww, center, ss; };

__device__ __forceinline__ PilTaps pil_taps(int d, int in_size, int out_size) {
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  PilTaps t;
  t.center = (d + 0.5) * scale;
  t.ss = 1.0 / filterscale;
  int xmin = (int)(t.center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(t.center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  t.xmin = xmin; t.n = xmax - xmin;
  double ww = 0.0;
  for (int x = 0; x < t.n; ++x) {
    double v = (x + xmin - t.center + 0.5) * t.ss;
    v = v < 0 ? -v : v;
    ww += v < 1.0 ? 1.0 - v : 0.0;
  }
  t.ww = ww;
  return t;
}
__device__ __forceinline__ int pil_coef(const PilTaps& t, int x) {
  double v = (x + t.xmin - t.center + 0.5) * t.ss;
  v = v < 0 ? -v : v;
  double w = v < 1.0 ? 1.0 - v : 0.0;
  if (t.ww != 0.0) w /= t.ww;
  return (int)(w < 0 ? -0.5 + w * 4194304.0 : 0.5 + w * 4194304.0);
}
__device__ __forceinline__ unsigned pil_clip8(int v) { v >>= 22; return (unsigned)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__device__ void pil_bilinear_pixel(const uint8_t* crop, int pitch_px, int cw, int ch, int S, int ox, int oy, unsigned out[3]) {
  const PilTaps tx = pil_taps(ox, cw, S), ty = pil_taps(oy, ch, S);
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int j = 0; j < ty.n; ++j) {
    const uint8_t* row = crop + ((size_t)(ty.xmin + j) * pitch_px + tx.xmin) * 3;
    int h[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int k = 0; k < tx.n; ++k) {
      const int c = pil_coef(tx, k);
      h[0] += (int)__ldg(row + 3 * k) * c; h[1] += (int)__ldg(row + 3 * k + 1) * c; h[2] += (int)__ldg(row + 3 * k + 2) * c;
    }
    const int cy = pil_coef(ty, j);
    for (int c = 0; c < 3; ++c) acc[c] += (int)pil_clip8(h[c]) * cy;
  }
  for (int c = 0; c < 3; ++c) out[c] = pil_clip8(acc[c]);
}

// MODE is a template parameter so that the hot modes (0, 1) do not carry the registers of the cv2 / PIL resamplers.
template <int MODE>
__global__ void __launch_bounds__(256) face_crop_kernel(const FaceArgs a) {
  const int total_raw = a.offs[a.B];
  if (total_raw > a.max_faces && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.status, 16);
  const int total = min(total_raw, a.max_faces);
  const int S = a.S;
  __shared__ double s_m[6];     // inverse affine (mode 1)
  __shared__ int s_box[4];      // integer crop box x1,y1,x2,y2 (exclusive ends)
  __shared__ unsigned short s_tab[32 * 32 * 4];
  __shared__ int s_ad[MAX_S], s_bd[MAX_S], s_x0[MAX_S], s_y0[MAX_S];   // OpenCV's adelta / bdelta / per-row X0, Y0
  if (MODE == 1)
    for (int i = threadIdx.x; i < 32 * 32 * 4 / 2; i += blockDim.x)
      reinterpret_cast<uint32_t*>(s_tab)[i] = reinterpret_cast<const uint32_t*>(g_bilin)[i];
  for (int f = blockIdx.x; f < total; f += gridDim.x) {
    // locate (image, slot)
    int lo = 0, hi = a.B;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (a.offs[mid] <= f) lo = mid; else hi = mid; }
    const int b = lo, slot = f - a.offs[lo];
    const float* bx = a.box + ((size_t)b * a.capf + slot) * 5;
    if (threadIdx.x == 0) {
      if (a.face_img != nullptr) a.face_img[f] = b;
      if (MODE != 1) {
        // extract_face: margin in fp32 like numpy float32 scalars (detect_face.py:358-368)
        const float m0 = div_rn(mul_rn((float)a.margin, sub_rn(bx[2], bx[0])), (float)(a.S - a.margin));
        const float m1 = div_rn(mul_rn((float)a.margin, sub_rn(bx[3], bx[1])), (float)(a.S - a.margin));
        s_box[0] = (int)fmaxf(sub_rn(bx[0], mul_rn(m0, 0.5f)), 0.f);
        s_box[1] = (int)fmaxf(sub_rn(bx[1], mul_rn(m1, 0.5f)), 0.f);
        s_box[2] = (int)fminf(add_rn(bx[2], mul_rn(m0, 0.5f)), (float)a.W);
        s_box[3] = (int)fminf(add_rn(bx[3], mul_rn(m1, 0.5f)), (float)a.H);
      } else {
        // get_face_from_boxes (demo_image.py:179-182)
        s_box[0] = max((int)bx[0], 0);
        s_box[1] = max((int)bx[1], 0);
        s_box[2] = min((int)add_rn(bx[2], 1.0f), a.W);
        s_box[3] = min((int)add_rn(bx[3], 1.0f), a.H);
        // landmarks relative to the UNclamped float corner (demo_image.py:236-239), then Umeyama(moved -> template)
        const float* p = a.pts + ((size_t)b * a.capf + slot) * 10;
        double sx[5], sy[5], mx = 0, my = 0, tx = 0, ty = 0;
        for (int j = 0; j < 5; ++j) {
          sx[j] = (double)sub_rn(p[2 * j], bx[0]); sy[j] = (double)sub_rn(p[2 * j + 1], bx[1]);
          mx += sx[j]; my += sy[j]; tx += (double)a.tmpl[2 * j]; ty += (double)a.tmpl[2 * j + 1];
        }
        mx /= 5; my /= 5; tx /= 5; ty /= 5;
        double sxx = 0, num_a = 0, num_b = 0;
        for (int j = 0; j < 5; ++j) {
          const double xs = sx[j] - mx, ys = sy[j] - my, xd = (double)a.tmpl[2 * j] - tx, yd = (double)a.tmpl[2 * j + 1] - ty;
          sxx += xs * xs + ys * ys;
          num_a += xs * xd + ys * yd;
          num_b += xs * yd - ys * xd;
        }
        const double ca = num_a / sxx, cb = num_b / sxx;       // M = [[ca, -cb, t0], [cb, ca, t1]]
        const double t0 = tx - (ca * mx - cb * my), t1 = ty - (cb * mx + ca * my);
        // cv2.invertAffineTransform
        double D = ca * ca + cb * cb;
        D = D != 0 ? 1.0 / D : 0.0;
        const double A11 = ca * D, A22 = ca * D, A12 = cb * D, A21 = -cb * D;
        s_m[0] = A11; s_m[1] = A12; s_m[2] = -A11 * t0 - A12 * t1;
        s_m[3] = A21; s_m[4] = A22; s_m[5] = -A21 * t0 - A22 * t1;
      }
    }
    __syncthreads();
    if (MODE == 1) {
      // per-column / per-row fixed-point terms (AB_BITS = 10), exactly OpenCV's adelta/bdelta and X0/Y0 incl. round_delta
      for (int i = threadIdx.x; i < S; i += blockDim.x) {
        s_ad[i] = (int)llrint(s_m[0] * (double)i * 1024.0);
        s_bd[i] = (int)llrint(s_m[3] * (double)i * 1024.0);
        s_x0[i] = (int)llrint((s_m[1] * (double)i + s_m[2]) * 1024.0) + 16;
        s_y0[i] = (int)llrint((s_m[4] * (double)i + s_m[5]) * 1024.0) + 16;
      }
      __syncthreads();
    }
    const int x1 = s_box[0], y1 = s_box[1], cw = s_box[2] - s_box[0], ch = s_box[3] - s_box[1];
    const uint8_t* img = a.frames + (size_t)b * a.H * a.W * 3;
    if (MODE == 0) {
      for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int oy = i / S, ox = i - oy * S;
        unsigned r = 0, g = 0, bl = 0;
        if (cw > 0 && ch > 0) {
          const int ys = (oy * ch) / S, ye = ((oy + 1) * ch + S - 1) / S;
          const int xs = (ox * cw) / S, xe = ((ox + 1) * cw + S - 1) / S;
          unsigned s0 = 0, s1 = 0, s2 = 0;
          for (int y = ys; y < ye; ++y) {
            const uint8_t* row = img + ((size_t)(y1 + y) * a.W + (x1 + xs)) * 3;
            for (int x = 0; x < xe - xs; ++x) { s0 += __ldg(row + 3 * x); s1 += __ldg(row + 3 * x + 1); s2 += __ldg(row + 3 * x + 2); }
          }
          const float kh = (float)(ye - ys), kw = (float)(xe - xs);
          r = (unsigned)div_rn(div_rn((float)s0, kh), kw);     // .byte(): truncation
          g = (unsigned)div_rn(div_rn((float)s1, kh), kw);
          bl = (unsigned)div_rn(div_rn((float)s2, kh), kw);
        }
        store_px(a, (size_t)f * S * S + i, r, g, bl);
      }
    } else if (MODE == 2 || MODE == 3) {
      const uint8_t* crop = img + ((size_t)y1 * a.W + x1) * 3;
      for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int oy = i / S, ox = i - oy * S;
        unsigned c3[3] = {0, 0, 0};
        if (cw > 0 && ch > 0) {
          if (MODE == 2) cv_area_pixel(crop, a.W, cw, ch, S, ox, oy, c3);
          else pil_bilinear_pixel(crop, a.W, cw, ch, S, ox, oy, c3);
        }
        store_px(a, (size_t)f * S * S + i, c3[0], c3[1], c3[2]);
      }
    } else {
      // cv2.warpAffine, INTER_LINEAR fixed point: AB_BITS = 10, INTER_BITS = 5, weights 2^15
      const bool words_ok = a.word_loads != 0;
      for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int oy = i / S, ox = i - oy * S;
        const int X = (s_x0[oy] + s_ad[ox]) >> 5, Y = (s_y0[oy] + s_bd[ox]) >> 5;
        int sx = X >> 5, sy = Y >> 5;
        sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);
        sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
        const int fx = X & 31, fy = Y & 31;
        const unsigned short* wt = s_tab + (fy * 32 + fx) * 4;
        int acc[3] = {0, 0, 0};
        if (words_ok && sx >= 0 && sx + 1 < cw && sy >= 0 && sy + 1 < ch) {
          // interior pixel: the two taps of a row are 6 consecutive bytes -> two (three when the span starts on byte 3) aligned
          // 32-bit loads per row instead of six byte loads (the same bytes, the same integer arithmetic)
          const uint8_t* p0 = img + ((size_t)(y1 + sy) * a.W + (x1 + sx)) * 3;
#pragma unroll
          for (int row = 0; row < 2; ++row) {
            const uintptr_t pa = reinterpret_cast<uintptr_t>(p0 + (size_t)row * a.W * 3);
            const uint32_t* al = reinterpret_cast<const uint32_t*>(pa & ~(uintptr_t)3);
            const unsigned sh = (unsigned)(pa & 3) * 8u;
            const uint32_t w0 = __ldg(al), w1 = __ldg(al + 1);
            const uint32_t lo = __funnelshift_r(w0, w1, sh);
            const uint32_t hi = sh == 24u ? __funnelshift_r(w1, __ldg(al + 2), 24u) : (w1 >> sh);
            const int wl = (int)wt[2 * row], wr = (int)wt[2 * row + 1];
            acc[0] += wl * (int)(lo & 0xffu) + wr * (int)(lo >> 24);
            acc[1] += wl * (int)((lo >> 8) & 0xffu) + wr * (int)(hi & 0xffu);
            acc[2] += wl * (int)((lo >> 16) & 0xffu) + wr * (int)((hi >> 8) & 0xffu);
          }
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int yy = sy + (t >> 1), xx = sx + (t & 1);
            if (yy >= 0 && yy < ch && xx >= 0 && xx < cw) {
              const uint8_t* px = img + ((size_t)(y1 + yy) * a.W + (x1 + xx)) * 3;
              const int w = (int)wt[t];
              acc[0] += w * (int)__ldg(px); acc[1] += w * (int)__ldg(px + 1); acc[2] += w * (int)__ldg(px + 2);
            }
          }
        }
        unsigned c3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int v = (acc[c] + (1 << 14)) >> 15;
          c3[c] = (unsigned)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
        store_px(a, (size_t)f * S * S + i, c3[0], c3[1], c3[2]);
      }
    }
    __syncthreads();
  }
}

__global__ void scan_counts_kernel2(const int* __restrict__ count, int B, int cap, int* __restrict__ offs) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int s = 0;
    for (int b = 0; b < B; ++b) { offs[b] = s; s += min(count[b], cap); }
    offs[B] = s;
  }
}

// OpenCV imgwarp.cpp initInterTab2D(INTER_LINEAR, fixpt = true) in float32 arithmetic
void build_bilinear_table(unsigned short* tab) {
  float t1[32][2];
  for (int i = 0; i < 32; ++i) {
    const float x = (float)i * (1.0f / 32);
    t1[i][0] = 1.0f - x;
    t1[i][1] = x;
  }
  for (int i = 0; i < 32; ++i)
    for (int j = 0; j < 32; ++j) {
      int iv[4];
      int isum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          const float v = t1[i][k1] * t1[j][k2];
          long r = lrintf(v * 32768.0f);
          r = r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
          iv[k1 * 2 + k2] = (int)r;
          isum += (int)r;
        }
      if (isum != 32768) {
        const int diff = isum - 32768;
        int mk = 0, Mk = 0;
        for (int k = 0; k < 4; ++k) {
          if (iv[k] < iv[mk]) mk = k;
          else if (iv[k] > iv[Mk]) Mk = k;
        }
        if (diff < 0) iv[Mk] -= diff; else iv[mk] -= diff;
      }
      for (int k = 0; k < 4; ++k) tab[(i * 32 + j) * 4 + k] = (unsigned short)iv[k];
    }
}

}  // namespace

extern "C" int vnfr_face_crops(const uint8_t* frames, int B, int H, int W, int capf, const int32_t* count, const float* box,
                               const float* pts, int mode, int image_size, int margin, const float* template_host, int dtype,
                               int max_faces, int32_t* offs, uint8_t* face_u8, void* face_half, int32_t* face_img, int32_t* status,
                               int half_layout, void* stream) {
  VNFR_REQUIRE(frames && count && box && offs && face_half && status, "null pointer");
  VNFR_REQUIRE(image_size <= MAX_S, "image_size larger than 256 is not supported");
  VNFR_REQUIRE(mode == 0 || mode == 2 || mode == 3 || (mode == 1 && pts != nullptr && template_host != nullptr),
               "mode must be 0 / 2 / 3 (extract: tensor / ndarray / PIL resampler) or 1 (align: needs landmarks and a template)");
  VNFR_REQUIRE(image_size > 0 && margin >= 0 && margin < image_size, "bad image_size / margin");
  if (B == 0 || max_faces == 0) return VNFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  static VnfrPerDevice tab_set_once = {};
  if (vnfr_first_on_device(tab_set_once)) {
    static unsigned short tab[32 * 32 * 4];
    build_bilinear_table(tab);
    VNFR_CUDA(cudaMemcpyToSymbolAsync(g_bilin, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st));
  }
  scan_counts_kernel2<<<1, 32, 0, st>>>(count, B, capf, offs);
  ++g_vnfr_launches;
  FaceArgs a;
  a.frames = frames; a.B = B; a.H = H; a.W = W; a.capf = capf; a.S = image_size; a.mode = mode; a.margin = margin;
  a.max_faces = max_faces; a.f16 = dtype == 1; a.s2d = half_layout == 1;
  // VNFR_FACE_CROP_BYTES=1 keeps the byte loads everywhere (A/B switch; tests compare the two paths bit for bit)
  a.word_loads = (((size_t)W * 3) % 4 == 0 && ((uintptr_t)frames & 3) == 0 && getenv("VNFR_FACE_CROP_BYTES") == nullptr) ? 1 : 0;
  a.count = count; a.box = box; a.pts = pts; a.offs = offs;
  for (int i = 0; i < 10; ++i) a.tmpl[i] = template_host ? template_host[i] : 0.f;
  a.face_u8 = face_u8; a.face_h = (unsigned short*)face_half; a.face_img = face_img; a.status = status;
  int grid = max_faces < 148 * 8 ? max_faces : 148 * 8;
  switch (mode) {
    case 0: face_crop_kernel<0><<<grid, 256, 0, st>>>(a); break;
    case 1: face_crop_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 2: face_crop_kernel<2><<<grid, 256, 0, st>>>(a); break;
    default: face_crop_kernel<3><<<grid, 256, 0, st>>>(a); break;
  }
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
