// Final face extraction: one gather kernel that turns detections into 160x160 (or SxS) encoder inputs.
//
// mode 0 ("extract"): MTCNN.extract / extract_face / crop_resize for torch.Tensor frames + fixed_image_standardization
//   (mtcnn.py:458-518; detect_face.py:317-322, :342-378): margin-expand, int-truncate + clamp the box, adaptive-average
//   resize the crop to SxS, truncate to u8 (`.byte()`), then (x-127.5)/128.
// mode 1 ("align"): the demo_video path (demo_image.py:174-199 get_face_from_boxes, :236-239 move_landmark_to_box;
//   align_face.py:51-57 alignment): integer crop, 5-point least-squares similarity (Umeyama with scale; closed form in
//   2-D) onto the template, cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) in OpenCV's fixed-point arithmetic, then
//   transforms_default (data_loader/__init__.py:27-34, 52-56).
// Both write the u8 face (HWC, what the reference's glue returns) and the standardised 16-bit NHWC8 tensor the encoder
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

__global__ void __launch_bounds__(256) face_crop_kernel(const FaceArgs a) {
  const int total_raw = a.offs[a.B];
  if (total_raw > a.max_faces && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.status, 16);
  const int total = min(total_raw, a.max_faces);
  const int S = a.S;
  __shared__ double s_m[6];     // inverse affine (mode 1)
  __shared__ int s_box[4];      // integer crop box x1,y1,x2,y2 (exclusive ends)
  __shared__ unsigned short s_tab[32 * 32 * 4];
  __shared__ int s_ad[MAX_S], s_bd[MAX_S], s_x0[MAX_S], s_y0[MAX_S];   // OpenCV's adelta / bdelta / per-row X0, Y0
  if (a.mode == 1)
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
      if (a.mode == 0) {
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
    if (a.mode == 1) {
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
    if (a.mode == 0) {
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
    } else {
      // cv2.warpAffine, INTER_LINEAR fixed point: AB_BITS = 10, INTER_BITS = 5, weights 2^15
      for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        const int oy = i / S, ox = i - oy * S;
        const int X = (s_x0[oy] + s_ad[ox]) >> 5, Y = (s_y0[oy] + s_bd[ox]) >> 5;
        int sx = X >> 5, sy = Y >> 5;
        sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);
        sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
        const int fx = X & 31, fy = Y & 31;
        const unsigned short* wt = s_tab + (fy * 32 + fx) * 4;
        int acc[3] = {0, 0, 0};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int yy = sy + (t >> 1), xx = sx + (t & 1);
          if (yy >= 0 && yy < ch && xx >= 0 && xx < cw) {
            const uint8_t* px = img + ((size_t)(y1 + yy) * a.W + (x1 + xx)) * 3;
            const int w = (int)wt[t];
            acc[0] += w * (int)__ldg(px); acc[1] += w * (int)__ldg(px + 1); acc[2] += w * (int)__ldg(px + 2);
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
  VNFR_REQUIRE(mode == 0 || (mode == 1 && pts != nullptr && template_host != nullptr), "align mode needs landmarks and a template");
  VNFR_REQUIRE(image_size > 0 && margin >= 0 && margin < image_size, "bad image_size / margin");
  if (B == 0 || max_faces == 0) return VNFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  static bool tab_set = false;
  if (!tab_set) {
    static unsigned short tab[32 * 32 * 4];
    build_bilinear_table(tab);
    VNFR_CUDA(cudaMemcpyToSymbolAsync(g_bilin, tab, sizeof(tab), 0, cudaMemcpyHostToDevice, st));
    tab_set = true;
  }
  scan_counts_kernel2<<<1, 32, 0, st>>>(count, B, capf, offs);
  ++g_vnfr_launches;
  FaceArgs a;
  a.frames = frames; a.B = B; a.H = H; a.W = W; a.capf = capf; a.S = image_size; a.mode = mode; a.margin = margin;
  a.max_faces = max_faces; a.f16 = dtype == 1; a.s2d = half_layout == 1;
  a.count = count; a.box = box; a.pts = pts; a.offs = offs;
  for (int i = 0; i < 10; ++i) a.tmpl[i] = template_host ? template_host[i] : 0.f;
  a.face_u8 = face_u8; a.face_h = (unsigned short*)face_half; a.face_img = face_img; a.status = status;
  int grid = max_faces < 148 * 8 ? max_faces : 148 * 8;
  face_crop_kernel<<<grid, 256, 0, st>>>(a);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
