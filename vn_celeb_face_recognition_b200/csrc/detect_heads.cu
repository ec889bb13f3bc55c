// R-Net and O-Net: batched crop -> area-resize -> normalise -> network heads, one persistent kernel per net.
//
// Replaces the per-candidate Python loops detect_face.py:108-114 / :136-143 (slice + imresample(24|48) + cat +
// normalise; ~256 launch chains per 1080p frame in the reference), fixed_batch_process (:16-23) and RNet.forward /
// ONet.forward (mtcnn.py:84-99, :138-157).  Each CTA takes G candidates at a time: the crops are area-resized straight
// from the u8 frame into shared memory (exact integer window sums, sum/kh/kw in fp32 like torch), every layer's
// activations stay in shared memory, weights ([K][Cout] fp32, L2-resident) are read with warp-uniform 16-byte loads.
// Arithmetic is fp32 FMA: the `score > threshold` decisions of stages 2/3 must match the fp32 reference.
#include "common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>

extern long long g_vnfr_launches;

namespace {

constexpr int NT = 512;   // threads per CTA

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

// Packed fp32 pair arithmetic of sm_100 (FFMA2): two IEEE fp32 FMAs per lane in ONE issue slot.  The convolutions here
// are issue-bound (FMA + shared-memory loads + address arithmetic compete for the scheduler), so halving the FMA
// instruction count raises the FMA pipe utilisation; results are bit-identical to scalar fmaf.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2f(float lo, float hi) {
  return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo2f(f32x2 v) { return __uint_as_float((unsigned)(v & 0xFFFFFFFFull)); }
__device__ __forceinline__ float hi2f(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ void fma2(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// ---------------------------------------------------------------------------------------------------------------
// conv (valid, stride 1) + bias + PReLU on shared-memory activations for output channels [c_begin, c_end).
//   in  [G][CIN][IH][IW]   out [G][c_end - c_begin][OH][OW] (channel c stored at index c - c_begin)
//   w   [CIN*KH*KW][COUT] global, k = (ci*KH + ky)*KW + kx
// Thread item = (CH output channels) x (POS positions p, p+NQ, ..., p+(POS-1)NQ) x (one of KS slices of the input
// channels): CH*POS FMAs per POS shared-memory loads + CH/4 warp-uniform 16-byte weight loads.  The first version
// (4 x 4 tiles) was bound by the load/store unit (ncu: LSU wavefronts 82 % of peak, FMA pipe 28 %); 8-channel tiles
// halve the loads per FMA.  KS > 1 splits the reduction over thread groups for the small late layers (partial sums
// through `scratch` [KS][out size]), which otherwise leave most of the CTA idle.
template <int CIN, int COUT, int KH, int KW, int IH, int IW, int G, int CH, int POS, int KS>
__device__ __forceinline__ void conv_prelu_smem(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ scratch,
                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                const float* __restrict__ alpha, int c_begin, int c_end) {
  constexpr int OH = IH - KH + 1, OW = IW - KW + 1, NPOS = G * OH * OW, NQ = (NPOS + POS - 1) / POS;
  constexpr int CI_PER = CIN / KS;
  static_assert(CIN % KS == 0 && CH % 4 == 0, "bad tiling");
  const int ngroups = (c_end - c_begin) / CH;
  const int cn = c_end - c_begin;
  const int per_slice = ngroups * NQ;
  for (int item = threadIdx.x; item < KS * per_slice; item += NT) {
    const int ks = item / per_slice, it2 = item - ks * per_slice;
    const int cg = it2 / NQ, q = it2 - cg * NQ;
    const int c0 = c_begin + CH * cg;
    int off[POS];
    bool ok[POS];
#pragma unroll
    for (int j = 0; j < POS; ++j) {
      const int p = q + j * NQ;
      ok[j] = p < NPOS;
      const int pp = ok[j] ? p : 0;
      const int g = pp / (OH * OW), r = pp - g * (OH * OW);
      const int oy = r / OW, ox = r - oy * OW;
      off[j] = (g * CIN * IH + oy) * IW + ox;
    }
    f32x2 acc[POS][CH / 2];                              // channel pairs (c, c+1), packed FFMA2
#pragma unroll
    for (int c4 = 0; c4 < CH / 4; ++c4) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ks == 0) b4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * c4));
#pragma unroll
      for (int j = 0; j < POS; ++j) { acc[j][2 * c4] = pack2f(b4.x, b4.y); acc[j][2 * c4 + 1] = pack2f(b4.z, b4.w); }
    }
    const float* wp = w + c0;
    const int ci0 = ks * CI_PER;
#pragma unroll 1
    for (int ci = ci0; ci < ci0 + CI_PER; ++ci) {
#pragma unroll
      for (int ky = 0; ky < KH; ++ky)
#pragma unroll
        for (int kx = 0; kx < KW; ++kx) {
          f32x2 wv[CH / 2];
#pragma unroll
          for (int c4 = 0; c4 < CH / 4; ++c4) {
            const ulonglong2 w4 = __ldg(reinterpret_cast<const ulonglong2*>(wp + ((ci * KH + ky) * KW + kx) * COUT) + c4);
            wv[2 * c4] = w4.x; wv[2 * c4 + 1] = w4.y;
          }
          const int o = (ci * IH + ky) * IW + kx;
#pragma unroll
          for (int j = 0; j < POS; ++j) {
            const float v = in[off[j] + o];
            const f32x2 vv = pack2f(v, v);
#pragma unroll
            for (int c = 0; c < CH / 2; ++c) fma2(acc[j][c], wv[c], vv);
          }
        }
    }
    float* dst = KS == 1 ? out : scratch + ks * (G * cn * OH * OW);
#pragma unroll
    for (int j = 0; j < POS; ++j) {
      if (!ok[j]) continue;
      const int p = q + j * NQ;
      const int g = p / (OH * OW), r = p - g * (OH * OW);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float v = (c & 1) ? hi2f(acc[j][c >> 1]) : lo2f(acc[j][c >> 1]);
        dst[(g * cn + CH * cg + c) * (OH * OW) + r] = KS == 1 ? prelu(v, __ldg(alpha + c0 + c)) : v;
      }
    }
  }
  if (KS > 1) {
    __syncthreads();
    const int total = G * cn * OH * OW;
    for (int i = threadIdx.x; i < total; i += NT) {
      float sum = scratch[i];
#pragma unroll
      for (int k = 1; k < KS; ++k) sum += scratch[k * total + i];
      const int c = (i / (OH * OW)) % cn;
      out[i] = prelu(sum, __ldg(alpha + c_begin + c));
    }
  }
}

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Same convolution with the weights STAGED through shared memory: when the CTA owns (almost) all of the SM's shared
// memory the L1 cache is a few KB, every warp-uniform weight load goes to L2 and the FMAs stall on it (ncu source view
// of the first version: 36 % of all warp-stall samples were long-scoreboard waits on exactly those loads).  Chunks of
// CC input channels of [K][COUT] weights (for each of the KS reduction slices) are copied with cp.async into a double
// buffer `wbuf` [2][KS][CC*KH*KW*COUT] one chunk ahead of the FMAs that read them as 16-byte broadcast LDS.
template <int CIN, int COUT, int KH, int KW, int IH, int IW, int G, int CH, int POS, int KS, int CC>
__device__ __forceinline__ void conv_prelu_smem_ws(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ scratch,
                                                   float* __restrict__ wbuf, const float* __restrict__ w,
                                                   const float* __restrict__ bias, const float* __restrict__ alpha, int c_begin,
                                                   int c_end) {
  constexpr int OH = IH - KH + 1, OW = IW - KW + 1, NPOS = G * OH * OW, NQ = (NPOS + POS - 1) / POS;
  constexpr int CI_PER = CIN / KS, NCHUNK = CI_PER / CC, CHUNK_F = CC * KH * KW * COUT, BUF_F = KS * CHUNK_F;
  static_assert(CIN % KS == 0 && CI_PER % CC == 0 && CH % 4 == 0 && CHUNK_F % 4 == 0, "bad tiling");
  const int ngroups = (c_end - c_begin) / CH;
  const int cn = c_end - c_begin;
  const int per_slice = ngroups * NQ;
  const int n_items = KS * per_slice;
  auto stage = [&](int chunk, int buf) {
    // slice ks, chunk -> input channels [ks*CI_PER + chunk*CC, +CC): one contiguous block of CHUNK_F floats
    for (int i = threadIdx.x; i < BUF_F / 4; i += NT) {
      const int ks = i / (CHUNK_F / 4), j = i - ks * (CHUNK_F / 4);
      cp_async16(wbuf + buf * BUF_F + ks * CHUNK_F + 4 * j, w + (size_t)((ks * CI_PER + chunk * CC) * KH * KW) * COUT + 4 * j);
    }
    cp_async_commit_group();
  };
  for (int base = 0; base < n_items; base += NT) {
    const int item = base + threadIdx.x;
    const bool active = item < n_items;
    const int ks = active ? item / per_slice : 0, it2 = active ? item - ks * per_slice : 0;
    const int cg = it2 / NQ, q = it2 - cg * NQ;
    const int c0 = c_begin + CH * cg;
    int off[POS];
    bool ok[POS];
#pragma unroll
    for (int j = 0; j < POS; ++j) {
      const int p = q + j * NQ;
      ok[j] = active && p < NPOS;
      const int pp = ok[j] ? p : 0;
      const int g = pp / (OH * OW), r = pp - g * (OH * OW);
      const int oy = r / OW, ox = r - oy * OW;
      off[j] = (g * CIN * IH + oy) * IW + ox;
    }
    f32x2 acc[POS][CH / 2];                              // channel pairs (c, c+1)
#pragma unroll
    for (int c4 = 0; c4 < CH / 4; ++c4) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ks == 0) b4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * c4));
#pragma unroll
      for (int j = 0; j < POS; ++j) { acc[j][2 * c4] = pack2f(b4.x, b4.y); acc[j][2 * c4 + 1] = pack2f(b4.z, b4.w); }
    }
    stage(0, 0);
    for (int chunk = 0; chunk < NCHUNK; ++chunk) {
      const int buf = chunk & 1;
      if (chunk + 1 < NCHUNK) { stage(chunk + 1, buf ^ 1); cp_async_wait_group<1>(); } else { cp_async_wait_group<0>(); }
      __syncthreads();                                   // chunk `chunk` is visible to every thread
      const float* wb = wbuf + buf * BUF_F + ks * CHUNK_F + c0;
      const int ci0 = ks * CI_PER + chunk * CC;
#pragma unroll 1
      for (int cc = 0; cc < CC; ++cc) {
#pragma unroll
        for (int ky = 0; ky < KH; ++ky)
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) {
            f32x2 wv[CH / 2];
#pragma unroll
            for (int c4 = 0; c4 < CH / 4; ++c4) {
              const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wb + ((cc * KH + ky) * KW + kx) * COUT + 4 * c4);
              wv[2 * c4] = w4.x; wv[2 * c4 + 1] = w4.y;
            }
            const int o = ((ci0 + cc) * IH + ky) * IW + kx;
#pragma unroll
            for (int j = 0; j < POS; ++j) {
              const float v = in[off[j] + o];
              const f32x2 vv = pack2f(v, v);
#pragma unroll
              for (int c = 0; c < CH / 2; ++c) fma2(acc[j][c], wv[c], vv);
            }
          }
      }
      __syncthreads();                                   // everyone is done with `buf` before it is restaged
    }
    float* dst = KS == 1 ? out : scratch + ks * (G * cn * OH * OW);
#pragma unroll
    for (int j = 0; j < POS; ++j) {
      if (!ok[j]) continue;
      const int p = q + j * NQ;
      const int g = p / (OH * OW), r = p - g * (OH * OW);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const float v = (c & 1) ? hi2f(acc[j][c >> 1]) : lo2f(acc[j][c >> 1]);
        dst[(g * cn + CH * cg + c) * (OH * OW) + r] = KS == 1 ? prelu(v, __ldg(alpha + c0 + c)) : v;
      }
    }
  }
  if (KS > 1) {
    __syncthreads();
    const int total = G * cn * OH * OW;
    for (int i = threadIdx.x; i < total; i += NT) {
      float sum = scratch[i];
#pragma unroll
      for (int k = 1; k < KS; ++k) sum += scratch[k * total + i];
      const int c = (i / (OH * OW)) % cn;
      out[i] = prelu(sum, __ldg(alpha + c_begin + c));
    }
  }
}

// MaxPool2d(K, stride 2, ceil_mode=True) on [NC][IH][IW] -> [NC][OH][OW]  (windows clipped at the border)
template <int K, int IH, int IW>
__device__ __forceinline__ void maxpool_smem(const float* __restrict__ in, float* __restrict__ out, int nc) {
  constexpr int OH = (IH - K + 1) / 2 + 1, OW = (IW - K + 1) / 2 + 1;
  for (int i = threadIdx.x; i < nc * OH * OW; i += NT) {
    const int c = i / (OH * OW), r = i - c * (OH * OW);
    const int oy = r / OW, ox = r - oy * OW;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int y = 2 * oy + ky, x = 2 * ox + kx;
        if (y < IH && x < IW) m = fmaxf(m, in[(c * IH + y) * IW + x]);
      }
    out[i] = m;
  }
}

// Fully connected + bias + PReLU: in [G][K] smem, w [K][COUT] global, out [G][COUT] smem.  The K range is split over the
// 16 warps (each lane owns COUT/32 consecutive outputs), partial sums are combined through `scratch` [16][G][COUT].
template <int K, int COUT, int G>
__device__ __forceinline__ void fc_prelu_smem(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ scratch,
                                              const float* __restrict__ w, const float* __restrict__ bias,
                                              const float* __restrict__ alpha) {
  constexpr int V = COUT / 32, NW = NT / 32, KS = (K + NW - 1) / NW;
  static_assert(V == 4 || V == 8, "COUT must be 128 or 256");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[G][V];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[g][v] = 0.f;
  const int k0 = warp * KS, k1 = min(K, k0 + KS);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    float wv[V];
    const float4* wr = reinterpret_cast<const float4*>(w + (size_t)k * COUT + lane * V);
#pragma unroll
    for (int v4 = 0; v4 < V / 4; ++v4) {
      const float4 t = __ldg(wr + v4);
      wv[4 * v4] = t.x; wv[4 * v4 + 1] = t.y; wv[4 * v4 + 2] = t.z; wv[4 * v4 + 3] = t.w;
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float x = in[g * K + k];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[g][v] = fmaf(wv[v], x, acc[g][v]);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int v = 0; v < V; ++v) scratch[(warp * G + g) * COUT + lane * V + v] = acc[g][v];
  __syncthreads();
  for (int i = threadIdx.x; i < G * COUT; i += NT) {
    const int j = i % COUT;
    float s = __ldg(bias + j);
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) s += scratch[ww * G * COUT + i];
    out[i] = prelu(s, __ldg(alpha + j));
  }
}

// flat crop index -> (image, slot) through the exclusive scan of per-image counts
__device__ __forceinline__ void locate(const int* __restrict__ offs, int B, int flat, int& b, int& slot) {
  int lo = 0, hi = B;            // largest b with offs[b] <= flat
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offs[mid] <= flat) lo = mid; else hi = mid;
  }
  b = lo;
  slot = flat - offs[lo];
}

// ---------------------------------------------------------------------------------------------------------------
// packed weight layouts (floats); produced by models/mtcnn.py::_pack_rnet / _pack_onet
struct RW {   // R-Net
  static constexpr int W1 = 0, B1 = W1 + 27 * 28, A1 = B1 + 28;
  static constexpr int W2 = A1 + 28, B2 = W2 + 252 * 48, A2 = B2 + 48;
  static constexpr int W3 = A2 + 48, B3 = W3 + 192 * 64, A3 = B3 + 64;
  static constexpr int W4 = A3 + 64, B4 = W4 + 576 * 128, A4 = B4 + 128;
  static constexpr int W5 = A4 + 128, B5 = W5 + 128 * 8, END = B5 + 8;   // heads: cols 0-1 logits, 2-5 reg
};
struct OW_ {  // O-Net
  static constexpr int W1 = 0, B1 = W1 + 27 * 32, A1 = B1 + 32;
  static constexpr int W2 = A1 + 32, B2 = W2 + 288 * 64, A2 = B2 + 64;
  static constexpr int W3 = A2 + 64, B3 = W3 + 576 * 64, A3 = B3 + 64;
  static constexpr int W4 = A3 + 64, B4 = W4 + 256 * 128, A4 = B4 + 128;
  static constexpr int W5 = A4 + 128, B5 = W5 + 1152 * 256, A5 = B5 + 256;
  static constexpr int W6 = A5 + 256, B6 = W6 + 256 * 16, END = B6 + 16;  // heads: 0-1 logits, 2-5 reg, 6-15 landmarks
};

struct HeadArgs {
  const uint8_t* frames;
  int B, H, W, cap;
  const int* count;      // [B]
  const int4* pad;       // [B][cap]  (x, y, ex, ey)
  const int* offs;       // [B+1] exclusive scan of min(count, cap)
  const float* w;
  float* prob;           // [B][cap]
  float4* reg;           // [B][cap]
  float* lmk;            // [B][cap][10] (O-Net)
  float* crops;          // workspace [crop_cap][3][S][S]: the resized, normalised crops (written by crop_kernel)
  void* p1;              // O-Net tensor-core path: pooled conv1 map, split parts per pixel: [crop][23][23][96] bf16 (hi | mid | lo)
                         // or [crop][23][23][64] fp16 (hi | lo)
  int split_mode;        // 1 = three bf16 parts, 2 = two fp16 parts
  const float* c2;       // O-Net tensor-core path: conv2 + PReLU output [crop][21*21][64] fp32 (written by sv_conv)
  void* p3;              // O-Net tensor-core conv3: pooled conv2 map as fp16 hi | lo parts [crop][10][10][128] (nullable)
  const float* c3;       // O-Net tensor-core conv3: conv3 + PReLU output [crop][8*8][64] fp32 (nullable: conv3 on the FMA pipe)
  int crop_cap;          // crops the workspace holds; flat indices beyond it are dropped and flagged in *status (bit 5)
  int* status;
};

// Crop + area-resize + normalise of EVERY candidate of the batch, as its own high-occupancy kernel: the window sums are
// byte gathers whose latency needs many warps in flight, which the network kernels (16 warps per SM, all shared memory
// taken by activations) cannot offer -- inside rnet_kernel this phase was 31 % of the time (profiles/).  One CTA per
// crop at a time, one thread per output pixel; arithmetic identical to the reference's (exact integer sums).
//
// Mapping: one thread per output pixel over the flat (crop, pixel) index space (adjacent lanes read adjacent windows,
// so a warp-level load covers a contiguous span of the frame row).  Measured alternatives that were SLOWER on the 1080p workload: a row-task
// scheme with packed column sums and two barriers per output row (562 us vs 354 us per 3.5 k R-Net crops) and a
// warp-per-output-pixel scheme for large boxes (1 072 us).
constexpr int CROP_THREADS = 256;

template <int S>
__global__ void __launch_bounds__(CROP_THREADS) crop_kernel(const HeadArgs a) {
  const int total_raw = a.offs[a.B];
  if (total_raw > a.crop_cap && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.status, 32);
  const int total = min(total_raw, a.crop_cap);
  // flat (crop, output pixel) index space: every thread slot is used whatever the crop count, no barriers
  const long long n_out = (long long)total * (S * S);
  for (long long idx = blockIdx.x * (long long)CROP_THREADS + threadIdx.x; idx < n_out; idx += (long long)gridDim.x * CROP_THREADS) {
    const int flat = (int)(idx / (S * S));
    const int i = (int)(idx - (long long)flat * (S * S));
    int b, slot;
    locate(a.offs, a.B, flat, b, slot);
    const int4 pad = __ldg(a.pad + (size_t)b * a.cap + slot);
    const uint8_t* frame = a.frames + (size_t)b * a.H * a.W * 3;
    float* dst = a.crops + (size_t)flat * 3 * S * S;
    const int x0 = pad.x - 1, y0 = pad.y - 1;
    const int cw = pad.z - x0, ch = pad.w - y0;
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;                   // detect_face.py:110 skips empty boxes: the crop stays zero
    if (cw > 0 && ch > 0) {
      const int oy = i / S, ox = i - oy * S;
      const int ys = (oy * ch) / S, ye = ((oy + 1) * ch + S - 1) / S;
      const int xs = (ox * cw) / S, xe = ((ox + 1) * cw + S - 1) / S;
      unsigned s0 = 0, s1 = 0, s2 = 0;
      // (aligned 4-byte loads with rotating channel accumulators were measured at parity with these byte loads)
      for (int y = ys; y < ye; ++y) {
        const uint8_t* row = frame + ((size_t)(y0 + y) * a.W + (x0 + xs)) * 3;
        for (int x = 0; x < xe - xs; ++x) {
          s0 += __ldg(row + 3 * x); s1 += __ldg(row + 3 * x + 1); s2 += __ldg(row + 3 * x + 2);
        }
      }
      const float kh = (float)(ye - ys), kw = (float)(xe - xs);
      r0 = mul_rn(sub_rn(div_rn(div_rn((float)s0, kh), kw), 127.5f), 0.0078125f);
      r1 = mul_rn(sub_rn(div_rn(div_rn((float)s1, kh), kw), 127.5f), 0.0078125f);
      r2 = mul_rn(sub_rn(div_rn(div_rn((float)s2, kh), kw), 127.5f), 0.0078125f);
    }
    dst[i] = r0; dst[S * S + i] = r1; dst[2 * S * S + i] = r2;
  }
}

// Warp-per-output-row variant (the default): the byte gathers above are bound by L1 wavefronts -- a warp-level byte load
// whose lanes sit kw*3 bytes apart touches 4-5 cache lines, ~600 wavefronts per output row of a 150-pixel box.  Here one
// warp owns (crop, output row): its lanes read the source span of that row as CONSECUTIVE aligned 32-bit words (one
// wavefront per 128 bytes), keep exact column sums of their 4 byte columns over the kh window rows in two registers of
// packed u16 lanes (kh <= 257, checked on the host), park them in the warp's shared-memory slice and then add up the kw
// column sums of each of the row's 3*S outputs.  No CTA barrier; boxes wider than the slice take the byte-gather path.
constexpr int CR_WARPS = 8, CR_CAP = 2048;      // warps per CTA, byte columns per warp slice

template <int S>
__global__ void __launch_bounds__(CR_WARPS * 32) crop_rows_kernel(const HeadArgs a) {
  __shared__ __align__(16) uint16_t s_col[CR_WARPS][CR_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* col = s_col[warp];
  const int total_raw = a.offs[a.B];
  if (total_raw > a.crop_cap && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.status, 32);
  const int total = min(total_raw, a.crop_cap);
  const long long n_tasks = (long long)total * S;
  const size_t rowbytes = (size_t)a.W * 3;
  for (long long task = (long long)blockIdx.x * CR_WARPS + warp; task < n_tasks; task += (long long)gridDim.x * CR_WARPS) {
    const int flat = (int)(task / S), oy = (int)(task - (long long)flat * S);
    int b, slot;
    locate(a.offs, a.B, flat, b, slot);
    const int4 pad = __ldg(a.pad + (size_t)b * a.cap + slot);
    float* dst = a.crops + (size_t)flat * 3 * S * S + oy * S;
    const int x0 = pad.x - 1, y0 = pad.y - 1;
    const int cw = pad.z - x0, ch = pad.w - y0;
    if (cw <= 0 || ch <= 0) {                               // detect_face.py:110 skips empty boxes: the crop stays zero
      for (int t = lane; t < 3 * S; t += 32) { const int c = t / S; dst[c * S * S + (t - c * S)] = 0.f; }
      continue;
    }
    const int ys = (oy * ch) / S, ye = ((oy + 1) * ch + S - 1) / S;
    const float kh = (float)(ye - ys);
    const uint8_t* row0 = a.frames + ((size_t)b * a.H + (y0 + ys)) * rowbytes + (size_t)x0 * 3;   // first byte of the span
    const int delta = (int)((uintptr_t)row0 & 3);           // the same for every row: rowbytes % 4 == 0 (host check)
    const int span = delta + cw * 3;                        // byte columns counted from the aligned start
    if (span <= CR_CAP) {
      const uint8_t* base = row0 - delta;
      const int nwords = (span + 3) >> 2;
      for (int wi = lane; wi < nwords; wi += 32) {
        const uint8_t* p = base + 4 * (size_t)wi;
        uint32_t e = 0, o = 0;
        int y = ys;
        for (; y + 4 <= ye; y += 4) {
          uint32_t w[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) w[u] = __ldg(reinterpret_cast<const uint32_t*>(p + u * rowbytes));
#pragma unroll
          for (int u = 0; u < 4; ++u) { e += w[u] & 0x00FF00FFu; o += (w[u] >> 8) & 0x00FF00FFu; }
          p += 4 * rowbytes;
        }
        for (; y < ye; ++y) {
          const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
          e += w & 0x00FF00FFu; o += (w >> 8) & 0x00FF00FFu;
          p += rowbytes;
        }
        *reinterpret_cast<uint2*>(col + 4 * wi) = make_uint2(__byte_perm(e, o, 0x5410), __byte_perm(e, o, 0x7632));
      }
      __syncwarp();
      for (int t = lane; t < 3 * S; t += 32) {
        const int c = t / S, ox = t - c * S;
        const int xs = (ox * cw) / S, xe = ((ox + 1) * cw + S - 1) / S;
        const uint16_t* cp = col + delta + 3 * xs + c;
        unsigned sum = 0;
        for (int x = xs; x < xe; ++x, cp += 3) sum += *cp;
        dst[c * S * S + ox] = mul_rn(sub_rn(div_rn(div_rn((float)sum, kh), (float)(xe - xs)), 127.5f), 0.0078125f);
      }
      __syncwarp();
    } else {
      for (int t = lane; t < 3 * S; t += 32) {
        const int c = t / S, ox = t - c * S;
        const int xs = (ox * cw) / S, xe = ((ox + 1) * cw + S - 1) / S;
        unsigned sum = 0;
        for (int y = ys; y < ye; ++y) {
          const uint8_t* q = row0 + (size_t)(y - ys) * rowbytes + 3 * xs + c;
          for (int x = xs; x < xe; ++x, q += 3) sum += __ldg(q);
        }
        dst[c * S * S + ox] = mul_rn(sub_rn(div_rn(div_rn((float)sum, kh), (float)(xe - xs)), 127.5f), 0.0078125f);
      }
    }
  }
}

// crop stage launcher: the warp-per-row kernel when the frame layout allows aligned word loads and u16 column sums
template <int S>
void launch_crops(const HeadArgs& a, cudaStream_t st) {
  static const bool force_gather = getenv("VNFR_CROP_GATHER") != nullptr;
  const bool rows_ok = ((size_t)a.W * 3) % 4 == 0 && ((uintptr_t)a.frames % 4) == 0 && (a.H + S - 1) / S + 1 <= 257;
  if (rows_ok && !force_gather) crop_rows_kernel<S><<<148 * 6, CR_WARPS * 32, 0, st>>>(a);
  else crop_kernel<S><<<148 * 8, CROP_THREADS, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------------- R-Net
constexpr int RG = 4;     // candidates per CTA pass
constexpr int R_A = RG * 3 * 24 * 24;        // 6912   input / pool2 / fc4 out
constexpr int R_B = RG * 28 * 11 * 11;       // 13552  pooled conv1 / conv3 out
constexpr int R_C = RG * 48 * 9 * 9;         // 15552  conv1 channel-slab temp (8 ch: 15488) / conv2 out / fc scratch
constexpr int R_SMEM = (R_A + R_B + R_C) * 4;

// optional per-phase cycle counters of CTA 0 (tools/heads_probe.py); null = off
__device__ long long* g_heads_dbg = nullptr;
#define HD_MARK(i) do { if (dbg) { const long long t__ = clock64(); dbg[i] += t__ - tlast; tlast = t__; } } while (0)

__global__ void __launch_bounds__(NT, 1) rnet_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? g_heads_dbg : nullptr;
  if (dbg) dbg += 16;                          // R-Net counters live in slots 16..31
  long long tlast = dbg ? clock64() : 0;
  float* A = sm; float* Bf = sm + R_A; float* Cf = Bf + R_B;
  const int total = min(a.offs[a.B], a.crop_cap);
  const float* w = a.w;
  for (int base = blockIdx.x * RG; base < total; base += gridDim.x * RG) {
    __shared__ int s_b[RG], s_slot[RG];
    if (threadIdx.x < RG) {
      int b = 0, slot = 0;
      if (base + threadIdx.x < total) locate(a.offs, a.B, base + threadIdx.x, b, slot);
      s_b[threadIdx.x] = (base + threadIdx.x < total) ? b : -1;
      s_slot[threadIdx.x] = slot;
    }
    __syncthreads();
    HD_MARK(0);
    {
      // the RG crops of this pass are consecutive in the workspace: one coalesced 16-byte copy
      const int nvalid = min(RG, total - base) * 3 * 576;
      const float4* src = reinterpret_cast<const float4*>(a.crops + (size_t)base * 3 * 576);
      for (int i = threadIdx.x; i < RG * 3 * 576 / 4; i += NT)
        reinterpret_cast<float4*>(A)[i] = 4 * i < nvalid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    HD_MARK(1);
    // conv1 3->28 (3x3) + PReLU in slabs of 8 channels -> maxpool 3/2 ceil -> Bf [RG][28][11][11]
    for (int c0 = 0; c0 < 28; c0 += 8) {
      const int c1 = min(28, c0 + 8), cn = c1 - c0;
      conv_prelu_smem<3, 28, 3, 3, 24, 24, RG, 4, 8, 1>(A, Cf, nullptr, w + RW::W1, w + RW::B1, w + RW::A1, c0, c1);   // Cf [RG][cn][22][22]
      __syncthreads();
      HD_MARK(2);
      for (int g = 0; g < RG; ++g) {
        // pool channel slab of candidate g into its place in Bf
        constexpr int OH = 11;
        for (int i = threadIdx.x; i < cn * OH * OH; i += NT) {
          const int c = i / (OH * OH), r = i - c * (OH * OH);
          const int oy = r / OH, ox = r - oy * OH;
          float m = -CUDART_INF_F;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int y = 2 * oy + ky, x = 2 * ox + kx;
              if (y < 22 && x < 22) m = fmaxf(m, Cf[((g * cn + c) * 22 + y) * 22 + x]);
            }
          Bf[((g * 28 + c0 + c) * OH + oy) * OH + ox] = m;
        }
      }
      __syncthreads();
      HD_MARK(3);
    }
    conv_prelu_smem_ws<28, 48, 3, 3, 11, 11, RG, 8, 4, 1, 4>(Bf, Cf, nullptr, A, w + RW::W2, w + RW::B2, w + RW::A2, 0, 48);   // Cf [RG][48][9][9]
    __syncthreads();
    HD_MARK(4);
    maxpool_smem<3, 9, 9>(Cf, A, RG * 48);                                                            // A  [RG][48][4][4]
    __syncthreads();
    HD_MARK(5);
    conv_prelu_smem<48, 64, 2, 2, 4, 4, RG, 4, 3, 2>(A, Bf, Cf, w + RW::W3, w + RW::B3, w + RW::A3, 0, 64);   // Bf [RG][64][3][3] = [RG][576]
    __syncthreads();
    HD_MARK(6);
    fc_prelu_smem<576, 128, RG>(Bf, A, Cf, w + RW::W4, w + RW::B4, w + RW::A4);                        // A  [RG][128]
    __syncthreads();
    HD_MARK(7);
    if (threadIdx.x < RG * 8) {
      const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
      float s = __ldg(w + RW::B5 + j);
      for (int k = 0; k < 128; ++k) s = fmaf(__ldg(w + RW::W5 + k * 8 + j), A[g * 128 + k], s);
      Cf[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x < RG && s_b[threadIdx.x] >= 0) {
      const int g = threadIdx.x;
      const float l0 = Cf[g * 8], l1 = Cf[g * 8 + 1];
      const float mx = fmaxf(l0, l1);
      const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
      const size_t o = (size_t)s_b[g] * a.cap + s_slot[g];
      const int4 pd = a.pad[o];
      const bool empty = !(pd.w > pd.y - 1 && pd.z > pd.x - 1);      // detect_face.py:110 would skip this crop
      a.prob[o] = empty ? 0.f : e1 / (e0 + e1);
      a.reg[o] = make_float4(Cf[g * 8 + 2], Cf[g * 8 + 3], Cf[g * 8 + 4], Cf[g * 8 + 5]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ R-Net, tensor-core conv2
// conv2 (28 -> 48, 3x3) is 64 % of R-Net's FLOPs and took 40 % of rnet_kernel.  Same scheme as O-Net below: two fp16
// parts per fp32 operand, three products, fp32 accumulation in TMEM (sv_conv.cu, VnfrConvOp.split3 = 2), all crops of the
// batch in one launch:
//   rnet_front_kernel: crop -> conv1 + PReLU -> maxpool 3/2 -> fp16 split, NHWC [crop][11][11][hi 32 | lo 32] (28 real channels)
//   sv_conv_kernel   : conv2 + bias + PReLU -> fp32 NHWC [crop][81][48]
//   rnet_back_kernel : maxpool 3/2 -> conv3 -> dense4 -> heads (small buffers: two CTAs per SM)
__global__ void __launch_bounds__(NT, 1) rnet_front_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm; float* Bf = sm + R_A; float* Cf = Bf + R_B;
  const int total = min(a.offs[a.B], a.crop_cap);
  const float* w = a.w;
  unsigned short* p1 = reinterpret_cast<unsigned short*>(a.p1);
  for (int base = blockIdx.x * RG; base < total; base += gridDim.x * RG) {
    __syncthreads();
    {
      const int nvalid = min(RG, total - base) * 3 * 576;
      const float4* src = reinterpret_cast<const float4*>(a.crops + (size_t)base * 3 * 576);
      for (int i = threadIdx.x; i < RG * 3 * 576 / 4; i += NT)
        reinterpret_cast<float4*>(A)[i] = 4 * i < nvalid ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int c0 = 0; c0 < 28; c0 += 8) {
      const int c1 = min(28, c0 + 8), cn = c1 - c0;
      conv_prelu_smem<3, 28, 3, 3, 24, 24, RG, 4, 8, 1>(A, Cf, nullptr, w + RW::W1, w + RW::B1, w + RW::A1, c0, c1);   // Cf [RG][cn][22][22]
      __syncthreads();
      for (int g = 0; g < RG; ++g) {
        constexpr int OH = 11;
        for (int i = threadIdx.x; i < cn * OH * OH; i += NT) {
          const int c = i / (OH * OH), r = i - c * (OH * OH);
          const int oy = r / OH, ox = r - oy * OH;
          float m = -CUDART_INF_F;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int y = 2 * oy + ky, x = 2 * ox + kx;
              if (y < 22 && x < 22) m = fmaxf(m, Cf[((g * cn + c) * 22 + y) * 22 + x]);
            }
          Bf[((g * 28 + c0 + c) * OH + oy) * OH + ox] = m;
        }
      }
      __syncthreads();
    }
    // two-part fp16 split, pixel-major, channel fastest (channels 28..31 of each part are zero padding)
    const int ng = min(RG, total - base);
    for (int i = threadIdx.x; i < ng * 121 * 32; i += NT) {
      const int c = i & 31, t = i >> 5;
      const int g = t / 121, px = t - g * 121;
      const float x = c < 28 ? Bf[(g * 28 + c) * 121 + px] : 0.f;
      const __half hi = __float2half_rn(x);
      const __half lo = __float2half_rn(x - __half2float(hi));
      unsigned short* q = p1 + ((size_t)(base + g) * 121 + px) * 64 + c;
      q[0] = __half_as_ushort(hi); q[32] = __half_as_ushort(lo);
    }
  }
}

// ---- R-Net / O-Net front with conv1 (3 -> 28 / 32, 3x3) on the tensor cores as well (default; VNFR_RNET_FRONT_FMA=1 /
// VNFR_ONET_FRONT_FMA=1 select the FMA front kernels).  The FMA versions are bound by shared-memory loads (ncu: 70 % LSU
// wavefronts, 30 % FMA): K = 27 gives every loaded activation only 4-8 FMAs.  Here the crop becomes an "x-im2col" A
// operand in shared memory: one 32-byte swizzled row per pixel (y, x) holding k = 3 dx + c -> crop[c][y][x + dx] (9 of 16
// fp16, hi and lo parts in two buffers); tap ky reads the SAME buffer from a start address shifted by S ky rows (the
// shifted-view trick), so a band of conv rows is MT row tiles x 3 taps x 3 products (hi*hi, hi*lo, lo*hi) tcgen05.mma of
// M 128, N 32, K 16 into MT TMEM accumulators.  The 8 warps read their TMEM lanes back (+ bias, PReLU) into a
// shared-memory map (aliasing the A buffers), max-pool 3/2 and write the two-part fp16 split that conv2's shifted-view
// convolution reads.  R-Net: the whole 22x22 map is one band (5 tiles); O-Net: bands of 4 pooled rows = 9 conv rows (4
// tiles, one conv row recomputed per band).  Two CTAs per SM.
constexpr int RT_THREADS = 512;         // 2 CTAs x 16 warps per SM: every phase between the barriers is latency-bound

template <int S, int COUT, int PB, int W1, int B1, int A1>
struct FrontTc {
  static constexpr int OH = S - 2;                             // conv1 output edge
  static constexpr int PH = (OH - 3 + 1) / 2 + 1;              // pooled edge (3/2, ceil mode)
  static constexpr int CB = (2 * PB + 1 < OH) ? 2 * PB + 1 : OH;   // conv rows per band
  static constexpr int MT = ((CB - 1) * S + OH + 127) / 128;   // accumulator row tiles per band (raster of width S)
  static constexpr int A_ROWS = MT * 128 + 2 * S;              // rows the last accumulator rows reach through the ky shifts
  static constexpr int A_PART = (A_ROWS * 32 + 255) / 256 * 256;
  static constexpr int MAP_BYTES = CB * OH * 32 * 4;
  static constexpr int REGION = ((2 * A_PART > MAP_BYTES ? 2 * A_PART : MAP_BYTES) + 1023) / 1024 * 1024;
  static constexpr int B_BYTES = 3 * 2 * 1024;                 // [ky][part][32 cout rows][16 k] fp16, 32-byte swizzled rows
  static constexpr int SMEM = 1024 + B_BYTES + REGION + 3 * S * S * 4;
  static constexpr unsigned TCOLS = MT * 32 <= 128 ? 128u : 256u;
};

template <int S, int COUT, int PB, int W1, int B1, int A1>
__global__ void __launch_bounds__(RT_THREADS, 2) head_front_tc_kernel(const HeadArgs a) {
  using F = FrontTc<S, COUT, PB, W1, B1, A1>;
  constexpr int OH = F::OH, PH = F::PH, MT = F::MT;
  extern __shared__ __align__(16) uint8_t rt_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float4 s_ba[16];                                    // conv1 bias (8 quads) and PReLU slopes (8 quads)
  uint8_t* base = rt_raw + ((1024u - (smem_u32(rt_raw) & 1023u)) & 1023u);
  uint8_t* s_b = base;                                           // B operand
  uint8_t* s_reg = base + F::B_BYTES;                            // A parts, later the conv1 map [rows][OH][32] fp32
  float* s_crop = reinterpret_cast<float*>(base + F::B_BYTES + F::REGION);   // [3][S][S]
  float* s_map = reinterpret_cast<float*>(s_reg);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_hi = smem_u32(s_reg), a_lo = a_hi + F::A_PART, b_addr = smem_u32(s_b), bar = smem_u32(&s_bar);
  const float* w = a.w;
  // ---- once per CTA: B operand (conv1 weights as fp16 hi / lo, k = 3 dx + c), barrier, TMEM
  for (int i = tid; i < 3 * 32 * 16; i += RT_THREADS) {
    const int k = i & 15, n = (i >> 4) & 31, ky = i >> 9;
    float wv = 0.f;
    if (k < 9 && n < COUT) { const int dx = k / 3, c = k - 3 * dx; wv = __ldg(w + W1 + ((c * 3 + ky) * 3 + dx) * COUT + n); }
    const __half hi = __float2half_rn(wv);
    const __half lo = __float2half_rn(wv - __half2float(hi));
    const int off = n * 32 + (((k >> 3) ^ ((n >> 2) & 1)) << 4) + 2 * (k & 7);
    *reinterpret_cast<unsigned short*>(s_b + (ky * 2 + 0) * 1024 + off) = __half_as_ushort(hi);
    *reinterpret_cast<unsigned short*>(s_b + (ky * 2 + 1) * 1024 + off) = __half_as_ushort(lo);
  }
  if (tid < 64) {
    const int ch = tid & 31;
    reinterpret_cast<float*>(s_ba)[tid] = ch < COUT ? __ldg(w + (tid < 32 ? B1 : A1) + ch) : 0.f;
  }
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(F::TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  uint32_t phase = 0;
  const int total = min(a.offs[a.B], a.crop_cap);
  unsigned short* p1 = reinterpret_cast<unsigned short*>(a.p1);
  for (int flat = blockIdx.x; flat < total; flat += gridDim.x) {
    // ---- crop -> shared memory
    {
      const float4* src = reinterpret_cast<const float4*>(a.crops + (size_t)flat * 3 * S * S);
      for (int i = tid; i < 3 * S * S / 4; i += RT_THREADS) reinterpret_cast<float4*>(s_crop)[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int P0 = 0; P0 < PH; P0 += PB) {
      const int P1 = min(PH, P0 + PB);
      const int y0 = 2 * P0, y1 = min(OH, 2 * (P1 - 1) + 3), ncr = y1 - y0;      // conv rows [y0, y1) of this band
      // ---- x-im2col A operand: band pixel p = S yy + x -> 16 halves, k = 3 dx + c (k >= 9 and x + dx >= S: zero)
      for (int pix = tid; pix < (ncr + 2) * S; pix += RT_THREADS) {
        const int x = pix % S;
        const float* src = s_crop + y0 * S + pix;
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int dx = k / 3, c = k - 3 * dx;
          v[k] = (k < 9 && x + dx < S) ? src[c * S * S + dx] : 0.f;
        }
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const __half2 h = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
          const __half2 l = __floats2half2_rn(v[2 * q] - __low2float(h), v[2 * q + 1] - __high2float(h));
          hw[q] = *reinterpret_cast<const uint32_t*>(&h);
          lw[q] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const uint32_t row = a_hi + 32u * (uint32_t)pix;
        const uint32_t sw = ((row >> 7) & 1u) << 4;
        tc::sts128(row + sw, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        tc::sts128(row + (sw ^ 16u), make_uint4(hw[4], hw[5], hw[6], hw[7]));
        tc::sts128(row + F::A_PART + sw, make_uint4(lw[0], lw[1], lw[2], lw[3]));
        tc::sts128(row + F::A_PART + (sw ^ 16u), make_uint4(lw[4], lw[5], lw[6], lw[7]));
      }
      tc::fence_proxy_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      // ---- conv1 of the band on the tensor cores
      if (warp == 0) {
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t idesc = tc::make_idesc_f16(32, 1);
#pragma unroll 1
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {
              const uint32_t shift = 32u * (uint32_t)(128 * mt + S * ky);
              const uint64_t ah = tc::make_sw_desc(a_hi + shift, 32, 0), al = tc::make_sw_desc(a_lo + shift, 32, 0);
              const uint64_t bh = tc::make_sw_desc(b_addr + (uint32_t)(ky * 2) * 1024u, 32, 0);
              const uint64_t bl = tc::make_sw_desc(b_addr + (uint32_t)(ky * 2 + 1) * 1024u, 32, 0);
              tc::umma_bf16(tmem_base + 32u * mt, ah, bh, idesc, ky != 0);
              tc::umma_bf16(tmem_base + 32u * mt, ah, bl, idesc, 1);
              tc::umma_bf16(tmem_base + 32u * mt, al, bh, idesc, 1);
            }
          }
          tc::umma_commit(bar);
        }
        __syncwarp();
      }
      tc::mbar_wait(bar, phase);
      phase ^= 1u;
      tc::tc_fence_after();
      __syncthreads();               // every thread has seen the MMAs complete: the A buffers may be overwritten by the map
      // ---- TMEM -> + bias, PReLU -> conv1 map [ncr][OH][32] fp32 (warp group g = warp >> 2 takes row tiles g, g + 4, ...)
      {
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int mt = warp >> 2; mt < MT; mt += RT_THREADS / 128) {
          float acc[32];
          __syncwarp();
          tc::tmem_ld16_issue(t_lane + 32u * mt, acc);
          tc::tmem_ld16_issue(t_lane + 32u * mt + 16u, acc + 16);
          tc::tmem_ld_wait(acc);
          tc::tmem_ld_wait(acc + 16);
          const int r = 128 * mt + (warp & 3) * 32 + lane;
          const int yy = r / S, x = r - S * yy;
          if (yy < ncr && x < OH) {
            // map row = pixel, 8 float4 slots XOR-swizzled by the pixel index: lanes (= consecutive pixels) are 128 B apart,
            // unswizzled their 128-bit stores all hit the same four banks (ncu: 72 % of the kernel's shared wavefronts)
            const int mrow = yy * OH + x;
            float4* dst = reinterpret_cast<float4*>(s_map + mrow * 32);
#pragma unroll
            for (int q = 0; q < COUT / 4; ++q) {
              const float4 bb = s_ba[q], aa = s_ba[8 + q];
              dst[q ^ (mrow & 7)] = make_float4(prelu(acc[4 * q] + bb.x, aa.x), prelu(acc[4 * q + 1] + bb.y, aa.y),
                                                prelu(acc[4 * q + 2] + bb.z, aa.z), prelu(acc[4 * q + 3] + bb.w, aa.w));
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncthreads();
      // ---- maxpool 3/2 (ceil: windows clipped at OH) of pooled rows [P0, P1) + two-part fp16 split
      //      -> p1 [crop][PH*PH][hi 32 | lo 32] (COUT real channels)
      //      (a thread takes four channels: 9 x LDS.128 per 4 outputs)
      for (int i = tid; i < (P1 - P0) * PH * 8; i += RT_THREADS) {
        const int q = i & 7, pp = i >> 3;
        const int oyl = pp / PH, ox = pp - PH * oyl;
        const int oy = P0 + oyl;
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * q < COUT) {
          m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int y = 2 * oy + ky, x = 2 * ox + kx;
              if (y < OH && x < OH) {
                const int mrow = (y - y0) * OH + x;
                const float4 v = *reinterpret_cast<const float4*>(s_map + mrow * 32 + 4 * (q ^ (mrow & 7)));
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
              }
            }
        }
        const __half2 h0 = __floats2half2_rn(m.x, m.y), h1 = __floats2half2_rn(m.z, m.w);
        const __half2 l0 = __floats2half2_rn(m.x - __low2float(h0), m.y - __high2float(h0));
        const __half2 l1 = __floats2half2_rn(m.z - __low2float(h1), m.w - __high2float(h1));
        unsigned short* d = p1 + ((size_t)flat * (PH * PH) + oy * PH + ox) * 64 + 4 * q;
        *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(d + 32) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
      }
      __syncthreads();               // the map region is rebuilt as the A operand of the next band / crop
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F::TCOLS) : "memory");
  }
}

// R-Net: 24x24 crops, 28 channels, one band; O-Net: 48x48 crops, 32 channels, bands of 4 pooled rows
#define RNET_FRONT_TC 24, 28, 11, RW::W1, RW::B1, RW::A1
#define ONET_FRONT_TC 48, 32, 4, OW_::W1, OW_::B1, OW_::A1

constexpr int RB_A = RG * 48 * 16, RB_B = RG * 576, RB_C = 16 * RG * 128;       // pool2 / dense4 out, conv3 out, scratch
constexpr int RB_SMEM = (RB_A + RB_B + RB_C) * 4;

__global__ void __launch_bounds__(NT, 2) rnet_back_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm; float* Bf = sm + RB_A; float* Cf = Bf + RB_B;
  const int total = min(a.offs[a.B], a.crop_cap);
  const float* w = a.w;
  for (int base = blockIdx.x * RG; base < total; base += gridDim.x * RG) {
    __shared__ int s_b[RG], s_slot[RG];
    __syncthreads();
    if (threadIdx.x < RG) {
      int b = 0, slot = 0;
      if (base + threadIdx.x < total) locate(a.offs, a.B, base + threadIdx.x, b, slot);
      s_b[threadIdx.x] = (base + threadIdx.x < total) ? b : -1;
      s_slot[threadIdx.x] = slot;
    }
    // maxpool 3/2 (9 -> 4: every window is complete) straight from the fp32 NHWC conv2 output; channel fastest
    for (int i = threadIdx.x; i < RG * 16 * 48; i += NT) {
      const int c = i % 48, t = i / 48;
      const int g = t >> 4, pos = t & 15;
      const int oy = pos >> 2, ox = pos & 3;
      float m = 0.f;
      if (base + g < total) {
        const float* src = a.c2 + ((size_t)(base + g) * 81 + (2 * oy) * 9 + 2 * ox) * 48 + c;
        m = -CUDART_INF_F;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) m = fmaxf(m, __ldg(src + (ky * 9 + kx) * 48));
      }
      A[(g * 48 + c) * 16 + pos] = m;
    }
    __syncthreads();
    conv_prelu_smem<48, 64, 2, 2, 4, 4, RG, 4, 3, 2>(A, Bf, Cf, w + RW::W3, w + RW::B3, w + RW::A3, 0, 64);   // Bf [RG][64][3][3] = [RG][576]
    __syncthreads();
    fc_prelu_smem<576, 128, RG>(Bf, A, Cf, w + RW::W4, w + RW::B4, w + RW::A4);                        // A  [RG][128]
    __syncthreads();
    if (threadIdx.x < RG * 8) {
      const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
      float sacc = __ldg(w + RW::B5 + j);
      for (int k = 0; k < 128; ++k) sacc = fmaf(__ldg(w + RW::W5 + k * 8 + j), A[g * 128 + k], sacc);
      Cf[threadIdx.x] = sacc;
    }
    __syncthreads();
    if (threadIdx.x < RG && s_b[threadIdx.x] >= 0) {
      const int g = threadIdx.x;
      const float l0 = Cf[g * 8], l1 = Cf[g * 8 + 1];
      const float mx = fmaxf(l0, l1);
      const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
      const size_t o = (size_t)s_b[g] * a.cap + s_slot[g];
      const int4 pd = a.pad[o];
      const bool empty = !(pd.w > pd.y - 1 && pd.z > pd.x - 1);      // detect_face.py:110 would skip this crop
      a.prob[o] = empty ? 0.f : e1 / (e0 + e1);
      a.reg[o] = make_float4(Cf[g * 8 + 2], Cf[g * 8 + 3], Cf[g * 8 + 4], Cf[g * 8 + 5]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------- O-Net
constexpr int OG = 4;                        // crops whose dense5 + heads run together (dense5's 1.18 MB of weights are
                                             // then read from L2 once per 4 crops instead of once per crop)
constexpr int O_A = 3 * 48 * 48;             // 6912   input / pool2 out (6400) / pool3 out / fc5 out [OG][256]
constexpr int O_B = 32 * 23 * 23;            // 16928  pooled conv1 / conv3 out (4096)
constexpr int O_C = 64 * 21 * 21;            // 28224  conv1 slab temp (8 ch: 16928) / conv2 out / split-K + fc scratch
constexpr int O_F = OG * 1152;               // 4608   conv4 outputs of the group = dense5 inputs
constexpr int O_SMEM = (O_A + O_B + O_C + O_F) * 4;

__global__ void __launch_bounds__(NT, 1) onet_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? g_heads_dbg : nullptr;
  long long tlast = dbg ? clock64() : 0;
  float* A = sm; float* Bf = sm + O_A; float* Cf = Bf + O_B; float* F = Cf + O_C;
  __shared__ int s_b[OG], s_slot[OG], s_empty[OG];
  const int total = min(a.offs[a.B], a.crop_cap);
  const float* w = a.w;
  // crop k of this CTA is flat index blockIdx.x + k*gridDim.x (balanced to +-1 crop per CTA)
  for (int k0 = 0; blockIdx.x + k0 * gridDim.x < total; k0 += OG) {
    for (int g = 0; g < OG; ++g) {
      const int flat = blockIdx.x + (k0 + g) * gridDim.x;
      if (flat >= total) {
        if (threadIdx.x == 0) s_b[g] = -1;
        for (int i = threadIdx.x; i < 1152; i += NT) F[g * 1152 + i] = 0.f;
        continue;                              // uniform across the CTA
      }
      if (threadIdx.x == 0) { int b, slot; locate(a.offs, a.B, flat, b, slot); s_b[g] = b; s_slot[g] = slot; }
      __syncthreads();
      const size_t o = (size_t)s_b[g] * a.cap + s_slot[g];
      const int4 pd = a.pad[o];
      if (threadIdx.x == 0) s_empty[g] = !(pd.w > pd.y - 1 && pd.z > pd.x - 1);
      HD_MARK(0);
      {
        const float4* src = reinterpret_cast<const float4*>(a.crops + (size_t)flat * O_A);
        for (int i = threadIdx.x; i < O_A / 4; i += NT) reinterpret_cast<float4*>(A)[i] = __ldg(src + i);
      }
      __syncthreads();
      HD_MARK(1);
      for (int c0 = 0; c0 < 32; c0 += 8) {
        conv_prelu_smem_ws<3, 32, 3, 3, 48, 48, 1, 8, 5, 1, 3>(A, Cf, nullptr, Cf + 8 * 46 * 46, w + OW_::W1, w + OW_::B1, w + OW_::A1, c0, c0 + 8);   // Cf [8][46][46]
        __syncthreads();
        HD_MARK(2);
        maxpool_smem<3, 46, 46>(Cf, Bf + c0 * 23 * 23, 8);                                               // Bf [32][23][23]
        __syncthreads();
        HD_MARK(3);
      }
      // weights staged through A (the crop is no longer needed): 2 x 4 channels x 9 taps x 64 = 2 x 9 216 B
      conv_prelu_smem_ws<32, 64, 3, 3, 23, 23, 1, 8, 7, 1, 4>(Bf, Cf, nullptr, A, w + OW_::W2, w + OW_::B2, w + OW_::A2, 0, 64);   // Cf [64][21][21]
      __syncthreads();
      HD_MARK(4);
      maxpool_smem<3, 21, 21>(Cf, A, 64);                                                                // A  [64][10][10]
      __syncthreads();
      HD_MARK(5);
      conv_prelu_smem_ws<64, 64, 3, 3, 10, 10, 1, 8, 4, 4, 2>(A, Bf, Cf, Cf + 16384, w + OW_::W3, w + OW_::B3, w + OW_::A3, 0, 64);  // Bf [64][8][8]
      __syncthreads();
      HD_MARK(6);
      maxpool_smem<2, 8, 8>(Bf, A, 64);                                                                  // A  [64][4][4]
      __syncthreads();
      HD_MARK(7);
      conv_prelu_smem<64, 128, 2, 2, 4, 4, 1, 4, 3, 4>(A, F + g * 1152, Cf, w + OW_::W4, w + OW_::B4, w + OW_::A4, 0, 128);  // F[g] [128][3][3]
      __syncthreads();
      HD_MARK(8);
    }
    __syncthreads();
    fc_prelu_smem<1152, 256, OG>(F, A, Cf, w + OW_::W5, w + OW_::B5, w + OW_::A5);                        // A  [OG][256]
    __syncthreads();
    HD_MARK(9);
    if (threadIdx.x < OG * 16) {
      const int g = threadIdx.x >> 4, j = threadIdx.x & 15;
      float sacc = __ldg(w + OW_::B6 + j);
      for (int k = 0; k < 256; ++k) sacc = fmaf(__ldg(w + OW_::W6 + k * 16 + j), A[g * 256 + k], sacc);
      Cf[threadIdx.x] = sacc;
    }
    __syncthreads();
    if (threadIdx.x < OG * 16) {
      const int g = threadIdx.x >> 4, j = threadIdx.x & 15;
      if (s_b[g] >= 0) {
        const size_t o = (size_t)s_b[g] * a.cap + s_slot[g];
        const float* h = Cf + g * 16;
        if (j == 0) {
          const float l0 = h[0], l1 = h[1];
          const float mx = fmaxf(l0, l1);
          const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
          a.prob[o] = s_empty[g] ? 0.f : e1 / (e0 + e1);
          a.reg[o] = make_float4(h[2], h[3], h[4], h[5]);
        }
        if (j >= 6) a.lmk[o * 10 + (j - 6)] = h[j];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ O-Net, tensor-core conv2
// conv2 (32 -> 64, 3x3) is 63 % of O-Net's FLOPs; on the fp32 FMA pipe it ran at 64 % of peak and still took half of the
// kernel.  Here it runs on the tensor cores in split precision (three bf16 parts per operand, six products, fp32
// accumulation -- sv_conv.cu, VnfrConvOp.split3) over ALL crops of the batch in one launch:
//   onet_front_kernel: crop -> conv1 + PReLU -> maxpool 3/2 -> 3-way bf16 split, NHWC [crop][23][23][96]      (global)
//   sv_conv_kernel   : conv2 + bias + PReLU -> fp32 NHWC [crop][441][64]                                      (global)
//   onet_back_kernel : maxpool 3/2 -> conv3 -> maxpool 2/2 -> conv4 -> dense5 (4 crops per pass) -> heads
constexpr int OF_A = 3 * 48 * 48, OF_B = 32 * 23 * 23, OF_C = 8 * 46 * 46 + 2 * 864;
constexpr int OF_SMEM = (OF_A + OF_B + OF_C) * 4;

__device__ __forceinline__ unsigned short bf16_bits(float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  return *reinterpret_cast<const unsigned short*>(&h);
}
__device__ __forceinline__ float bf16_to_float(unsigned short b) { return __uint_as_float((unsigned)b << 16); }

__global__ void __launch_bounds__(NT, 1) onet_front_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm; float* Bf = sm + OF_A; float* Cf = Bf + OF_B;
  const int total = min(a.offs[a.B], a.crop_cap);
  const float* w = a.w;
  for (int flat = blockIdx.x; flat < total; flat += gridDim.x) {
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(a.crops + (size_t)flat * OF_A);
      for (int i = threadIdx.x; i < OF_A / 4; i += NT) reinterpret_cast<float4*>(A)[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int c0 = 0; c0 < 32; c0 += 8) {
      conv_prelu_smem_ws<3, 32, 3, 3, 48, 48, 1, 8, 5, 1, 3>(A, Cf, nullptr, Cf + 8 * 46 * 46, w + OW_::W1, w + OW_::B1, w + OW_::A1, c0, c0 + 8);
      __syncthreads();
      maxpool_smem<3, 46, 46>(Cf, Bf + c0 * 23 * 23, 8);
      __syncthreads();
    }
    // split into 16-bit parts, pixel-major: thread -> (pixel, channel) with the channel fastest (64-byte runs per part)
    if (a.split_mode == 2) {
      // two fp16 parts: x = hi + lo to 2^-22 relative (residuals below 6e-5 land on fp16 subnormals, absolute step 6e-8)
      unsigned short* dst = reinterpret_cast<unsigned short*>(a.p1) + (size_t)flat * 529 * 64;
      for (int i = threadIdx.x; i < 529 * 32; i += NT) {
        const int px = i >> 5, c = i & 31;
        const float x = Bf[c * 529 + px];
        const __half hi = __float2half_rn(x);
        const __half lo = __float2half_rn(x - __half2float(hi));
        unsigned short* q = dst + (size_t)px * 64 + c;
        q[0] = __half_as_ushort(hi); q[32] = __half_as_ushort(lo);
      }
    } else {
      unsigned short* dst = reinterpret_cast<unsigned short*>(a.p1) + (size_t)flat * 529 * 96;
      for (int i = threadIdx.x; i < 529 * 32; i += NT) {
        const int px = i >> 5, c = i & 31;
        const float x = Bf[c * 529 + px];
        const unsigned short hi = bf16_bits(x);
        const float r1 = x - bf16_to_float(hi);
        const unsigned short mid = bf16_bits(r1);
        const unsigned short lo = bf16_bits(r1 - bf16_to_float(mid));
        unsigned short* q = dst + (size_t)px * 96 + c;
        q[0] = hi; q[32] = mid; q[64] = lo;
      }
    }
  }
}

constexpr int OB_A = 64 * 10 * 10 + 512, OB_B = 64 * 8 * 8, OB_C = 16384 + 9216;
constexpr int OB_SMEM = (OB_A + OB_B + OB_C + O_F) * 4;

// maxpool 3/2 (21 -> 10: every window is complete) of the conv2 output + two-part fp16 split, NHWC [crop][10][10][hi 64 | lo 64]:
// the input of conv3 on the tensor cores (64 -> 64, 3x3: 18 % of O-Net's FLOPs, 0.6 ms on the FMA pipe)
__global__ void __launch_bounds__(256) onet_mid_kernel(const HeadArgs a) {
  const int total = min(a.offs[a.B], a.crop_cap);
  const long long n = (long long)total * 100 * 64;
  unsigned short* p3 = reinterpret_cast<unsigned short*>(a.p3);
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < n; idx += (long long)gridDim.x * 256) {
    const int c = (int)(idx & 63);
    const long long t = idx >> 6;
    const int flat = (int)(t / 100), pos = (int)(t - (long long)flat * 100);
    const int oy = pos / 10, ox = pos - oy * 10;
    const float* src = a.c2 + (size_t)flat * 441 * 64 + c;
    float m = -CUDART_INF_F;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) m = fmaxf(m, __ldg(src + (size_t)((2 * oy + ky) * 21 + 2 * ox + kx) * 64));
    const __half hi = __float2half_rn(m);
    const __half lo = __float2half_rn(m - __half2float(hi));
    unsigned short* q = p3 + (size_t)t * 128 + c;
    q[0] = __half_as_ushort(hi); q[64] = __half_as_ushort(lo);
  }
}

// TC3: conv3 + PReLU already done on the tensor cores (a.c3); this kernel starts at the 2/2 max-pool
template <bool TC3>
__global__ void __launch_bounds__(NT, 1) onet_back_kernel(const HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm; float* Bf = sm + OB_A; float* Cf = Bf + OB_B; float* F = Cf + OB_C;
  __shared__ int s_b[OG], s_slot[OG], s_empty[OG];
  const int total = min(a.offs[a.B], a.crop_cap);
  const float* w = a.w;
  for (int k0 = 0; blockIdx.x + k0 * gridDim.x < total; k0 += OG) {
    for (int g = 0; g < OG; ++g) {
      const int flat = blockIdx.x + (k0 + g) * gridDim.x;
      if (flat >= total) {
        if (threadIdx.x == 0) s_b[g] = -1;
        for (int i = threadIdx.x; i < 1152; i += NT) F[g * 1152 + i] = 0.f;
        continue;
      }
      if (threadIdx.x == 0) {
        int b, slot; locate(a.offs, a.B, flat, b, slot); s_b[g] = b; s_slot[g] = slot;
        const int4 pd = a.pad[(size_t)b * a.cap + slot];
        s_empty[g] = !(pd.w > pd.y - 1 && pd.z > pd.x - 1);
      }
      if (TC3) {
        // maxpool 2/2 (8 -> 4) straight from the fp32 NHWC conv3 output; channel fastest
        const float* src = a.c3 + (size_t)flat * 64 * 64;
        for (int i = threadIdx.x; i < 16 * 64; i += NT) {
          const int pos = i >> 6, c = i & 63;
          const int oy = pos >> 2, ox = pos & 3;
          const float* q = src + (size_t)((2 * oy) * 8 + 2 * ox) * 64 + c;
          A[c * 16 + pos] = fmaxf(fmaxf(__ldg(q), __ldg(q + 64)), fmaxf(__ldg(q + 8 * 64), __ldg(q + 9 * 64)));
        }
        __syncthreads();
      } else {
        // maxpool 3/2 (21 -> 10: every window is complete) straight from the fp32 NHWC conv2 output; channel fastest
        const float* src = a.c2 + (size_t)flat * 441 * 64;
        for (int i = threadIdx.x; i < 100 * 64; i += NT) {
          const int pos = i >> 6, c = i & 63;
          const int oy = pos / 10, ox = pos - oy * 10;
          float m = -CUDART_INF_F;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) m = fmaxf(m, __ldg(src + (size_t)((2 * oy + ky) * 21 + 2 * ox + kx) * 64 + c));
          A[c * 100 + pos] = m;
        }
        __syncthreads();
        conv_prelu_smem_ws<64, 64, 3, 3, 10, 10, 1, 8, 4, 4, 2>(A, Bf, Cf, Cf + 16384, w + OW_::W3, w + OW_::B3, w + OW_::A3, 0, 64);
        __syncthreads();
        maxpool_smem<2, 8, 8>(Bf, A, 64);
        __syncthreads();
      }
      conv_prelu_smem<64, 128, 2, 2, 4, 4, 1, 4, 3, 4>(A, F + g * 1152, Cf, w + OW_::W4, w + OW_::B4, w + OW_::A4, 0, 128);
      __syncthreads();
    }
    __syncthreads();
    fc_prelu_smem<1152, 256, OG>(F, A, Cf, w + OW_::W5, w + OW_::B5, w + OW_::A5);
    __syncthreads();
    if (threadIdx.x < OG * 16) {
      const int g = threadIdx.x >> 4, j = threadIdx.x & 15;
      float sacc = __ldg(w + OW_::B6 + j);
      for (int k = 0; k < 256; ++k) sacc = fmaf(__ldg(w + OW_::W6 + k * 16 + j), A[g * 256 + k], sacc);
      Cf[threadIdx.x] = sacc;
    }
    __syncthreads();
    if (threadIdx.x < OG * 16) {
      const int g = threadIdx.x >> 4, j = threadIdx.x & 15;
      if (s_b[g] >= 0) {
        const size_t o = (size_t)s_b[g] * a.cap + s_slot[g];
        const float* h = Cf + g * 16;
        if (j == 0) {
          const float l0 = h[0], l1 = h[1];
          const float mx = fmaxf(l0, l1);
          const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
          a.prob[o] = s_empty[g] ? 0.f : e1 / (e0 + e1);
          a.reg[o] = make_float4(h[2], h[3], h[4], h[5]);
        }
        if (j >= 6) a.lmk[o * 10 + (j - 6)] = h[j];
      }
    }
    __syncthreads();
  }
}

__global__ void scan_counts_kernel(const int* __restrict__ count, int B, int cap, int* __restrict__ offs) {
  // B is small (frames per batch): a single thread does the exclusive scan
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int s = 0;
    for (int b = 0; b < B; ++b) { offs[b] = s; s += min(count[b], cap); }
    offs[B] = s;
  }
}

}  // namespace

// heads_chain.cu
int vnfr_heads_back_run(int onet, const VnfrHeadsBack* hb, const float* conv_map, int B, int cap, const int32_t* offs,
                        const int32_t* pad, float* prob, float* reg, float* lmk, int crop_cap, void* stream);

extern "C" int vnfr_heads_debug(long long* dev_buf) {      // debug hook (not part of include/vnfr_b200.h)
  VNFR_CUDA(cudaMemcpyToSymbol(g_heads_dbg, &dev_buf, sizeof(dev_buf)));
  return VNFR_OK;
}

extern "C" int vnfr_rnet_weight_floats(void) { return RW::END; }
extern "C" int vnfr_onet_weight_floats(void) { return OW_::END; }

static int run_head(bool onet, const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                    const float* weights, float* prob, float* reg, float* lmk, int32_t* offs, float* crops, int crop_cap,
                    int32_t* status, void* stream) {
  VNFR_REQUIRE(frames && count && pad && weights && prob && reg && offs && crops && status, "null pointer");
  VNFR_REQUIRE(crop_cap > 0 && ((uintptr_t)crops % 16) == 0, "crop workspace must hold at least one crop and be 16-byte aligned");
  VNFR_REQUIRE(!onet || lmk != nullptr, "O-Net needs a landmark buffer");
  if (B == 0) return VNFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  scan_counts_kernel<<<1, 32, 0, st>>>(count, B, cap, offs);
  ++g_vnfr_launches;
  HeadArgs a;
  a.frames = frames; a.B = B; a.H = H; a.W = W; a.cap = cap; a.count = count;
  a.pad = reinterpret_cast<const int4*>(pad); a.offs = offs; a.w = weights; a.prob = prob;
  a.reg = reinterpret_cast<float4*>(reg); a.lmk = lmk; a.crops = crops; a.crop_cap = crop_cap; a.status = status;
  a.p1 = nullptr; a.c2 = nullptr; a.split_mode = 0; a.p3 = nullptr; a.c3 = nullptr;
  static VnfrPerDevice attr_once = {};
  if (vnfr_first_on_device(attr_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(rnet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(onet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, O_SMEM));
  }
  if (onet) launch_crops<48>(a, st);
  else launch_crops<24>(a, st);
  ++g_vnfr_launches;
  // persistent grid: one CTA per SM (shared memory bound), each loops over the flat candidate list
  if (onet) onet_kernel<<<148 * 1, NT, O_SMEM, st>>>(a);
  else rnet_kernel<<<148 * 1, NT, R_SMEM, st>>>(a);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

// O-Net with conv2 on the tensor cores (see onet_front_kernel).  split_mode 1: w2_split bf16 [64][1728]
// (encoder_plan.pack_conv_split3), p1 101 568 B per crop; split_mode 2: w2_split fp16 [64][896] (pack_conv_split2), p1
// 67 712 B per crop; c2: 112 896 B per crop.
extern "C" int vnfr_onet_forward_tc(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                                    const float* weights, const void* w2_split, int split_mode, float* prob, float* reg, float* lmk,
                                    int32_t* offs, float* crops, void* p1, float* c2, const void* w3_split, void* p3, float* c3,
                                    int crop_cap, int32_t* status, const VnfrHeadsBack* back, void* stream) {
  VNFR_REQUIRE(frames && count && pad && weights && w2_split && prob && reg && lmk && offs && crops && p1 && c2 && status, "null pointer");
  VNFR_REQUIRE(back == nullptr || w3_split != nullptr, "the tensor-core back half of O-Net needs conv3 on the tensor cores (w3_split)");
  VNFR_REQUIRE(split_mode == 1 || split_mode == 2, "split_mode must be 1 (3 x bf16) or 2 (2 x fp16)");
  const bool tc3 = w3_split != nullptr;
  VNFR_REQUIRE(!tc3 || (p3 != nullptr && c3 != nullptr && ((uintptr_t)p3 % 16) == 0 && ((uintptr_t)c3 % 16) == 0),
               "conv3 on the tensor cores needs the p3 / c3 workspaces (16-byte aligned)");
  VNFR_REQUIRE(crop_cap > 0 && ((uintptr_t)crops % 16) == 0 && ((uintptr_t)p1 % 16) == 0 && ((uintptr_t)c2 % 16) == 0,
               "workspaces must hold at least one crop and be 16-byte aligned");
  if (B == 0) return VNFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  scan_counts_kernel<<<1, 32, 0, st>>>(count, B, cap, offs);
  ++g_vnfr_launches;
  HeadArgs a;
  a.frames = frames; a.B = B; a.H = H; a.W = W; a.cap = cap; a.count = count;
  a.pad = reinterpret_cast<const int4*>(pad); a.offs = offs; a.w = weights; a.prob = prob;
  a.reg = reinterpret_cast<float4*>(reg); a.lmk = lmk; a.crops = crops; a.crop_cap = crop_cap; a.status = status;
  a.p1 = p1; a.c2 = c2; a.split_mode = split_mode; a.p3 = tc3 ? p3 : nullptr; a.c3 = tc3 ? c3 : nullptr;
  static VnfrPerDevice attr_once = {};
  if (vnfr_first_on_device(attr_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(onet_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OF_SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(head_front_tc_kernel<ONET_FRONT_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   FrontTc<ONET_FRONT_TC>::SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(onet_back_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, OB_SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(onet_back_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OB_SMEM));
  }
  launch_crops<48>(a, st);
  static const bool front_fma = getenv("VNFR_ONET_FRONT_FMA") != nullptr;
  // (the tensor-core front writes the two-part fp16 split only: split_mode 1 keeps the FMA front)
  if (front_fma || split_mode != 2) onet_front_kernel<<<148, NT, OF_SMEM, st>>>(a);
  else head_front_tc_kernel<ONET_FRONT_TC><<<148 * 2, RT_THREADS, FrontTc<ONET_FRONT_TC>::SMEM, st>>>(a);
  g_vnfr_launches += 2;
  VNFR_CHECK_LAUNCH();
  // conv2 on the tensor cores; the tensor maps are re-encoded only when a pointer or the capacity changes
  static VnfrConvOp op;
  static const void* key[4] = {nullptr, nullptr, nullptr, nullptr};
  static int key_cap = -1, key_mode = 0;
  if (key[0] != p1 || key[1] != w2_split || key[2] != (const void*)c2 || key[3] != (const void*)weights || key_cap != crop_cap ||
      key_mode != split_mode) {
    memset(&op, 0, sizeof(op));
    op.in = p1; op.weights = w2_split; op.bias = weights + OW_::B2; op.prelu_alpha = weights + OW_::A2;
    op.out_f32 = c2; op.out_f32_pitch = 64;
    const int parts = split_mode == 2 ? 2 : 3;
    op.n_img = crop_cap; op.in_h = 23; op.in_w = 23; op.cin = 32 * parts; op.in_pitch = 32 * parts;
    op.kh = 3; op.kw = 3; op.stride = 1; op.pad_h = 0; op.pad_w = 0; op.out_h = 21; op.out_w = 21;
    op.cout = 64; op.cout_pad = 64; op.k_pad = split_mode == 2 ? 896 : 1728; op.block_n = 64; op.n_split = 64;
    op.relu = 0; op.dtype = split_mode == 2 ? 1 : 0; op.reserved[0] = 32; op.split3 = split_mode;
    op.n_img_dev = offs + B;                     // total candidate count, written by scan_counts_kernel
    const int rc = vnfr_conv_prepare(&op);
    if (rc != VNFR_OK) return rc;
    VNFR_REQUIRE(op.a_mode == 3, "split-precision conv2 did not qualify for the shifted-view kernel");
    key[0] = p1; key[1] = w2_split; key[2] = c2; key[3] = weights; key_cap = crop_cap; key_mode = split_mode;
  }
  op.n_img_dev = offs + B;
  {
    const int rc = vnfr_conv_run(&op, stream);
    if (rc != VNFR_OK) return rc;
  }
  if (tc3) {
    // conv3 (64 -> 64, 3x3 on the pooled 10x10 map) on the tensor cores, two fp16 parts / three products
    onet_mid_kernel<<<148 * 8, 256, 0, st>>>(a);
    ++g_vnfr_launches;
    VNFR_CHECK_LAUNCH();
    static VnfrConvOp op3;
    static const void* key3[4] = {nullptr, nullptr, nullptr, nullptr};
    static int key3_cap = -1;
    if (key3[0] != p3 || key3[1] != w3_split || key3[2] != (const void*)c3 || key3[3] != (const void*)weights || key3_cap != crop_cap) {
      memset(&op3, 0, sizeof(op3));
      op3.in = p3; op3.weights = w3_split; op3.bias = weights + OW_::B3; op3.prelu_alpha = weights + OW_::A3;
      op3.out_f32 = c3; op3.out_f32_pitch = 64;
      op3.n_img = crop_cap; op3.in_h = 10; op3.in_w = 10; op3.cin = 128; op3.in_pitch = 128;
      op3.kh = 3; op3.kw = 3; op3.stride = 1; op3.pad_h = 0; op3.pad_w = 0; op3.out_h = 8; op3.out_w = 8;
      op3.cout = 64; op3.cout_pad = 64; op3.k_pad = 1728; op3.block_n = 64; op3.n_split = 64;
      op3.relu = 0; op3.dtype = 1; op3.reserved[0] = 64; op3.split3 = 2;
      op3.n_img_dev = offs + B;
      const int rc = vnfr_conv_prepare(&op3);
      if (rc != VNFR_OK) return rc;
      VNFR_REQUIRE(op3.a_mode == 3, "split-precision conv3 did not qualify for the shifted-view kernel");
      key3[0] = p3; key3[1] = w3_split; key3[2] = c3; key3[3] = weights; key3_cap = crop_cap;
    }
    op3.n_img_dev = offs + B;
    const int rc = vnfr_conv_run(&op3, stream);
    if (rc != VNFR_OK) return rc;
    if (back != nullptr) return vnfr_heads_back_run(1, back, c3, B, cap, offs, pad, prob, reg, lmk, crop_cap, stream);
    onet_back_kernel<true><<<148, NT, OB_SMEM, st>>>(a);
  } else {
    onet_back_kernel<false><<<148, NT, OB_SMEM, st>>>(a);
  }
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

// R-Net with conv2 on the tensor cores (see rnet_front_kernel).  w2_split: fp16 [48][896] (encoder_plan.pack_conv_split2, sv_ck
// 32); p1: fp16 [crop_cap][11][11][64], c2: fp32 [crop_cap][81][48] workspaces (15 488 B and 15 552 B per crop).
extern "C" int vnfr_rnet_forward_tc(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                                    const float* weights, const void* w2_split, float* prob, float* reg, int32_t* offs,
                                    float* crops, void* p1, float* c2, int crop_cap, int32_t* status, const VnfrHeadsBack* back,
                                    void* stream) {
  VNFR_REQUIRE(frames && count && pad && weights && w2_split && prob && reg && offs && crops && p1 && c2 && status, "null pointer");
  VNFR_REQUIRE(crop_cap > 0 && ((uintptr_t)crops % 16) == 0 && ((uintptr_t)p1 % 16) == 0 && ((uintptr_t)c2 % 16) == 0,
               "workspaces must hold at least one crop and be 16-byte aligned");
  if (B == 0) return VNFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  scan_counts_kernel<<<1, 32, 0, st>>>(count, B, cap, offs);
  ++g_vnfr_launches;
  HeadArgs a;
  a.frames = frames; a.B = B; a.H = H; a.W = W; a.cap = cap; a.count = count;
  a.pad = reinterpret_cast<const int4*>(pad); a.offs = offs; a.w = weights; a.prob = prob;
  a.reg = reinterpret_cast<float4*>(reg); a.lmk = nullptr; a.crops = crops; a.crop_cap = crop_cap; a.status = status;
  a.p1 = p1; a.c2 = c2; a.split_mode = 2; a.p3 = nullptr; a.c3 = nullptr;
  static VnfrPerDevice attr_once = {};
  if (vnfr_first_on_device(attr_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(rnet_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(head_front_tc_kernel<RNET_FRONT_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   FrontTc<RNET_FRONT_TC>::SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(rnet_back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RB_SMEM));
  }
  launch_crops<24>(a, st);
  static const bool front_fma = getenv("VNFR_RNET_FRONT_FMA") != nullptr;
  if (front_fma) rnet_front_kernel<<<148, NT, R_SMEM, st>>>(a);
  else head_front_tc_kernel<RNET_FRONT_TC><<<148 * 2, RT_THREADS, FrontTc<RNET_FRONT_TC>::SMEM, st>>>(a);
  g_vnfr_launches += 2;
  VNFR_CHECK_LAUNCH();
  static VnfrConvOp op;
  static const void* key[4] = {nullptr, nullptr, nullptr, nullptr};
  static int key_cap = -1;
  if (key[0] != p1 || key[1] != w2_split || key[2] != (const void*)c2 || key[3] != (const void*)weights || key_cap != crop_cap) {
    memset(&op, 0, sizeof(op));
    op.in = p1; op.weights = w2_split; op.bias = weights + RW::B2; op.prelu_alpha = weights + RW::A2;
    op.out_f32 = c2; op.out_f32_pitch = 48;
    op.n_img = crop_cap; op.in_h = 11; op.in_w = 11; op.cin = 64; op.in_pitch = 64;
    op.kh = 3; op.kw = 3; op.stride = 1; op.pad_h = 0; op.pad_w = 0; op.out_h = 9; op.out_w = 9;
    op.cout = 48; op.cout_pad = 48; op.k_pad = 896; op.block_n = 48; op.n_split = 48;
    op.relu = 0; op.dtype = 1; op.reserved[0] = 32; op.split3 = 2;
    op.n_img_dev = offs + B;                     // total candidate count, written by scan_counts_kernel
    const int rc = vnfr_conv_prepare(&op);
    if (rc != VNFR_OK) return rc;
    VNFR_REQUIRE(op.a_mode == 3, "split-precision conv2 did not qualify for the shifted-view kernel");
    key[0] = p1; key[1] = w2_split; key[2] = c2; key[3] = weights; key_cap = crop_cap;
  }
  op.n_img_dev = offs + B;
  {
    const int rc = vnfr_conv_run(&op, stream);
    if (rc != VNFR_OK) return rc;
  }
  if (back != nullptr) return vnfr_heads_back_run(0, back, c2, B, cap, offs, pad, prob, reg, nullptr, crop_cap, stream);
  rnet_back_kernel<<<148 * 2, NT, RB_SMEM, st>>>(a);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_rnet_forward(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                                 const float* weights, float* prob, float* reg, int32_t* offs, float* crops, int crop_cap,
                                 int32_t* status, void* stream) {
  return run_head(false, frames, B, H, W, cap, count, pad, weights, prob, reg, nullptr, offs, crops, crop_cap, status, stream);
}

extern "C" int vnfr_onet_forward(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                                 const float* weights, float* prob, float* reg, float* lmk, int32_t* offs, float* crops,
                                 int crop_cap, int32_t* status, void* stream) {
  return run_head(true, frames, B, H, W, cap, count, pad, weights, prob, reg, lmk, offs, crops, crop_cap, status, stream);
}
