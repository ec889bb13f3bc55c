// P-Net as one fully-convolutional sweep over EVERY pyramid level of EVERY frame in a single launch, with softmax,
// the `>= threshold` test of generateBoundingBox and the candidate compaction fused into the epilogue.
//
// Replaces mtcnn.py:38-49 (conv3x3(3->10)+PReLU, maxpool2 ceil, conv3x3(10->16)+PReLU, conv3x3(16->32)+PReLU,
// {1x1->2 softmax, 1x1->4}) run once per scale at detect_face.py:73, and detect_face.py:203-218 (mask, nonzero,
// gathers).  The dense prob/reg maps are never written (only on request, for parity tests).
//
// Persistent CTAs (3 per SM) walk over 16x16 tiles of output cells: the 42x42x3 input patch, the pooled conv1 map
// (20x20x10) and the conv2 map (18x18x16) live in shared memory; conv3 + both heads are held in registers (two cells
// x 32 channels per thread).  All 6 632 weights are copied once per CTA into shared memory in K-major order with the
// output channel innermost, so one 16-byte broadcast load feeds 4 FMAs per cell (the first version kept them in
// __constant__ memory: 26 KB of indexed constant loads thrashed the constant cache -- 12.7 % issue utilisation, see
// profiles/).  Arithmetic is fp32 FMA throughout: thresholded decisions must match the fp32 reference.
#include "common.cuh"
#include <math_constants.h>
#include <string.h>

extern long long g_vnfr_launches;

namespace {

constexpr int T = 16;                 // output tile edge
constexpr int PT = T + 4;             // pooled conv1 tile edge (20)
constexpr int C2T = T + 2;            // conv2 tile edge (18)
constexpr int IT = 2 * PT + 2;        // input tile edge (42)
constexpr int ITP = IT + 1;           // padded row pitch
#ifndef PNET_MIN_CTAS
#define PNET_MIN_CTAS 2
#endif
constexpr int NTHR = 128;             // threads per CTA: conv3 phase = 2 cells x 32 channels per thread

// host-packed weights (floats), torch layouts [co][ci][ky][kx] -- what vnfr_pnet_set_weights receives
constexpr int W1 = 0, B1 = W1 + 270, A1 = B1 + 10;
constexpr int W2 = A1 + 10, B2 = W2 + 1440, A2 = B2 + 16;
constexpr int W3 = A2 + 16, B3 = W3 + 4608, A3 = B3 + 32;
constexpr int W41 = A3 + 32, B41 = W41 + 64, W42 = B41 + 2, B42 = W42 + 128;
constexpr int PNET_FLOATS = B42 + 4;  // 6632

// device layout: K-major with the output channel innermost (16-byte vector loads broadcast to the whole warp)
//   conv1 [27][12] (10 used), conv2 [90][16], conv3 [144][32], heads [32][8] (0-1 logits, 2-5 reg), then bias / PReLU rows
constexpr int D_W1 = 0, D_B1 = D_W1 + 27 * 12, D_A1 = D_B1 + 12;
constexpr int D_W2 = D_A1 + 12, D_B2 = D_W2 + 90 * 16, D_A2 = D_B2 + 16;
constexpr int D_W3 = D_A2 + 16, D_B3 = D_W3 + 144 * 32, D_A3 = D_B3 + 32;
constexpr int D_W4 = D_A3 + 32, D_B4 = D_W4 + 32 * 8;
constexpr int D_FLOATS = D_B4 + 8;    // 6796 floats (all offsets are multiples of 4)

__device__ float g_pnet_w[D_FLOATS];  // repacked weights (global, L2-resident; copied into shared memory per CTA)

constexpr int S_BUF = 3 * IT * ITP > 16 * C2T * C2T ? 3 * IT * ITP : 16 * C2T * C2T;   // input patch, later conv2 map
constexpr int S_P = 10 * PT * PT;
constexpr int PNET_SMEM = (D_FLOATS + S_BUF + S_P) * 4;

struct PnetParams {
  int B, n_levels;
  int lh[VNFR_MAX_LEVELS], lw[VNFR_MAX_LEVELS], oh[VNFR_MAX_LEVELS], ow[VNFR_MAX_LEVELS];
  int tiles_x[VNFR_MAX_LEVELS], tile_off[VNFR_MAX_LEVELS + 1];
  long long level_off[VNFR_MAX_LEVELS], map_off[VNFR_MAX_LEVELS];
};

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

__global__ void __launch_bounds__(NTHR, PNET_MIN_CTAS) pnet_kernel(const __grid_constant__ PnetParams p, const float* __restrict__ levels,
                                                    float thr, int cap, int* __restrict__ cand_count,
                                                    uint32_t* __restrict__ cand_cell, float* __restrict__ cand_score,
                                                    float4* __restrict__ cand_reg, float* __restrict__ dense_prob,
                                                    float* __restrict__ dense_reg) {
  extern __shared__ __align__(16) float smem[];
  float* s_w = smem;                                  // weights
  float* s_buf = smem + D_FLOATS;                     // input patch [3][IT][ITP], then conv2 map [16][C2T][C2T]
  float* s_p = s_buf + S_BUF;                         // pooled conv1 map [10][PT][PT], then head partials
  const int tid = threadIdx.x;
  for (int i = tid; i < D_FLOATS / 4; i += NTHR) reinterpret_cast<float4*>(s_w)[i] = reinterpret_cast<const float4*>(g_pnet_w)[i];

  const int tiles_per_img = p.tile_off[p.n_levels];
  const int total_tiles = tiles_per_img * p.B;
  // persistent CTAs; tiles are numbered frame-major so that concurrently running CTAs work on the same frame's levels
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int t = tile - b * tiles_per_img;
    int l = 0;
    while (l + 1 < p.n_levels && t >= p.tile_off[l + 1]) ++l;
    const int lt = t - p.tile_off[l];
    const int ty0 = (lt / p.tiles_x[l]) * T, tx0 = (lt % p.tiles_x[l]) * T;
    const int lh = p.lh[l], lw = p.lw[l], oh = p.oh[l], ow = p.ow[l];
    __syncthreads();        // previous tile finished with the shared buffers (and the weights are loaded)

    // ---- input patch (zero outside the level)
    {
      const float* src = levels + p.level_off[l] + (size_t)b * 3 * lh * lw;
      const int gy0 = 2 * ty0, gx0 = 2 * tx0;
      // 8 independent loads in flight per thread (a load -> store chain per element left the L2 latency exposed: the
      // ncu source view charged 20 % of all stall samples to these stores waiting on their loads)
      constexpr int NEL = 3 * IT * IT, UNR = 8;
      for (int i0 = tid; i0 < NEL; i0 += NTHR * UNR) {
        float v[UNR];
        int dsto[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int i = i0 + u * NTHR;
          const int ii = i < NEL ? i : 0;
          const int c = ii / (IT * IT), r = ii - c * (IT * IT);
          const int y = r / IT, x = r - y * IT;
          const int gy = gy0 + y, gx = gx0 + x;
          dsto[u] = i < NEL ? (c * IT + y) * ITP + x : -1;
          v[u] = (i < NEL && gy < lh && gx < lw) ? __ldg(src + ((size_t)c * lh + gy) * lw + gx) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (dsto[u] >= 0) s_buf[dsto[u]] = v[u];
      }
    }
    __syncthreads();

    // ---- conv1 (3->10, 3x3) + PReLU + maxpool 2x2/2 ceil_mode: one pooled position (4 conv positions x 10 ch) per item
    {
      const int c1h = lh - 2, c1w = lw - 2;        // valid conv1 extent of this level
      for (int pos = tid; pos < PT * PT; pos += NTHR) {
        const int py = pos / PT, px = pos - py * PT;
        float patch[3][4][4];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int y = 0; y < 4; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) patch[c][y][x] = s_buf[(c * IT + 2 * py + y) * ITP + 2 * px + x];
        float acc[4][10];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int co = 0; co < 10; ++co) acc[q][co] = s_w[D_B1 + co];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float4* wr = reinterpret_cast<const float4*>(s_w + D_W1 + ((ci * 3 + ky) * 3 + kx) * 12);
              const float4 wa = wr[0], wb = wr[1];
              const float2 wc = *reinterpret_cast<const float2*>(wr + 2);
              const float w[10] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y};
#pragma unroll
              for (int co = 0; co < 10; ++co) {
                acc[0][co] = fmaf(w[co], patch[ci][ky][kx], acc[0][co]);
                acc[1][co] = fmaf(w[co], patch[ci][ky][kx + 1], acc[1][co]);
                acc[2][co] = fmaf(w[co], patch[ci][ky + 1][kx], acc[2][co]);
                acc[3][co] = fmaf(w[co], patch[ci][ky + 1][kx + 1], acc[3][co]);
              }
            }
        const int cy = 2 * (ty0 + py), cx = 2 * (tx0 + px);
        const bool vy1 = cy + 1 < c1h, vx1 = cx + 1 < c1w, v00 = cy < c1h && cx < c1w;
#pragma unroll
        for (int co = 0; co < 10; ++co) {
          const float al = s_w[D_A1 + co];
          float m = v00 ? prelu(acc[0][co], al) : 0.f;     // windows clipped at the border (ceil_mode) use valid cells only
          if (v00 && vx1) m = fmaxf(m, prelu(acc[1][co], al));
          if (v00 && vy1) m = fmaxf(m, prelu(acc[2][co], al));
          if (v00 && vy1 && vx1) m = fmaxf(m, prelu(acc[3][co], al));
          s_p[(co * PT + py) * PT + px] = m;
        }
      }
    }
    __syncthreads();

    // ---- conv2 (10->16, 3x3) + PReLU: item = 3 positions (p, p+108, p+216) x 16 channels
    if (tid < 108) {
      int off[3], oo[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int pos = tid + 108 * j;
        const int y = pos / C2T, x = pos - y * C2T;
        off[j] = y * PT + x;
        oo[j] = y * C2T + x;
      }
      float acc[3][16];
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int co = 0; co < 16; ++co) acc[j][co] = s_w[D_B2 + co];
#pragma unroll 2
      for (int ci = 0; ci < 10; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4* wr = reinterpret_cast<const float4*>(s_w + D_W2 + ((ci * 3 + ky) * 3 + kx) * 16);
            const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
            const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
            const int o = (ci * PT + ky) * PT + kx;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float v = s_p[off[j] + o];
#pragma unroll
              for (int co = 0; co < 16; ++co) acc[j][co] = fmaf(w[co], v, acc[j][co]);
            }
          }
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int co = 0; co < 16; ++co) s_buf[co * C2T * C2T + oo[j]] = prelu(acc[j][co], s_w[D_A2 + co]);
    }
    __syncthreads();

    // ---- conv3 (16->32, 3x3) + PReLU + heads: thread = cells (y, x) and (y + 8, x), 32 channels each in registers
    const int y = tid >> 4, x = tid & 15;
    // packed fp32 pairs (FFMA2, fma.rn.f32x2): accp[h][q] = channels (2q, 2q+1) of cell h -- half the FMA issue slots,
    // bit-identical results
    unsigned long long accp[2][16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const unsigned long long b2 = *reinterpret_cast<const unsigned long long*>(s_w + D_B3 + 2 * q);
      accp[0][q] = b2; accp[1][q] = b2;
    }
#pragma unroll 1
    for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v0 = s_buf[(ci * C2T + y + ky) * C2T + x + kx];
          const float v1 = s_buf[(ci * C2T + y + 8 + ky) * C2T + x + kx];
          const unsigned long long vv0 = (unsigned long long)__float_as_uint(v0) | ((unsigned long long)__float_as_uint(v0) << 32);
          const unsigned long long vv1 = (unsigned long long)__float_as_uint(v1) | ((unsigned long long)__float_as_uint(v1) << 32);
          const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(s_w + D_W3 + ((ci * 3 + ky) * 3 + kx) * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const ulonglong2 w4 = wr[q];
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accp[0][2 * q]) : "l"(w4.x), "l"(vv0));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accp[1][2 * q]) : "l"(w4.x), "l"(vv1));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accp[0][2 * q + 1]) : "l"(w4.y), "l"(vv0));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accp[1][2 * q + 1]) : "l"(w4.y), "l"(vv1));
          }
        }
    }
    float acc[2][32];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        acc[h][2 * q] = __uint_as_float((unsigned)(accp[h][q] & 0xFFFFFFFFull));
        acc[h][2 * q + 1] = __uint_as_float((unsigned)(accp[h][q] >> 32));
      }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = s_w[D_B4 + j];
#pragma unroll
      for (int co = 0; co < 32; ++co) {
        const float v = prelu(acc[h][co], s_w[D_A3 + co]);
        const float4 wa = *reinterpret_cast<const float4*>(s_w + D_W4 + co * 8);
        const float4 wb = *reinterpret_cast<const float4*>(s_w + D_W4 + co * 8 + 4);
        o[0] = fmaf(wa.x, v, o[0]); o[1] = fmaf(wa.y, v, o[1]); o[2] = fmaf(wa.z, v, o[2]); o[3] = fmaf(wa.w, v, o[3]);
        o[4] = fmaf(wb.x, v, o[4]); o[5] = fmaf(wb.y, v, o[5]);
      }
      // softmax over the two logits (torch: exp(x - max) / sum)
      const float mx = fmaxf(o[0], o[1]);
      const float e0 = expf(o[0] - mx), e1 = expf(o[1] - mx);
      const float prob = e1 / (e0 + e1);
      const int gy = ty0 + y + 8 * h, gx = tx0 + x;
      const bool valid = gy < oh && gx < ow;
      if (valid && dense_prob != nullptr) {
        const size_t cells = (size_t)oh * ow;
        const size_t cell = (size_t)gy * ow + gx;
        dense_prob[p.map_off[l] + (size_t)b * cells + cell] = prob;
        if (dense_reg != nullptr) {
          float* dr = dense_reg + 4 * (p.map_off[l] + (size_t)b * cells);
          dr[cell] = o[2]; dr[cells + cell] = o[3]; dr[2 * cells + cell] = o[4]; dr[3 * cells + cell] = o[5];
        }
      }
      // generateBoundingBox: probs >= thresh (detect_face.py:209) -> warp-ballot stream compaction into the segment
      const bool pass = valid && prob >= thr;
      const unsigned ballot = __ballot_sync(0xffffffffu, pass);
      if (ballot != 0u) {
        const int seg = b * p.n_levels + l;
        const int lane = tid & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(cand_count + seg, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) {
          const int slot = base + __popc(ballot & ((1u << lane) - 1u));
          if (slot < cap) {
            const size_t oidx = (size_t)seg * cap + slot;
            cand_cell[oidx] = ((uint32_t)gy << 16) | (uint32_t)gx;
            cand_score[oidx] = prob;
            cand_reg[oidx] = make_float4(o[2], o[3], o[4], o[5]);
          }
        }
      }
    }
  }
}

}  // namespace

extern "C" int vnfr_pnet_set_weights(const float* packed_host, int n_floats, void* stream) {
  VNFR_REQUIRE(packed_host != nullptr && n_floats == PNET_FLOATS, "P-Net packed weights must hold 6632 floats");
  static float d[D_FLOATS];          // repack [co][k] -> [k][co] (padded); static: the async copy reads it after return
  memset(d, 0, sizeof(d));
  const float* h = packed_host;
  for (int co = 0; co < 10; ++co) { for (int k = 0; k < 27; ++k) d[D_W1 + k * 12 + co] = h[W1 + co * 27 + k]; d[D_B1 + co] = h[B1 + co]; d[D_A1 + co] = h[A1 + co]; }
  for (int co = 0; co < 16; ++co) { for (int k = 0; k < 90; ++k) d[D_W2 + k * 16 + co] = h[W2 + co * 90 + k]; d[D_B2 + co] = h[B2 + co]; d[D_A2 + co] = h[A2 + co]; }
  for (int co = 0; co < 32; ++co) { for (int k = 0; k < 144; ++k) d[D_W3 + k * 32 + co] = h[W3 + co * 144 + k]; d[D_B3 + co] = h[B3 + co]; d[D_A3 + co] = h[A3 + co]; }
  for (int c = 0; c < 32; ++c) {
    d[D_W4 + c * 8 + 0] = h[W41 + c]; d[D_W4 + c * 8 + 1] = h[W41 + 32 + c];
    for (int j = 0; j < 4; ++j) d[D_W4 + c * 8 + 2 + j] = h[W42 + j * 32 + c];
  }
  d[D_B4 + 0] = h[B41]; d[D_B4 + 1] = h[B41 + 1];
  for (int j = 0; j < 4; ++j) d[D_B4 + 2 + j] = h[B42 + j];
  VNFR_CUDA(cudaMemcpyToSymbolAsync(g_pnet_w, d, sizeof(d), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return VNFR_OK;
}

extern "C" int vnfr_pnet_sweep_compact(const VnfrPyramid* pyr, const float* levels, float threshold, int cap,
                                       int32_t* cand_count, uint32_t* cand_cell, float* cand_score, float* cand_reg,
                                       float* dense_prob, float* dense_reg, void* stream) {
  VNFR_REQUIRE(pyr != nullptr && cap > 0, "bad arguments");
  const int tiles = pyr->tile_off[pyr->n_levels];
  if (pyr->B == 0 || tiles == 0) return VNFR_OK;
  PnetParams p;
  p.B = pyr->B; p.n_levels = pyr->n_levels;
  for (int l = 0; l < pyr->n_levels; ++l) {
    p.lh[l] = pyr->lh[l]; p.lw[l] = pyr->lw[l]; p.oh[l] = pyr->oh[l]; p.ow[l] = pyr->ow[l];
    p.tiles_x[l] = pyr->tiles_x[l]; p.tile_off[l] = pyr->tile_off[l];
    p.level_off[l] = pyr->level_off[l]; p.map_off[l] = pyr->map_off[l];
  }
  p.tile_off[pyr->n_levels] = tiles;
  static bool attr = false;
  if (!attr) {
    VNFR_CUDA(cudaFuncSetAttribute(pnet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PNET_SMEM));
    attr = true;
  }
  const int total = pyr->B * tiles;
  const int grid = total < 148 * PNET_MIN_CTAS ? total : 148 * PNET_MIN_CTAS;      // persistent CTAs
  pnet_kernel<<<grid, NTHR, PNET_SMEM, (cudaStream_t)stream>>>(p, levels, threshold, cap, cand_count, cand_cell, cand_score,
                                                              reinterpret_cast<float4*>(cand_reg), dense_prob, dense_reg);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
