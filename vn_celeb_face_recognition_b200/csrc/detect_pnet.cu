// P-Net as one fully-convolutional sweep over EVERY pyramid level of EVERY frame in a single launch, with softmax,
// the `>= threshold` test of generateBoundingBox and the candidate compaction fused into the epilogue.
//
// Replaces mtcnn.py:38-49 (conv3x3(3->10)+PReLU, maxpool2 ceil, conv3x3(10->16)+PReLU, conv3x3(16->32)+PReLU,
// {1x1->2 softmax, 1x1->4}) run once per scale at detect_face.py:73, and detect_face.py:203-218 (mask, nonzero,
// gathers).  The dense prob/reg maps are never written (only on request, for parity tests).
//
// One CTA computes a 16x16 tile of output cells: the 42x42x3 input patch, the pooled conv1 map (20x20x10) and the
// conv2 map (18x18x16) live in shared memory; conv3 + both heads are held in registers (one cell per thread).  All
// 6 632 weights sit in __constant__ memory and every inner loop is fully unrolled so that each FMA takes its weight
// as a constant-bank operand (no load instruction); arithmetic is fp32 FMA throughout (thresholded decisions must
// match the fp32 reference).
#include "common.cuh"
#include <math_constants.h>

extern long long g_vnfr_launches;

namespace {

constexpr int T = 16;                 // output tile edge
constexpr int PT = T + 4;             // pooled conv1 tile edge (20)
constexpr int C2T = T + 2;            // conv2 tile edge (18)
constexpr int IT = 2 * PT + 2;        // input tile edge (42)
constexpr int ITP = IT + 1;           // padded row pitch

// packed weight offsets (floats), torch layouts [co][ci][ky][kx]
constexpr int W1 = 0, B1 = W1 + 270, A1 = B1 + 10;
constexpr int W2 = A1 + 10, B2 = W2 + 1440, A2 = B2 + 16;
constexpr int W3 = A2 + 16, B3 = W3 + 4608, A3 = B3 + 32;
constexpr int W41 = A3 + 32, B41 = W41 + 64, W42 = B41 + 2, B42 = W42 + 128;
constexpr int PNET_FLOATS = B42 + 4;  // 6632

__constant__ float c_w[PNET_FLOATS];

struct PnetParams {
  int B, n_levels;
  int lh[VNFR_MAX_LEVELS], lw[VNFR_MAX_LEVELS], oh[VNFR_MAX_LEVELS], ow[VNFR_MAX_LEVELS];
  int tiles_x[VNFR_MAX_LEVELS], tile_off[VNFR_MAX_LEVELS + 1];
  long long level_off[VNFR_MAX_LEVELS], map_off[VNFR_MAX_LEVELS];
};

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

__global__ void __launch_bounds__(256) pnet_kernel(const __grid_constant__ PnetParams p, const float* __restrict__ levels,
                                                   float thr, int cap, int* __restrict__ cand_count,
                                                   uint32_t* __restrict__ cand_cell, float* __restrict__ cand_score,
                                                   float4* __restrict__ cand_reg, float* __restrict__ dense_prob,
                                                   float* __restrict__ dense_reg) {
  // the input patch is dead once conv1 is done, so the conv2 map reuses its storage (37.7 KB static in total)
  __shared__ __align__(16) float s_buf[3 * IT * ITP > 16 * C2T * C2T ? 3 * IT * ITP : 16 * C2T * C2T];
  __shared__ float s_p[10][PT][PT];
  float (*s_in)[IT][ITP] = reinterpret_cast<float (*)[IT][ITP]>(s_buf);
  float (*s_c2)[C2T][C2T] = reinterpret_cast<float (*)[C2T][C2T]>(s_buf);

  const int tiles_per_img = p.tile_off[p.n_levels];
  const int b = blockIdx.x / tiles_per_img;
  const int t = blockIdx.x - b * tiles_per_img;
  int l = 0;
  while (l + 1 < p.n_levels && t >= p.tile_off[l + 1]) ++l;
  const int lt = t - p.tile_off[l];
  const int ty0 = (lt / p.tiles_x[l]) * T, tx0 = (lt % p.tiles_x[l]) * T;
  const int lh = p.lh[l], lw = p.lw[l], oh = p.oh[l], ow = p.ow[l];
  const int tid = threadIdx.x;

  // ---- input patch (zero outside the level)
  {
    const float* src = levels + p.level_off[l] + (size_t)b * 3 * lh * lw;
    const int gy0 = 2 * ty0, gx0 = 2 * tx0;
    for (int i = tid; i < 3 * IT * IT; i += 256) {
      const int c = i / (IT * IT), r = i - c * (IT * IT);
      const int y = r / IT, x = r - y * IT;
      const int gy = gy0 + y, gx = gx0 + x;
      s_in[c][y][x] = (gy < lh && gx < lw) ? __ldg(src + ((size_t)c * lh + gy) * lw + gx) : 0.f;
    }
  }
  __syncthreads();

  // ---- conv1 (3->10, 3x3) + PReLU + maxpool 2x2 stride 2 ceil_mode: one pooled position per thread-iteration
  {
    const int c1h = lh - 2, c1w = lw - 2;        // valid conv1 extent of this level
    for (int pos = tid; pos < PT * PT; pos += 256) {
      const int py = pos / PT, px = pos - py * PT;
      float patch[3][4][4];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int y = 0; y < 4; ++y)
#pragma unroll
          for (int x = 0; x < 4; ++x) patch[c][y][x] = s_in[c][2 * py + y][2 * px + x];
      const int cy = 2 * (ty0 + py), cx = 2 * (tx0 + px);
      const bool vy1 = cy + 1 < c1h, vx1 = cx + 1 < c1w, v00 = cy < c1h && cx < c1w;
#pragma unroll
      for (int co = 0; co < 10; ++co) {
        float a00 = c_w[B1 + co], a01 = a00, a10 = a00, a11 = a00;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float w = c_w[W1 + ((co * 3 + ci) * 3 + ky) * 3 + kx];
              a00 = fmaf(w, patch[ci][ky][kx], a00);
              a01 = fmaf(w, patch[ci][ky][kx + 1], a01);
              a10 = fmaf(w, patch[ci][ky + 1][kx], a10);
              a11 = fmaf(w, patch[ci][ky + 1][kx + 1], a11);
            }
        const float al = c_w[A1 + co];
        float m = v00 ? prelu(a00, al) : 0.f;       // windows clipped at the border (ceil_mode) use valid cells only
        if (v00 && vx1) m = fmaxf(m, prelu(a01, al));
        if (v00 && vy1) m = fmaxf(m, prelu(a10, al));
        if (v00 && vy1 && vx1) m = fmaxf(m, prelu(a11, al));
        s_p[co][py][px] = m;
      }
    }
  }
  __syncthreads();

  // ---- conv2 (10->16, 3x3) + PReLU: work item = (position, half of the output channels)
  for (int item = tid; item < 2 * C2T * C2T; item += 256) {
    const int half = item / (C2T * C2T);
    const int pos = item - half * (C2T * C2T);
    const int y = pos / C2T, x = pos - y * C2T;
    float in[10][3][3];
#pragma unroll
    for (int ci = 0; ci < 10; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) in[ci][ky][kx] = s_p[ci][y + ky][x + kx];
    if (half == 0) {
#pragma unroll
      for (int co = 0; co < 8; ++co) {
        float a = c_w[B2 + co];
#pragma unroll
        for (int ci = 0; ci < 10; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a = fmaf(c_w[W2 + ((co * 10 + ci) * 3 + ky) * 3 + kx], in[ci][ky][kx], a);
        s_c2[co][y][x] = prelu(a, c_w[A2 + co]);
      }
    } else {
#pragma unroll
      for (int co = 8; co < 16; ++co) {
        float a = c_w[B2 + co];
#pragma unroll
        for (int ci = 0; ci < 10; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a = fmaf(c_w[W2 + ((co * 10 + ci) * 3 + ky) * 3 + kx], in[ci][ky][kx], a);
        s_c2[co][y][x] = prelu(a, c_w[A2 + co]);
      }
    }
  }
  __syncthreads();

  // ---- conv3 (16->32, 3x3) + PReLU + heads: one output cell per thread, 32 accumulators in registers
  const int y = tid >> 4, x = tid & 15;
  float acc[32];
#pragma unroll
  for (int co = 0; co < 32; ++co) acc[co] = c_w[B3 + co];
#pragma unroll
  for (int ci = 0; ci < 16; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float v = s_c2[ci][y + ky][x + kx];
#pragma unroll
        for (int co = 0; co < 32; ++co) acc[co] = fmaf(c_w[W3 + ((co * 16 + ci) * 3 + ky) * 3 + kx], v, acc[co]);
      }
  float a0 = c_w[B41], a1 = c_w[B41 + 1];
  float r0 = c_w[B42], r1 = c_w[B42 + 1], r2 = c_w[B42 + 2], r3 = c_w[B42 + 3];
#pragma unroll
  for (int co = 0; co < 32; ++co) {
    const float v = prelu(acc[co], c_w[A3 + co]);
    a0 = fmaf(c_w[W41 + co], v, a0);
    a1 = fmaf(c_w[W41 + 32 + co], v, a1);
    r0 = fmaf(c_w[W42 + co], v, r0);
    r1 = fmaf(c_w[W42 + 32 + co], v, r1);
    r2 = fmaf(c_w[W42 + 64 + co], v, r2);
    r3 = fmaf(c_w[W42 + 96 + co], v, r3);
  }
  // softmax over the two logits (torch: exp(x - max) / sum)
  const float mx = fmaxf(a0, a1);
  const float e0 = expf(a0 - mx), e1 = expf(a1 - mx);
  const float prob = e1 / (e0 + e1);

  const int gy = ty0 + y, gx = tx0 + x;
  const bool valid = gy < oh && gx < ow;
  if (valid && dense_prob != nullptr) {
    const size_t cells = (size_t)oh * ow;
    const size_t cell = (size_t)gy * ow + gx;
    dense_prob[p.map_off[l] + (size_t)b * cells + cell] = prob;
    if (dense_reg != nullptr) {
      float* dr = dense_reg + 4 * (p.map_off[l] + (size_t)b * cells);
      dr[cell] = r0; dr[cells + cell] = r1; dr[2 * cells + cell] = r2; dr[3 * cells + cell] = r3;
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
        const size_t o = (size_t)seg * cap + slot;
        cand_cell[o] = ((uint32_t)gy << 16) | (uint32_t)gx;
        cand_score[o] = prob;
        cand_reg[o] = make_float4(r0, r1, r2, r3);
      }
    }
  }
}

}  // namespace

extern "C" int vnfr_pnet_set_weights(const float* packed_host, int n_floats, void* stream) {
  VNFR_REQUIRE(packed_host != nullptr && n_floats == PNET_FLOATS, "P-Net packed weights must hold 6632 floats");
  VNFR_CUDA(cudaMemcpyToSymbolAsync(c_w, packed_host, sizeof(float) * PNET_FLOATS, 0, cudaMemcpyHostToDevice, (cudaStream_t)stream));
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
  pnet_kernel<<<pyr->B * tiles, 256, 0, (cudaStream_t)stream>>>(p, levels, threshold, cap, cand_count, cand_cell, cand_score,
                                                               reinterpret_cast<float4*>(cand_reg), dense_prob, dense_reg);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
