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
// profiles/).  conv1 / conv2 / heads are fp32 FMA: thresholded decisions must match the fp32 reference.
//
// conv3 (16 -> 32, 3x3: 57 % of the FLOPs) runs on the tensor cores in split precision (default; VNFR_PNET_FMA=1 keeps it
// on the FMA pipe): the conv2 epilogue writes the 18x18x16 map pixel-major as two fp16 parts (hi | lo, 32 bytes = one
// K = 16 slice per pixel, 32-byte swizzle rows) into shared memory; output cell (y, x) is accumulator row r = 18 y + x
// and tap (ky, kx) reads the SAME buffer from a start address shifted by 18 ky + kx rows (the shifted-view trick of
// sv_conv.cu), so one elected thread issues 3 row tiles x 9 taps x 3 products (hi*hi, hi*lo, lo*hi) tcgen05.mma of
// M = 128, N = 32, K = 16 into three 32-column TMEM accumulators; the four warps then read their TMEM lanes back, add
// bias, PReLU and run the heads / softmax / compaction exactly as before.  The two columns x = 16, 17 of every raster
// row and rows r >= 288 are computed and discarded (16/18 useful).
#include "common.cuh"
#include "tc_common.cuh"
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

// The repacked weights live in a CALLER-OWNED device buffer (vnfr_pnet_pack_weights -> upload -> vnfr_pnet_sweep_compact):
// [D_FLOATS fp32 | TC_B_BYTES of fp16 conv3 parts]; every CTA copies them into shared memory.  No process-global state:
// two detectors with different weights can run on two streams.

// tensor-core conv3: B operand = conv3 weights as fp16 hi / lo parts, [tap][part][32 cout rows][16 cin] in the 32-byte
// swizzled K-major layout (16-byte chunk c of row n at n*32 + ((c ^ ((n >> 2) & 1)) << 4)); A operand = the conv2 map,
// one 32-byte row per pixel of the 18-wide raster (432 rows: the last accumulator rows reach row 383 + 38)
constexpr int TC_B_BYTES = 9 * 2 * 1024;
constexpr int TC_A_PART_BYTES = 432 * 32;
constexpr int PNET_PACKED_BYTES = D_FLOATS * 4 + TC_B_BYTES;      // D_FLOATS * 4 is a multiple of 16

constexpr int S_BUF = 3 * IT * ITP > 16 * C2T * C2T ? 3 * IT * ITP : 16 * C2T * C2T;   // input patch, later conv2 map
constexpr int S_BUF_BYTES = (S_BUF * 4 > 2 * TC_A_PART_BYTES ? S_BUF * 4 : 2 * TC_A_PART_BYTES);   // ... or the two A parts
constexpr int S_P = 10 * PT * PT;
// layout (from a 1024-byte aligned base): B weights | input patch / conv2 map (fp32 or fp16 parts) | weights | pooled conv1
constexpr int PNET_SMEM = 1024 + TC_B_BYTES + S_BUF_BYTES + (D_FLOATS + S_P) * 4;
// the tensor-core variant does not keep the fp32 conv3 weights in shared memory: 70 KB per CTA, three CTAs per SM
constexpr int TC_W_SKIP = 144 * 32;
constexpr int PNET_SMEM_TC = PNET_SMEM - TC_W_SKIP * 4;
constexpr int PNET_TC_CTAS = 3;
static_assert(TC_B_BYTES % 1024 == 0 && S_BUF_BYTES % 1024 == 0, "operand buffers must keep the swizzle phase");

struct PnetParams {
  int B, n_levels;
  int lh[VNFR_MAX_LEVELS], lw[VNFR_MAX_LEVELS], oh[VNFR_MAX_LEVELS], ow[VNFR_MAX_LEVELS];
  int tiles_x[VNFR_MAX_LEVELS], tile_off[VNFR_MAX_LEVELS + 1];
  long long level_off[VNFR_MAX_LEVELS], map_off[VNFR_MAX_LEVELS];
};

__device__ __forceinline__ float prelu(float v, float a) { return v > 0.f ? v : v * a; }

// heads (1x1 -> 2 logits + 4 regressions) + softmax + `>= thr` + warp-ballot compaction of ONE cell per lane; acc = conv3
// output before PReLU.  Warp-collective: every lane calls it, `valid` masks the cell.
__device__ __forceinline__ void pnet_cell_epilogue(const PnetParams& p, const float* __restrict__ s_w, const float (&acc)[32], int b, int l,
                                                   int gy, int gx, bool valid, int oh, int ow, float thr, int cap,
                                                   int* __restrict__ cand_count, uint32_t* __restrict__ cand_cell,
                                                   float* __restrict__ cand_score, float4* __restrict__ cand_reg,
                                                   float* __restrict__ dense_prob, float* __restrict__ dense_reg) {
  // the six head outputs as three packed fp32 pairs (fma.rn.f32x2: bit-identical, half the FMA issue slots)
  unsigned long long op[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) op[j] = *reinterpret_cast<const unsigned long long*>(s_w + D_B4 + 2 * j);
#pragma unroll
  for (int co = 0; co < 32; ++co) {
    const unsigned vb = __float_as_uint(prelu(acc[co], s_w[D_A3 + co]));
    const unsigned long long vv = (unsigned long long)vb | ((unsigned long long)vb << 32);
    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(s_w + D_W4 + co * 8);
    const unsigned long long wb = *reinterpret_cast<const unsigned long long*>(s_w + D_W4 + co * 8 + 4);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(op[0]) : "l"(wa.x), "l"(vv));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(op[1]) : "l"(wa.y), "l"(vv));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(op[2]) : "l"(wb), "l"(vv));
  }
  float o[6];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    o[2 * j] = __uint_as_float((unsigned)(op[j] & 0xffffffffull));
    o[2 * j + 1] = __uint_as_float((unsigned)(op[j] >> 32));
  }
  // softmax over the two logits (torch: exp(x - max) / sum)
  const float mx = fmaxf(o[0], o[1]);
  const float e0 = expf(o[0] - mx), e1 = expf(o[1] - mx);
  const float prob = e1 / (e0 + e1);
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
    const int lane = threadIdx.x & 31;
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

template <bool TC>
__global__ void __launch_bounds__(NTHR, TC ? PNET_TC_CTAS : PNET_MIN_CTAS) pnet_kernel(const __grid_constant__ PnetParams p, const float* __restrict__ levels,
                                                    float thr, int cap, int* __restrict__ cand_count,
                                                    uint32_t* __restrict__ cand_cell, float* __restrict__ cand_score,
                                                    float4* __restrict__ cand_reg, float* __restrict__ dense_prob,
                                                    float* __restrict__ dense_reg, const float* __restrict__ g_pnet_w) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint4* g_pnet_w3h = reinterpret_cast<const uint4*>(g_pnet_w + D_FLOATS);
  __shared__ __align__(8) uint64_t s_bar;             // tensor-core path: MMAs of the current tile have completed
  __shared__ uint32_t s_tmem;
  uint8_t* smem_al = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_bw = smem_al;                                                   // conv3 B operand (tensor-core path)
  float* s_buf = reinterpret_cast<float*>(smem_al + TC_B_BYTES);             // input patch [3][IT][ITP], then conv2 map:
                                                                             // fp32 [16][C2T][C2T] or fp16 hi | lo A operand
  float* s_w = reinterpret_cast<float*>(smem_al + TC_B_BYTES + S_BUF_BYTES); // weights
  float* s_p = s_w + D_FLOATS - (TC ? TC_W_SKIP : 0);                        // pooled conv1 map [10][PT][PT]
  const float* s_wt = s_w - (TC ? TC_W_SKIP : 0);     // rows after the conv3 weights (bias / PReLU of conv3, heads): s_wt[D_...]
  const int tid = threadIdx.x;
  for (int i = tid; i < D_FLOATS / 4; i += NTHR) {
    if (TC && i >= D_W3 / 4 && i < D_B3 / 4) continue;                       // fp32 conv3 weights: not needed on the tensor-core path
    reinterpret_cast<float4*>(s_w)[(TC && i >= D_B3 / 4) ? i - TC_W_SKIP / 4 : i] = reinterpret_cast<const float4*>(g_pnet_w)[i];
  }
  const uint32_t a_hi = smem_u32(s_buf), a_lo = a_hi + TC_A_PART_BYTES, b_addr = smem_u32(s_bw), bar = smem_u32(&s_bar);
  uint32_t tmem_base = 0, phase = 0;
  if (TC) {
    for (int i = tid; i < TC_B_BYTES / 16; i += NTHR) reinterpret_cast<uint4*>(s_bw)[i] = g_pnet_w3h[i];
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    if (tid < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_proxy_async_smem();                      // the B operand was written through the generic proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tmem_base = s_tmem;
  }

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
        // packed fp32 pairs (fma.rn.f32x2): accp[q][h] = channels (2h, 2h+1) of conv position q
        unsigned long long accp[4][5];
#pragma unroll
        for (int h = 0; h < 5; ++h) {
          const unsigned long long b2 = *reinterpret_cast<const unsigned long long*>(s_w + D_B1 + 2 * h);
          accp[0][h] = b2; accp[1][h] = b2; accp[2][h] = b2; accp[3][h] = b2;
        }
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const unsigned long long* wr = reinterpret_cast<const unsigned long long*>(s_w + D_W1 + ((ci * 3 + ky) * 3 + kx) * 12);
              const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(wr), wb = *reinterpret_cast<const ulonglong2*>(wr + 2);
              const unsigned long long w[5] = {wa.x, wa.y, wb.x, wb.y, wr[4]};
              const float pv[4] = {patch[ci][ky][kx], patch[ci][ky][kx + 1], patch[ci][ky + 1][kx], patch[ci][ky + 1][kx + 1]};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const unsigned vb = __float_as_uint(pv[q]);
                const unsigned long long vv = (unsigned long long)vb | ((unsigned long long)vb << 32);
#pragma unroll
                for (int h = 0; h < 5; ++h) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accp[q][h]) : "l"(w[h]), "l"(vv));
              }
            }
        float acc[4][10];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int h = 0; h < 5; ++h) {
            acc[q][2 * h] = __uint_as_float((unsigned)(accp[q][h] & 0xffffffffull));
            acc[q][2 * h + 1] = __uint_as_float((unsigned)(accp[q][h] >> 32));
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
      // packed fp32 pairs (FFMA2, fma.rn.f32x2): accp[j][q] = channels (2q, 2q+1) of position j -- half the FMA issue slots of
      // the scalar loop (conv2 was 30 % of the kernel's instructions), bit-identical results
      unsigned long long accp[3][8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned long long b2 = *reinterpret_cast<const unsigned long long*>(s_w + D_B2 + 2 * q);
        accp[0][q] = b2; accp[1][q] = b2; accp[2][q] = b2;
      }
#pragma unroll 2
      for (int ci = 0; ci < 10; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(s_w + D_W2 + ((ci * 3 + ky) * 3 + kx) * 16);
            const ulonglong2 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
            const unsigned long long w[8] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y, w3.x, w3.y};
            const int o = (ci * PT + ky) * PT + kx;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const unsigned vb = __float_as_uint(s_p[off[j] + o]);
              const unsigned long long vv = (unsigned long long)vb | ((unsigned long long)vb << 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(accp[j][q]) : "l"(w[q]), "l"(vv));
            }
          }
      float acc[3][16];
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          acc[j][2 * q] = __uint_as_float((unsigned)(accp[j][q] & 0xffffffffull));
          acc[j][2 * q + 1] = __uint_as_float((unsigned)(accp[j][q] >> 32));
        }
      if (TC) {
        // fp16 hi | lo parts, one 32-byte swizzled row per pixel: 16-byte chunk c at row*32 + ((c ^ address bit 7) << 4)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint32_t hw[8], lw[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float v0 = prelu(acc[j][2 * q], s_w[D_A2 + 2 * q]), v1 = prelu(acc[j][2 * q + 1], s_w[D_A2 + 2 * q + 1]);
            const __half2 h = __floats2half2_rn(v0, v1);
            const __half2 lo = __floats2half2_rn(v0 - __low2float(h), v1 - __high2float(h));
            hw[q] = *reinterpret_cast<const uint32_t*>(&h);
            lw[q] = *reinterpret_cast<const uint32_t*>(&lo);
          }
          const uint32_t row = a_hi + 32u * (uint32_t)oo[j];
          const uint32_t sw = ((row >> 7) & 1u) << 4;
          tc::sts128(row + sw, make_uint4(hw[0], hw[1], hw[2], hw[3]));
          tc::sts128(row + (sw ^ 16u), make_uint4(hw[4], hw[5], hw[6], hw[7]));
          tc::sts128(row + TC_A_PART_BYTES + sw, make_uint4(lw[0], lw[1], lw[2], lw[3]));
          tc::sts128(row + TC_A_PART_BYTES + (sw ^ 16u), make_uint4(lw[4], lw[5], lw[6], lw[7]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int co = 0; co < 16; ++co) s_buf[co * C2T * C2T + oo[j]] = prelu(acc[j][co], s_w[D_A2 + co]);
      }
    }
    if (TC) { tc::fence_proxy_async_smem(); tc::tc_fence_before(); }
    __syncthreads();

    if (TC) {
      // ---- conv3 on the tensor cores: 3 row tiles x 9 taps x 3 products, fp32 accumulation in TMEM
      if (tid < 32) {
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t idesc = tc::make_idesc_f16(32, 1);
#pragma unroll 1
          for (int mt = 0; mt < 3; ++mt) {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t shift = 32u * (uint32_t)(128 * mt + (tap / 3) * C2T + (tap % 3));
              const uint64_t ah = tc::make_sw_desc(a_hi + shift, 32, 0), al = tc::make_sw_desc(a_lo + shift, 32, 0);
              const uint64_t bh = tc::make_sw_desc(b_addr + (uint32_t)(tap * 2) * 1024u, 32, 0);
              const uint64_t bl = tc::make_sw_desc(b_addr + (uint32_t)(tap * 2 + 1) * 1024u, 32, 0);
              tc::umma_bf16(tmem_base + 32u * mt, ah, bh, idesc, tap != 0);
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
      const uint32_t t_lane = tmem_base + ((uint32_t)(tid & ~31) << 16);          // this warp's 32 TMEM lanes
#pragma unroll 1
      for (int mt = 0; mt < 3; ++mt) {
        float acc[32];
        __syncwarp();
        tc::tmem_ld16_issue(t_lane + 32u * mt, acc);
        tc::tmem_ld16_issue(t_lane + 32u * mt + 16u, acc + 16);
        tc::tmem_ld_wait(acc);
        tc::tmem_ld_wait(acc + 16);
#pragma unroll
        for (int co = 0; co < 32; ++co) acc[co] += s_wt[D_B3 + co];
        const int r = 128 * mt + tid;                   // accumulator row = raster index in the 18-wide conv2 map
        const int cy = r / C2T, cx = r - cy * C2T;
        const int gy = ty0 + cy, gx = tx0 + cx;
        const bool valid = cy < T && cx < T && gy < oh && gx < ow;
        pnet_cell_epilogue(p, s_wt, acc, b, l, gy, gx, valid, oh, ow, thr, cap, cand_count, cand_cell, cand_score, cand_reg,
                           dense_prob, dense_reg);
      }
      tc::tc_fence_before();                            // TMEM reads are ordered before the next tile's MMAs (barrier at loop top)
      continue;
    }

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
      const int gy = ty0 + y + 8 * h, gx = tx0 + x;
      pnet_cell_epilogue(p, s_wt, acc[h], b, l, gy, gx, gy < oh && gx < ow, oh, ow, thr, cap, cand_count, cand_cell, cand_score,
                         cand_reg, dense_prob, dense_reg);
    }
  }
  if (TC) {
    tc::tc_fence_before();
    __syncthreads();
    if (tid < 32) {
      tc::tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
  }
}

}  // namespace

extern "C" int vnfr_pnet_packed_bytes(void) { return PNET_PACKED_BYTES; }

extern "C" int vnfr_pnet_pack_weights(const float* packed_host, int n_floats, void* out_host, int out_bytes) {
  VNFR_REQUIRE(packed_host != nullptr && n_floats == PNET_FLOATS, "P-Net packed weights must hold 6632 floats");
  VNFR_REQUIRE(out_host != nullptr && out_bytes == PNET_PACKED_BYTES, "out_host must hold vnfr_pnet_packed_bytes() bytes");
  float* d = reinterpret_cast<float*>(out_host);          // repack [co][k] -> [k][co] (padded)
  memset(out_host, 0, PNET_PACKED_BYTES);
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
  // conv3 weights as fp16 hi / lo parts in the swizzled B-operand layout of the tensor-core path
  uint16_t* bw = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(out_host) + D_FLOATS * 4);
  for (int tap = 0; tap < 9; ++tap)
    for (int n = 0; n < 32; ++n)
      for (int ci = 0; ci < 16; ++ci) {
        const float wv = h[W3 + n * 144 + ci * 9 + tap];
        const __half hi = __float2half_rn(wv);
        const __half lo = __float2half_rn(wv - __half2float(hi));
        const int c = ci >> 3, e = ci & 7;
        const int off = n * 32 + ((c ^ ((n >> 2) & 1)) << 4) + 2 * e;          // bytes inside the (tap, part) block
        bw[((tap * 2 + 0) * 1024 + off) / 2] = __half_as_ushort(hi);
        bw[((tap * 2 + 1) * 1024 + off) / 2] = __half_as_ushort(lo);
      }
  return VNFR_OK;
}

extern "C" int vnfr_pnet_sweep_compact(const VnfrPyramid* pyr, const float* levels, const void* pnet_weights, float threshold, int cap,
                                       int32_t* cand_count, uint32_t* cand_cell, float* cand_score, float* cand_reg,
                                       float* dense_prob, float* dense_reg, void* stream) {
  VNFR_REQUIRE(pyr != nullptr && cap > 0, "bad arguments");
  VNFR_REQUIRE(pnet_weights != nullptr && ((uintptr_t)pnet_weights % 16) == 0, "pnet_weights must be a 16-byte aligned device buffer");
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
  static VnfrPerDevice attr_once = {};
  static const bool fma_conv3 = getenv("VNFR_PNET_FMA") != nullptr;
  if (vnfr_first_on_device(attr_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(pnet_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PNET_SMEM_TC));
    VNFR_CUDA(cudaFuncSetAttribute(pnet_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PNET_SMEM));
  }
  const int total = pyr->B * tiles;
  const int per_sm = fma_conv3 ? PNET_MIN_CTAS : PNET_TC_CTAS;
  const int grid = total < 148 * per_sm ? total : 148 * per_sm;      // persistent CTAs
  if (fma_conv3)
    pnet_kernel<false><<<grid, NTHR, PNET_SMEM, (cudaStream_t)stream>>>(p, levels, threshold, cap, cand_count, cand_cell, cand_score,
                                                                       reinterpret_cast<float4*>(cand_reg), dense_prob, dense_reg,
                                                                       reinterpret_cast<const float*>(pnet_weights));
  else
    pnet_kernel<true><<<grid, NTHR, PNET_SMEM_TC, (cudaStream_t)stream>>>(p, levels, threshold, cap, cand_count, cand_cell, cand_score,
                                                                      reinterpret_cast<float4*>(cand_reg), dense_prob, dense_reg,
                                                                      reinterpret_cast<const float*>(pnet_weights));
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
