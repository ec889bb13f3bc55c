// Scale-pyramid plan + area-resize/normalise of every level in one launch.
//
// Replaces detect_face.py:46 (u8 NHWC -> fp32 NCHW copy of the whole batch, never materialised here), :48-60 (scale
// list), :71-72 + :304-306 (imresample(area) per scale + (x-127.5)*0.0078125).  The u8 frame is the only input: every
// level pixel is the mean of its adaptive-average-pooling window [floor(i*H/oh), ceil((i+1)*H/oh)) computed as an exact
// integer sum followed by sum/kh/kw in fp32 -- bit-identical to torch's CPU adaptive_avg_pool2d on integer pixels.
// Main path: pyramid_strip_kernel -- ONE pass over the frame feeds every level: a CTA streams a strip of source rows
// once and keeps, per level, exact column sums of the level's current window rows in registers (see the kernel).
// Fallback (rows not 16-byte aligned, upscaling levels, windows taller than 257 rows): pyramid_rows_kernel, one task per
// output row, frame-major so that the passes over one frame hit L2.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

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


// ---------------------------------------------------------------------------------------------------------------------
// Strip kernel: every source row is streamed once per GROUP of up to PS_NL levels (two groups for the 9 levels of a
// 1080p / min_face_size 50 pyramid; the second pass over a frame is served by L2).
//   CTA = (frame b, level group, strip of source rows [ys, ye), column tile of 16-byte chunks [c0, c1)); thread t =
//   chunk c0 + t.  The CTA owns, for every level of its group, the output rows whose window STARTS in its strip and the
//   output columns whose window starts in its tile; it streams source rows ys, ys+1, ... through a PS_D-deep cp.async
//   ring (every thread fetches and later reads its own 16 bytes: no barrier) until the last owned window is complete.  Fine levels (short windows) get many short strips, coarse
//   levels fewer, taller ones (the rows read past the strip end are the tallest window of the group).
//   Row sums by running prefix: the thread keeps ONE running sum R of its 16 byte columns over all rows streamed so far
//   (8 registers of two u16 lanes each, wrapping) and, per level, a snapshot of R taken where the level's current window
//   starts.  Window sum = R - snapshot as plain 32-bit subtraction of the packed registers: with true lane sums (L, Hh)
//   the register holds (L + 65536 Hh) mod 2^32, so the difference is (l + 65536 h) mod 2^32 for the window's lane sums
//   l, h -- exact because a window has at most 257 rows (l, h <= 65535).  The per-row cost is independent of the number
//   of levels.
//   Adaptive-pooling windows of consecutive output rows overlap by at most one source row (downscaling), so the next
//   window's snapshot is R after this row or R before it.  What happens at row y for level l (first window starts /
//   window ends / next window includes this row) only depends on (y, l): a per-CTA table of bit masks in shared memory.
//   Flush: window column sums -> shared memory (u16, double buffered, ONE barrier per flush), then the CTA's threads add
//   up the kw column sums of every owned output pixel (window bounds from a per-CTA table) and write the fp32 row.
constexpr int PS_THREADS = 384;
constexpr int PS_NL = 5;           // levels per group (snapshots: 40 registers)
constexpr int PS_G = 5;            // groups per launch (VNFR_MAX_LEVELS = 24 <= 25)
constexpr int PS_D = 8;            // source rows in flight per CTA: cp.async ring in shared memory (power of two)
constexpr int PS_ROW = PS_THREADS * 16;   // bytes of one ring slot
constexpr int PS_TABLE = 1024;     // rows a CTA may stream (strip + the tallest window)

struct PyrStripGroup {
  int n_levels;
  int lh[PS_NL], lw[PS_NL];
  long long level_off[PS_NL];
  int strips, strip_h;             // source-row strips per frame
  int cta_off;                     // first CTA of this group inside a frame's CTA range
};

struct PyrStripParams {
  int B, H, W, n_groups;
  int tiles, tile_chunks;          // column tiles per row, owned 16-byte chunks per tile
  int total_chunks;                // W*3/16
  int ctas_per_frame;
  int tab_entries;                 // capacity of the window-bounds table (entries)
  PyrStripGroup g[PS_G];
};

__device__ __forceinline__ void ps_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ps_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ps_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct Packed8 { uint32_t v[8]; };   // [0..3] even bytes of the 4 words, [4..7] odd bytes

// Output row of one level from the window column sums in shared memory.
__device__ __forceinline__ void strip_reduce_row(const uint16_t* __restrict__ cbase, const uint32_t* __restrict__ tab, int n_own,
                                              float* __restrict__ orow, size_t plane, float kh, int tid) {
  for (int i = tid; i < 3 * n_own; i += PS_THREADS) {
    const int ch = (i >= n_own) + (i >= 2 * n_own), o = i - ch * n_own;
    const uint32_t t = tab[o];
    const int x0 = (int)(t & 0xFFFFu), kw = (int)(t >> 16);
    const uint16_t* cp = cbase + (3 * x0 + ch);
    uint32_t sum = 0;
    for (int x = 0; x < kw; ++x, cp += 3) sum += *cp;
    orow[ch * plane + o] = mul_rn(sub_rn(div_rn(div_rn((float)sum, kh), (float)kw), 127.5f), 0.0078125f);
  }
}

__global__ void __launch_bounds__(PS_THREADS, 2) pyramid_strip_kernel(const __grid_constant__ PyrStripParams p,
                                                                      const uint8_t* __restrict__ frames,
                                                                      float* __restrict__ levels) {
  extern __shared__ __align__(16) uint8_t ps_smem[];
  // layout: row ring [PS_D][PS_ROW] | masks [PS_TABLE] u32 | column sums 2 x [PS_THREADS*16] u16 | window bounds [tab_entries] u32
  uint8_t* s_ring = ps_smem;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(ps_smem + PS_D * PS_ROW);
  uint16_t* s_col = reinterpret_cast<uint16_t*>(ps_smem + PS_D * PS_ROW + PS_TABLE * 4);
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(ps_smem + PS_D * PS_ROW + PS_TABLE * 4 + 2 * PS_THREADS * 32);
  __shared__ int s_rows;
  __shared__ int s_taboff[PS_NL + 1], s_oxf[PS_NL];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.ctas_per_frame;
  int rem = blockIdx.x - b * p.ctas_per_frame;
  int gi = 0;
  while (gi + 1 < p.n_groups && rem >= p.g[gi + 1].cta_off) ++gi;
  const PyrStripGroup& G = p.g[gi];
  rem -= G.cta_off;
  const int tile = rem % p.tiles, strip = rem / p.tiles;
  const int H = p.H, W = p.W, nl = G.n_levels;
  const int ys = strip * G.strip_h, ye = min(H, ys + G.strip_h);
  const int c0 = tile * p.tile_chunks, c1 = min(c0 + p.tile_chunks, p.total_chunks);
  // owned pixels: first byte inside [16*c0, 16*c1)
  const int px0 = (16 * c0 + 2) / 3, px1 = (c1 == p.total_chunks) ? W : (16 * c1 + 2) / 3;

  // ---- per-CTA tables
  if (tid == 0) {
    int rows = 0, off = 0;
    for (int l = 0; l < nl; ++l) {
      const int lh = G.lh[l], lw = G.lw[l];
      const int oyf = (ys * lh + H - 1) / H, oye = min(lh, (ye * lh + H - 1) / H);
      if (oye > oyf) rows = max(rows, (oye * H + lh - 1) / lh - ys);
      const int oxf = (px0 * lw + W - 1) / W, oxe = (px1 == W) ? lw : (px1 * lw + W - 1) / W;
      s_oxf[l] = oxf;
      s_taboff[l] = off;
      off = min(off + max(oxe - oxf, 0), p.tab_entries);      // the host sizes the table so that this never clips
    }
    s_taboff[nl] = off;
    s_rows = min(rows, PS_TABLE);                              // likewise
  }
  __syncthreads();
  const int rows = s_rows;
  // bit l = the first owned window of level l starts at this row, bit 10+l = a window ends at this row (flush),
  // bit 20+l = the next window includes this row
  for (int r = tid; r < rows; r += PS_THREADS) {
    const int y = ys + r;
    uint32_t m = 0;
    for (int l = 0; l < nl; ++l) {
      const int lh = G.lh[l];
      const int oyf = (ys * lh + H - 1) / H, oye = min(lh, (ye * lh + H - 1) / H);
      const int a = (y * lh) / H, bb = ((y + 1) * lh + H - 1) / H - 1;       // windows containing row y: a..bb
      const bool own_a = a >= oyf && a < oye, own_b = bb >= oyf && bb < oye;
      const bool F = own_a && ((a + 1) * H + lh - 1) / lh == y + 1;
      if (oye > oyf && (oyf * H) / lh == y) m |= 1u << l;
      if (F) m |= 1u << (10 + l);
      if (F && bb > a && own_b) m |= 1u << (20 + l);
    }
    s_mask[r] = m;
  }
  // window bounds of the owned output columns: x0 | kw << 16
  for (int l = 0; l < nl; ++l) {
    const int lw = G.lw[l], oxf = s_oxf[l], n = s_taboff[l + 1] - s_taboff[l];
    for (int o = tid; o < n; o += PS_THREADS) {
      const unsigned ox = (unsigned)(oxf + o);
      const unsigned x0 = (ox * (unsigned)W) / (unsigned)lw;
      const unsigned x1 = ((ox + 1u) * (unsigned)W + (unsigned)lw - 1u) / (unsigned)lw;
      s_tab[s_taboff[l] + o] = x0 | ((x1 - x0) << 16);
    }
  }
  __syncthreads();

  Packed8 R;                      // running column sums
  Packed8 snap[PS_NL];            // R where the level's current window starts
#pragma unroll
  for (int k = 0; k < 8; ++k) R.v[k] = 0u;
#pragma unroll
  for (int l = 0; l < PS_NL; ++l)
#pragma unroll
    for (int k = 0; k < 8; ++k) snap[l].v[k] = 0u;

  const size_t rowbytes = (size_t)W * 3;
  const bool act = (c0 + tid) < p.total_chunks;
  const uint8_t* src = frames + ((size_t)b * H + ys) * rowbytes + (size_t)(c0 + tid) * 16;
  const uint16_t* cbase0 = s_col - 16 * c0;
  const uint32_t ring_addr = smem_u32(s_ring) + tid * 16;
  const uint4* ring_ptr = reinterpret_cast<const uint4*>(s_ring) + tid;
#pragma unroll
  for (int u = 0; u < PS_D; ++u) {
    if (act && u < rows) ps_cp_async16(ring_addr + u * PS_ROW, src + (size_t)u * rowbytes);
    ps_cp_async_commit();
  }
  int parity = 0;
  for (int r = 0; r < rows; ++r) {
    ps_cp_async_wait<PS_D - 1>();                       // this thread's 16 bytes of row r have landed
    const int slot = r & (PS_D - 1);
    uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
    if (act) w4 = ring_ptr[slot * (PS_ROW / 16)];
    const uint32_t m = s_mask[r];
    const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
    Packed8 row;
#pragma unroll
    for (int k = 0; k < 4; ++k) { row.v[k] = w[k] & 0x00FF00FFu; row.v[4 + k] = (w[k] >> 8) & 0x00FF00FFu; }
    if (m & 0x3FFu) {                       // first owned window of some level starts here: snapshot BEFORE this row
#pragma unroll
      for (int l = 0; l < PS_NL; ++l)
        if ((m >> l) & 1u) snap[l] = R;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) R.v[k] += row.v[k];
    // refill the slot (its data is in registers now) with row r + PS_D
    if (act && r + PS_D < rows) ps_cp_async16(ring_addr + slot * PS_ROW, src + (size_t)(r + PS_D) * rowbytes);
    ps_cp_async_commit();
    if (m & (0x3FFu << 10)) {
#pragma unroll
      for (int l = 0; l < PS_NL; ++l) {
        if ((m >> (10 + l)) & 1u) {
          // ---- flush level l: output row oy = the window that ends at source row y
          const bool incl = (m >> (20 + l)) & 1u;          // the next window includes this row
          uint32_t d[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            d[k] = R.v[k] - snap[l].v[k];
            snap[l].v[k] = incl ? R.v[k] - row.v[k] : R.v[k];
          }
          uint16_t* colb = s_col + parity * (PS_THREADS * 16);
          uint4 q0, q1;                                    // u16 sums of byte columns 0..7 / 8..15
          q0.x = __byte_perm(d[0], d[4], 0x5410); q0.y = __byte_perm(d[0], d[4], 0x7632);
          q0.z = __byte_perm(d[1], d[5], 0x5410); q0.w = __byte_perm(d[1], d[5], 0x7632);
          q1.x = __byte_perm(d[2], d[6], 0x5410); q1.y = __byte_perm(d[2], d[6], 0x7632);
          q1.z = __byte_perm(d[3], d[7], 0x5410); q1.w = __byte_perm(d[3], d[7], 0x7632);
          *reinterpret_cast<uint4*>(colb + 16 * tid) = q0;
          *reinterpret_cast<uint4*>(colb + 16 * tid + 8) = q1;
          __syncthreads();       // the other buffer is free again: every thread finished the previous flush's reads
          const int y = ys + r;
          const int lh = G.lh[l], lw = G.lw[l];
          const int oy = (y * lh) / H;
          const float kh = (float)(y + 1 - (oy * H) / lh);
          const size_t plane = (size_t)lh * lw;
          strip_reduce_row(cbase0 + parity * (PS_THREADS * 16), s_tab + s_taboff[l], s_taboff[l + 1] - s_taboff[l],
                           levels + G.level_off[l] + (size_t)b * 3 * plane + (size_t)oy * lw + s_oxf[l], plane, kh, tid);
          parity ^= 1;
        }
      }
    }
  }
  ps_cp_async_wait<0>();
}

// Strip plan of the whole pyramid; false = this geometry needs the fallback kernel.
bool plan_strips(const VnfrPyramid* pyr, PyrStripParams* q) {
  const int H = pyr->H, W = pyr->W, L = pyr->n_levels;
  if (((size_t)W * 3) % 16 != 0 || L > PS_NL * PS_G) return false;
  memset(q, 0, sizeof(*q));
  q->B = pyr->B; q->H = H; q->W = W;
  q->total_chunks = W * 3 / 16;
  int kw_max = 1;
  for (int l = 0; l < L; ++l) {
    const int lh = pyr->lh[l], lw = pyr->lw[l];
    if (lh > H || lw > W || lh < 1 || lw < 1) return false;           // upscaling: windows may overlap by more than a row
    if ((H + lh - 1) / lh + 1 > 257) return false;                    // u16 lanes: 257 * 255 = 65535
    kw_max = std::max(kw_max, (W + lw - 1) / lw + 1);
  }
  const int halo_chunks = (3 * kw_max + 15) / 16 + 1;
  if (halo_chunks >= PS_THREADS / 2) return false;
  const int own_max = PS_THREADS - halo_chunks;
  q->tiles = (q->total_chunks + own_max - 1) / own_max;
  q->tile_chunks = (q->total_chunks + q->tiles - 1) / q->tiles;
  int ctas = 0;
  for (int l0 = 0; l0 < L; l0 += PS_NL) {
    PyrStripGroup& g = q->g[q->n_groups++];
    g.n_levels = std::min(PS_NL, L - l0);
    int kh_max = 1, tab = 0;
    for (int l = 0; l < g.n_levels; ++l) {
      g.lh[l] = pyr->lh[l0 + l]; g.lw[l] = pyr->lw[l0 + l]; g.level_off[l] = pyr->level_off[l0 + l];
      kh_max = std::max(kh_max, (H + g.lh[l] - 1) / g.lh[l] + 1);
      // owned output columns of one tile (+2 for the rounding at both tile edges)
      tab += (int)(((long long)g.lw[l] * (16 * q->tile_chunks + 5) / 3 + W - 1) / W) + 2;
    }
    q->tab_entries = std::max(q->tab_entries, tab);
    // strips: rows streamed by all CTAs over the 2 x 148 resident CTAs + one CTA's rows for the tail
    double best = -1;
    for (int s = 1; s <= 256 && s <= H; ++s) {
      const int sh = (H + s - 1) / s, ns = (H + sh - 1) / sh;
      if (sh + kh_max > PS_TABLE) continue;
      const double rows = sh + kh_max;
      const double cost = (double)pyr->B * q->tiles * ns * rows / 296.0 + rows;
      if (best < 0 || cost < best) { best = cost; g.strip_h = sh; g.strips = ns; }
    }
    if (best < 0) return false;
    g.cta_off = ctas;
    ctas += g.strips * q->tiles;
  }
  q->ctas_per_frame = ctas;
  return true;
}

size_t strip_smem_bytes(const PyrStripParams& q) {
  return (size_t)PS_D * PS_ROW + (size_t)PS_TABLE * 4 + 2 * PS_THREADS * 32 + (size_t)q.tab_entries * 4;
}

}  // namespace

extern "C" int vnfr_pyramid_resize_norm(const VnfrPyramid* pyr, const uint8_t* frames, float* levels, void* stream) {
  VNFR_REQUIRE(pyr != nullptr, "pyramid plan is null");
  if (pyr->B == 0 || pyr->n_levels == 0) return VNFR_OK;
  VNFR_REQUIRE((long long)pyr->W * pyr->W < (1ll << 31) && (long long)pyr->H * pyr->H < (1ll << 31), "frame too large");
  static const bool force_rows = getenv("VNFR_PYRAMID_ROWS") != nullptr;      // profiling: the one-task-per-row kernel
  if (!force_rows && ((uintptr_t)frames % 16) == 0) {
    PyrStripParams q;
    if (plan_strips(pyr, &q) && strip_smem_bytes(q) <= 110 * 1024) {
      static VnfrPerDevice strip_attr_once = {};
      if (vnfr_first_on_device(strip_attr_once)) {
        VNFR_CUDA(cudaFuncSetAttribute(pyramid_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      }
      const long long grid = (long long)q.B * q.ctas_per_frame;
      pyramid_strip_kernel<<<(unsigned)grid, PS_THREADS, strip_smem_bytes(q), (cudaStream_t)stream>>>(q, frames, levels);
      ++g_vnfr_launches;
      VNFR_CHECK_LAUNCH();
      return VNFR_OK;
    }
  }
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
  const bool vec = ((size_t)p.W * 3) % 16 == 0 && ((uintptr_t)frames % 16) == 0;
  const int per_sm = smem > 0 ? (int)((220 * 1024) / (smem + 1024)) : 8;
  long long grid = 148LL * (per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm));
  if (grid > n_tasks) grid = n_tasks;
  static VnfrPerDevice attr_once = {};
  if (vnfr_first_on_device(attr_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(pyramid_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(pyramid_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (vec) pyramid_rows_kernel<16><<<(int)grid, PYR_THREADS, smem, (cudaStream_t)stream>>>(p, frames, levels);
  else pyramid_rows_kernel<1><<<(int)grid, PYR_THREADS, smem, (cudaStream_t)stream>>>(p, frames, levels);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
