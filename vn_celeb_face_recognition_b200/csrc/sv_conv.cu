// Shifted-view convolution on tcgen05: stride-1 k x k convolutions whose input band stays in shared memory.
//
//   out[p] = sum over taps (ky,kx) of  W_tap . in[p + ky*P + kx]        p = linear index into the zero-padded band
//
// Instead of gathering an im2col matrix (9x / 7x re-reads of every input pixel through cp.async, the limiter of the
// generic kernel in igemm_conv.cu), ONE TMA box load brings a band of input rows -- `nb` images x `rows_in` rows x `P`
// pixels x CK channels per plane, out-of-bounds pixels zero-filled by the TMA unit, which is the convolution's padding
// -- into 128B/64B-swizzled shared memory with one pixel per swizzle row.  The A operand of tap (ky,kx) is then simply
// the same buffer viewed from a start address shifted by (ky*P + kx) rows: the tap loop issues tcgen05.mma over
// shifted descriptors, no data moves.  Accumulator rows whose linear index falls on padding columns / halo rows are
// computed and discarded by the epilogue.  Weights are either resident in shared memory for the whole persistent CTA
// (small layers: stem, Block35) or streamed through a TMA ring shared by `MT` accumulators (Block17/Block8 1x7, 7x1,
// 1x3, 3x1 and the 192/256-channel 3x3s).
//
// Replaces the k x k BasicConv2d call sites of inception_resnet_v1.py:12-33 (conv2d_2a/2b, Block35 branch1/2 3x3,
// Block17 1x7/7x1, Block8 1x3/3x1, mixed_6a/7a stride-1 3x3).
//
// Warp roles (352 threads): warps 0-7 epilogue (TMEM lane quarter = warp & 3, two warps per quarter on alternate
// 16-column chunks), warp 8 A-band TMA producer, warp 9 TMEM alloc + MMA issue, warp 10 weight TMA producer.
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int SV_THREADS = 352;
constexpr int SV_EPI_THREADS = 256;

struct SvParams {
  ConvParams p;            // destinations / bias / residual / relu / dtype (used by the shared epilogue)
  int ck;                  // channels per plane (32 or 64); row_bytes = 2*ck = swizzle span
  int n_chunks, taps, ksteps;
  int P;                   // padded row pitch in pixels (in_w + 2*pad_w)
  int R;                   // output rows per band
  int rows_in;             // R + kh - 1
  int nb;                  // images per band
  int img_px;              // rows_in * P
  int tiles_per_band;      // 128-row accumulator tiles per band
  int bands_y, n_bands;
  int plane_bytes;         // bytes of one channel plane of the A band buffer (1024-aligned)
  int a_buf_bytes;         // n_chunks * plane_bytes
  int a_box_bytes;         // bytes one TMA box delivers (per plane)
  int b_resident;          // 1: all weight tiles live in shared memory; 0: ring of `stages`
  int b_tile_bytes;        // block_n * row_bytes
  int stages;
  int mt;                  // accumulator tiles sharing one weight stage (ring mode)
  int nbuf;                // TMEM accumulator buffers (2 * mt)
  int tmem_cols;           // columns per buffer
  int use_base_offset;     // descriptor base-offset field = (addr >> 7) & 7
  int epi_split;           // 1: the two epilogue warp groups take alternate tiles (narrow tiles)
  int epi_stage_bytes;     // > 0: per-epilogue-warp shared-memory slot (32 rows) through which narrow rows reach global memory
                           // as contiguous 16-byte chunks (see epilogue_row_narrow_staged)
  int a_bufs;              // band buffers: 2 (next band loads under this band's MMAs) or 1 (large bands)
  int order;               // MMA issue order inside a K step: 0 = accumulator-major, 1 = rotate over accumulators per K slice
  const int* n_img_dev;    // nullable: device-side count of valid images (bands beyond it are skipped)
  uint32_t koff[128];      // descriptor offset (16-byte units) of K step ks = (tap, chunk): chunk plane + (ky*P + kx) rows
};

// Optional cycle breakdown of CTA 0 (tools/sv_probe.py): [0] kernel, [1] MMA thread waiting for the A band, [2] for a
// drained accumulator, [3] for a weight stage, [5] epilogue thread 0 waiting for an accumulator, [6] epilogue work,
// [7] A producer waiting for a free band buffer.
__device__ long long* g_sv_dbg = nullptr;
#define SV_T0() (dbg ? clock64() : 0ll)
#define SV_ACC(i, t0) do { if (dbg) dbg[i] += clock64() - (t0); } while (0)

// Narrow tiles (cout = 32 / 64 = the whole NHWC pixel, 64 / 128 bytes): with one accumulator row per thread, a warp-level 16-byte
// store of epilogue_row_narrow touches 32 different 128-byte lines (ncu on conv2d_2b: LSU wavefronts at 70 % of peak, the top
// pipe of the kernel).  The valid rows of a warp are CONSECUTIVE output pixels (the raster only skips padding), so the warp
// parks its rows densely in a shared-memory slot (XOR-swizzled chunks) and streams the slot out with consecutive lanes on
// consecutive 16-byte chunks: 4 full lines per store instruction.
template <bool F16, int NC>
__device__ __forceinline__ void epilogue_row_narrow_staged(const ConvParams& p, const float* sb, uint32_t t_row, int m, bool row_ok,
                                                           uint32_t slot_smem, int lane) {
  constexpr int CH = NC / 8;                 // 16-byte chunks per row
  float v[NC];
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld16_issue(t_row + (uint32_t)(16 * c), v + 16 * c);
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld_wait(v + 16 * c);
  const unsigned mask = __ballot_sync(0xffffffffu, row_ok);
  if (mask == 0u) return;
  const int slot = __popc(mask & ((1u << lane) - 1u)), nvalid = __popc(mask);
  const int m_first = __shfl_sync(0xffffffffu, m, __ffs(mask) - 1);
  if (row_ok) {
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + 4 * i);
      v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    const uint32_t row = slot_smem + (uint32_t)slot * (uint32_t)(NC * 2);
#pragma unroll
    for (int qq = 0; qq < CH; ++qq)
      sts128(row + (uint32_t)((qq ^ (slot & (CH - 1))) << 4),
             make_uint4(pack2<F16>(v[8 * qq], v[8 * qq + 1]), pack2<F16>(v[8 * qq + 2], v[8 * qq + 3]),
                        pack2<F16>(v[8 * qq + 4], v[8 * qq + 5]), pack2<F16>(v[8 * qq + 6], v[8 * qq + 7])));
  }
  __syncwarp();
  uint4* g = reinterpret_cast<uint4*>(p.out0 + (size_t)m_first * NC);
  for (int i = lane; i < nvalid * CH; i += 32) {
    const int r = i / CH, ch = i - r * CH;
    g[i] = lds128(slot_smem + (uint32_t)r * (uint32_t)(NC * 2) + (uint32_t)((ch ^ (r & (CH - 1))) << 4));
  }
  __syncwarp();                              // the slot is rewritten by this warp's next tile
}

// The same for fp32 destinations (O-Net's split-precision conv2: 64 floats per row, + bias + PReLU): slot rows are padded by
// one chunk (17), which keeps both the row-per-thread writes and the chunk-per-lane reads conflict-free.
template <int NC>
__device__ __forceinline__ void epilogue_row_narrow_staged_f32(const ConvParams& p, const float* sb, uint32_t t_row, int m, bool row_ok,
                                                               uint32_t slot_smem, int lane) {
  constexpr int CH = NC / 4, RS = (CH + 1) * 16;      // chunks per row, padded row stride in bytes
  float v[NC];
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld16_issue(t_row + (uint32_t)(16 * c), v + 16 * c);
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld_wait(v + 16 * c);
  const unsigned mask = __ballot_sync(0xffffffffu, row_ok);
  if (mask == 0u) return;
  const int slot = __popc(mask & ((1u << lane) - 1u)), nvalid = __popc(mask);
  const int m_first = __shfl_sync(0xffffffffu, m, __ffs(mask) - 1);
  if (row_ok) {
    const uint32_t row = slot_smem + (uint32_t)slot * (uint32_t)RS;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + 4 * i);
      float o[4] = {v[4 * i] + b4.x, v[4 * i + 1] + b4.y, v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w};
      if (p.relu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.0f);
      }
      if (p.alpha != nullptr) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.alpha + 4 * i));
        o[0] = o[0] > 0.f ? o[0] : o[0] * a4.x; o[1] = o[1] > 0.f ? o[1] : o[1] * a4.y;
        o[2] = o[2] > 0.f ? o[2] : o[2] * a4.z; o[3] = o[3] > 0.f ? o[3] : o[3] * a4.w;
      }
      sts128(row + (uint32_t)(i * 16), make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])));
    }
  }
  __syncwarp();
  uint4* g = reinterpret_cast<uint4*>(p.out_f32 + (size_t)m_first * NC);
  for (int i = lane; i < nvalid * CH; i += 32) {
    const int r = i / CH, ch = i - r * CH;
    g[i] = lds128(slot_smem + (uint32_t)r * (uint32_t)RS + (uint32_t)(ch * 16));
  }
  __syncwarp();
}

template <bool F16, int MPS>   // MPS = tcgen05.mma instructions per K step = channels per plane / 16
__global__ void __launch_bounds__(SV_THREADS, 1)
sv_conv_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const SvParams q) {
  extern __shared__ uint8_t smem_raw[];
  const ConvParams& p = q.p;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;                                            // 2 band buffers
  const uint32_t smem_b = smem_a + (uint32_t)q.a_bufs * (uint32_t)q.a_buf_bytes;   // resident weights or ring
  const uint32_t b_bytes = (uint32_t)(q.b_resident ? q.ksteps : q.stages) * (uint32_t)q.b_tile_bytes;
  const uint32_t smem_bias = smem_b + b_bytes;                                  // 256 floats
  const uint32_t smem_koff = smem_bias + 1024u;                                 // 128 x u32 shifted-view offsets (>> 4)
  const uint32_t bars = smem_koff + 512u;
  // barriers: afull[2], aempty[2], bres, full[S], empty[S], tfull[nbuf], tempty[nbuf]
  const uint32_t bar_afull = bars, bar_aempty = bars + 16u, bar_bres = bars + 32u, bar_full = bars + 40u,
                 bar_empty = bar_full + 8u * q.stages, bar_tfull = bar_empty + 8u * q.stages,
                 bar_tempty = bar_tfull + 8u * q.nbuf, tmem_slot = bar_tempty + 8u * q.nbuf;
  const uint32_t smem_stage = (tmem_slot + 16u + 127u) & ~127u;                 // 8 epilogue-warp slots of q.epi_stage_bytes
  float* s_bias = reinterpret_cast<float*>(smem_raw + (smem_bias - smem_u32(smem_raw)));
  uint32_t* s_koff = reinterpret_cast<uint32_t*>(smem_raw + (smem_koff - smem_u32(smem_raw)));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // data-dependent batch: the number of valid images may live on the device (detector candidate counts)
  const int n_img_live = q.n_img_dev != nullptr ? min(p.n_img, __ldg(q.n_img_dev)) : p.n_img;
  const int n_bands_live = q.bands_y * ((n_img_live + q.nb - 1) / q.nb);
  long long* dbg = blockIdx.x == 0 ? g_sv_dbg : nullptr;
  if (lane != 0) dbg = nullptr;
  const long long t_kernel = SV_T0();
  const uint32_t row_bytes = 2u * (uint32_t)q.ck;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar_afull + 8u * i, 1); mbar_init(bar_aempty + 8u * i, 1); }
    mbar_init(bar_bres, 1);
    for (int s = 0; s < q.stages; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    for (int i = 0; i < q.nbuf; ++i) {
      mbar_init(bar_tfull + 8u * i, 1);
      mbar_init(bar_tempty + 8u * i, q.epi_split ? SV_EPI_THREADS / 2 : SV_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
  if (warp == 10 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(q.nbuf * q.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias is the same for every tile (single N tile): stage it once
  for (int i = tid; i < 256; i += SV_THREADS) s_bias[i] = i < p.cout ? __ldg(p.bias + i) : 0.f;
  pdl_launch_dependents();       // the next kernel may start its prologue on SMs this grid has left
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                    // nothing above read or wrote global data: now wait for the producer of our inputs
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 8) {
    // ================================================= epilogue
    const int qq = warp & 3;
    const int r = qq * 32 + lane;
    const int chalf = warp >> 2;
    int ab = 0, tseq = 0;
    uint32_t tph = 0;
    for (int band = blockIdx.x; band < n_bands_live; band += gridDim.x) {
      const int n0_img = (band / q.bands_y) * q.nb;
      const int y0 = (band % q.bands_y) * q.R;
      const int r_valid = min(q.R, p.out_h - y0);
      for (int t = 0; t < q.tiles_per_band; ++t) {
        // narrow tiles (cout <= 64): the two warp groups take alternate tiles (each warp then covers all columns of its
        // rows), which halves the per-tile wait -> load -> store -> arrive latency chain the MMA warp sees
        const bool mine = !q.epi_split || ((tseq & 1) == (warp >> 2));
        ++tseq;
        if (mine) {
          const int lin = t * 128 + r;
          const int il = lin / q.img_px;
          const int rem = lin - il * q.img_px;
          const int rr = rem / q.P;
          const int x = rem - rr * q.P;
          const int img = n0_img + il;
          const bool row_ok = il < q.nb && img < n_img_live && rr < r_valid && x < p.out_w;
          const int m = (img * p.out_h + y0 + rr) * p.out_w + x;
          const long long e0 = (dbg && warp == 0) ? clock64() : 0ll;
          mbar_wait(bar_tfull + 8u * ab, tph);
          const long long e1 = (dbg && warp == 0) ? clock64() : 0ll;
          tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(qq * 32) << 16) + (uint32_t)(ab * q.tmem_cols);
          if (q.epi_split && q.epi_stage_bytes > 0) {
            const uint32_t slot = smem_stage + (uint32_t)warp * (uint32_t)q.epi_stage_bytes;
            if (p.out_f32 != nullptr) epilogue_row_narrow_staged_f32<64>(p, s_bias, t_row, m, row_ok, slot, lane);
            else if (p.cout == 32) epilogue_row_narrow_staged<F16, 32>(p, s_bias, t_row, m, row_ok, slot, lane);
            else epilogue_row_narrow_staged<F16, 64>(p, s_bias, t_row, m, row_ok, slot, lane);
          } else if (q.epi_split) {
            if (p.cout == 32) epilogue_row_narrow<F16, 32>(p, s_bias, t_row, m, row_ok);
            else epilogue_row_narrow<F16, 64>(p, s_bias, t_row, m, row_ok);
          } else {
            epilogue_row<F16>(p, s_bias, t_row, m, row_ok, 0, p.cout, chalf);
          }
          tc_fence_before();
          mbar_arrive(bar_tempty + 8u * ab);
          if (dbg && warp == 0) { dbg[5] += e1 - e0; dbg[6] += clock64() - e1; }
        }
        if (++ab == q.nbuf) { ab = 0; tph ^= 1u; }
      }
    }
  } else if (warp == 8) {
    // ================================================= A band producer (TMA, one lane)
    if (lane == 0) {
      int j = 0;
      for (int band = blockIdx.x; band < n_bands_live; band += gridDim.x, ++j) {
        const int buf = q.a_bufs == 2 ? (j & 1) : 0;
        const int n0_img = (band / q.bands_y) * q.nb;
        const int y0 = (band % q.bands_y) * q.R;
        const long long p0 = SV_T0();
        mbar_wait(bar_aempty + 8u * buf, (uint32_t)(((q.a_bufs == 2 ? (j >> 1) : j) & 1) ^ 1));
        SV_ACC(7, p0);
        mbar_arrive_expect_tx(bar_afull + 8u * buf, (uint32_t)q.n_chunks * (uint32_t)q.a_box_bytes);
        for (int c = 0; c < q.n_chunks; ++c)
          tma_load_4d(smem_a + (uint32_t)buf * q.a_buf_bytes + (uint32_t)c * q.plane_bytes, &tmap_a, bar_afull + 8u * buf,
                      c * q.ck, -p.pad_w, y0 - p.pad_h, n0_img);
      }
    }
  } else if (warp == 10) {
    // ================================================= weight producer (TMA, one lane)
    if (lane == 0) {
      if (q.b_resident) {
        mbar_arrive_expect_tx(bar_bres, (uint32_t)q.ksteps * (uint32_t)q.b_tile_bytes);
        for (int ks = 0; ks < q.ksteps; ++ks)
          tma_load_2d(smem_b + (uint32_t)ks * q.b_tile_bytes, &tmap_w, bar_bres, ks * q.ck, 0);
      } else {
        int s = 0;
        uint32_t ph = 1;                         // producer starts on the "previous phase complete" parity
        for (int band = blockIdx.x; band < n_bands_live; band += gridDim.x) {
          for (int t = 0; t < q.tiles_per_band; t += q.mt) {
            for (int ks = 0; ks < q.ksteps; ++ks) {
              mbar_wait(bar_empty + 8u * s, ph);
              mbar_arrive_expect_tx(bar_full + 8u * s, (uint32_t)q.b_tile_bytes);
              tma_load_2d(smem_b + (uint32_t)s * q.b_tile_bytes, &tmap_w, bar_full + 8u * s, ks * q.ck, 0);
              if (++s == q.stages) { s = 0; ph ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // ================================================= MMA issuer.  The whole warp runs the loop convergently (every
    // index is a wrapping counter, every shifted-view offset a kernel-parameter table entry, so all of it lives in
    // uniform registers); only the tcgen05 instructions are predicated on one elected lane.
    const uint32_t idesc = make_idesc_f16(p.block_n, F16 ? 1 : 0);
    if (q.b_resident) { mbar_wait(bar_bres, 0); }
    const uint64_t b_desc0 = make_sw_desc(smem_b, row_bytes, 0);
    const uint32_t b_tile16 = (uint32_t)q.b_tile_bytes >> 4;
    const uint32_t tile16 = (128u * row_bytes) >> 4;
    int s = 0, ab = 0;                         // weight ring stage, TMEM buffer of the next tile
    uint32_t ph = 0, tph = 1;                  // their phase parities (consumer / "buffer drained")
    int j = 0;
    for (int band = blockIdx.x; band < n_bands_live; band += gridDim.x, ++j) {
      const int buf = q.a_bufs == 2 ? (j & 1) : 0;
      const long long m0 = SV_T0();
      mbar_wait(bar_afull + 8u * buf, (uint32_t)((q.a_bufs == 2 ? (j >> 1) : j) & 1));
      SV_ACC(1, m0);
      tc_fence_after();
      const uint64_t a_desc0 = make_sw_desc(smem_a + (uint32_t)buf * q.a_buf_bytes, row_bytes, 0);
      for (int t = 0; t < q.tiles_per_band; t += q.mt) {
        const int nt = min(q.mt, q.tiles_per_band - t);
        // accumulator buffers of this group: nt consecutive buffers starting at `ab` (mod nbuf).  Consecutive
        // tcgen05.mma into the SAME accumulator serialise on the tensor pipe's latency (measured ~235 cycles per
        // M128 x N32 x K16 instruction), so the issue order below rotates over the group's independent accumulators.
        uint32_t d[4];
        uint64_t a_t[4];
        int abu = ab;
        uint32_t tphu = tph;
        const long long m1 = SV_T0();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          d[u] = tmem_base + (uint32_t)(abu * q.tmem_cols);
          a_t[u] = a_desc0 + (uint64_t)((uint32_t)(t + u) * tile16);
          if (u < nt) {
            mbar_wait(bar_tempty + 8u * abu, tphu);
            if (++abu == q.nbuf) { abu = 0; tphu ^= 1u; }
          }
        }
        SV_ACC(2, m1);
        tc_fence_after();
        if (q.b_resident) {
          // resident weights: nothing to wait for inside the K loop, so the whole loop runs inside ONE elected region
          // (an elect + __syncwarp per K step costs more than the 2-4 MMAs it guards).  Issue order: for every K slice
          // rotate over the group's accumulators (independent instructions back to back).
          if (elect_one()) {
            if (q.order == 0) {
              for (int ks = 0; ks < q.ksteps; ++ks) {
                const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)ks * b_tile16);
                const uint64_t ko = (uint64_t)q.koff[ks];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (u < nt) {
#pragma unroll
                    for (int kk = 0; kk < MPS; ++kk)
                      umma_bf16(d[u], a_t[u] + ko + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (ks | kk) != 0);
                  }
                }
              }
            } else if (nt == 4) {
              for (int ks = 0; ks < q.ksteps; ++ks) {
                const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)ks * b_tile16);
                const uint64_t ko = (uint64_t)q.koff[ks];
#pragma unroll
                for (int kk = 0; kk < MPS; ++kk) {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    umma_bf16(d[u], a_t[u] + ko + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (ks | kk) != 0);
                }
              }
            } else {
              for (int ks = 0; ks < q.ksteps; ++ks) {
                const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)ks * b_tile16);
                const uint64_t ko = (uint64_t)q.koff[ks];
#pragma unroll
                for (int kk = 0; kk < MPS; ++kk) {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (u < nt) umma_bf16(d[u], a_t[u] + ko + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (ks | kk) != 0);
                }
              }
            }
          }
          __syncwarp();
        } else {
          for (int ks = 0; ks < q.ksteps; ++ks) {
            const long long m2 = SV_T0();
            mbar_wait(bar_full + 8u * s, ph);
            SV_ACC(3, m2);
            tc_fence_after();
            const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)s * b_tile16);
            const uint64_t ko = (uint64_t)q.koff[ks];
            if (elect_one()) {
              if (q.order == 0) {
                // N = 128: accumulator-major order measured faster (29.7 vs 37.0 us on Block17 1x7)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (u < nt) {
#pragma unroll
                    for (int kk = 0; kk < MPS; ++kk)
                      umma_bf16(d[u], a_t[u] + ko + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (ks | kk) != 0);
                  }
                }
              } else {
                // small N: back-to-back instructions into one accumulator serialise (126 cycles): rotate per K slice
#pragma unroll
                for (int kk = 0; kk < MPS; ++kk) {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (u < nt) umma_bf16(d[u], a_t[u] + ko + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (ks | kk) != 0);
                }
              }
              umma_commit(bar_empty + 8u * s);
            }
            __syncwarp();
            if (++s == q.stages) { s = 0; ph ^= 1u; }
          }
        }
        if (elect_one()) {
          int abc = ab;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (u < nt) { umma_commit(bar_tfull + 8u * abc); if (++abc == q.nbuf) abc = 0; }
        }
        __syncwarp();
        ab = abu; tph = tphu;
      }
      if (elect_one()) umma_commit(bar_aempty + 8u * buf);      // band buffer reusable once every MMA reading it completed
      __syncwarp();
    }
  }

  __syncthreads();
  if (dbg && tid == 0) dbg[0] += clock64() - t_kernel;
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(q.nbuf * q.tmem_cols)) : "memory");
  }
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Geometry of the shifted-view plan for one op; returns false when the op does not qualify / does not fit.
bool sv_plan(const VnfrConvOp* op, SvParams* q) {
  if (op->stride != 1 || (op->kh == 1 && op->kw == 1)) return false;
  if (op->cout > 256 || op->block_n != op->cout || op->cout_pad != op->cout) return false;
  const int ck = op->reserved[0];
  if (ck != 16 && ck != 32 && ck != 64) return false;
  const int row_bytes = 2 * ck;
  const int n_chunks = (op->cin + ck - 1) / ck;        // channel planes of the A band
  const int taps = op->kh * op->kw;
  // split precision: 1 = 3 bf16 parts / 6 products, 2 = 2 fp16 parts / 3 products
  if (op->split3 != 0 && op->split3 != 1 && op->split3 != 2) return false;
  const int n_parts = op->split3 == 2 ? 2 : 3;
  if (op->split3 && (n_chunks != n_parts || op->cin != n_parts * ck)) return false;
  if (op->split3 == 2 && op->dtype != 1) return false;   // the two-part split needs fp16's 11-bit significand
  const int kchunks = op->split3 == 2 ? 3 : (op->split3 ? 6 : n_chunks);        // K steps per tap
  if (op->k_pad < taps * kchunks * ck) return false;
  const int P = op->in_w + 2 * op->pad_w;
  if (P > 256) return false;
  const int b_tile = op->block_n * row_bytes;
  const int ksteps = taps * kchunks;
  if (ksteps > 128) return false;
  int tmem_cols = 32;
  while (tmem_cols < op->block_n) tmem_cols <<= 1;
  // narrow tiles whose destination rows are exactly the pixel (no channel slice of a wider buffer): staged stores
  const bool narrow = (op->cout == 32 || op->cout == 64) && op->residual == nullptr && (op->out_f32 != nullptr || op->n_split >= op->cout) &&
                      getenv("VNFR_SV_NO_EPI_SPLIT") == nullptr;
  // (64-wide rows only: for the 32-wide stem layers the staged path is slower even when the slots cost no band space --
  // conv2d_1a 141 -> 170 us, conv2d_2a 179 -> 187 us: 64-byte rows already share their 128-byte lines pairwise; conv2d_2b,
  // 64 wide: 245 -> 217 us)
  int stage_bytes = (narrow && op->cout == 64 && op->out_f32 == nullptr && op->out0 != nullptr && op->out0_pitch == op->cout && op->prelu_alpha == nullptr &&
                     ((uintptr_t)op->out0 % 16) == 0 && getenv("VNFR_SV_NO_STAGE") == nullptr) ? 32 * op->cout * 2 : 0;
  // fp32 destinations: O-Net conv2 only (64 floats per row, 32-channel planes).  Measured per half batch (ncu): O-Net conv2
  // 246 -> 208 us; O-Net conv3 (64-channel planes: the slots cost it band space) 83 -> 101 us and R-Net conv2 (48 floats per
  // row, which would also have to leave its two-warps-per-row epilogue) 158 -> 169 us, so those two keep the direct stores
  if (narrow && op->cout == 64 && ck == 32 && op->out_f32 != nullptr && op->out_f32_pitch == op->cout && ((uintptr_t)op->out_f32 % 16) == 0 &&
      (op->prelu_alpha == nullptr || ((uintptr_t)op->prelu_alpha % 16) == 0) && getenv("VNFR_SV_NO_STAGE") == nullptr)
    stage_bytes = 32 * (op->cout / 4 + 1) * 16;      // fp32 rows, padded by one 16-byte chunk
  const int budget = 212 * 1024 - 8 * stage_bytes;
  const bool b_res = (long long)ksteps * b_tile <= 80 * 1024;
  // accumulator tiles in flight per group: as many independent accumulators as TMEM holds twice over (max 4)
  int mt = 512 / (2 * tmem_cols);
  if (mt > 4) mt = 4;
  if (mt < 1) mt = 1;
  if (getenv("VNFR_SV_MT") != nullptr && atoi(getenv("VNFR_SV_MT")) >= 1 && atoi(getenv("VNFR_SV_MT")) < mt) mt = atoi(getenv("VNFR_SV_MT"));
  // streamed weights are only worth it when two accumulator tiles share every weight stage (otherwise the generic
  // kernel moves fewer bytes per useful output)
  if (!b_res && mt < 2) return false;
  // weight ring: deep enough to cover the TMA latency with small tiles (4 KB tiles of the split-precision O-Net conv2
  // need ~16 in flight), 64 KB at most
  int stages = 1;
  if (!b_res) {
    stages = (64 * 1024) / b_tile;
    if (stages < 4) stages = 4;
    if (stages > 16) stages = 16;
  }
  const int b_bytes = b_res ? ksteps * b_tile : stages * b_tile;
  const int px_align = 1024 / row_bytes;
  // candidates: (R, nb, band buffers); maximise useful accumulator rows per issued row, prefer fewer halo re-reads and
  // full weight-stage sharing; a single band buffer (no load/MMA overlap between bands) is allowed at a 10 % discount
  double best = -1;
  int bR = 0, bnb = 0, bbufs = 2;
  for (int bufs = 2; bufs >= 1; --bufs) {
    const int a_budget = (budget - b_bytes - 2048) / bufs;    // per band buffer
    for (int R = 1; R <= op->out_h; ++R) {
      const int rows_in = R + op->kh - 1;
      if (rows_in > 256) break;
      const int img_px = rows_in * P;
      const int nb_max = R == op->out_h ? 16 : 1;
      for (int nb = 1; nb <= nb_max && nb <= op->n_img; ++nb) {
        const int lin = (nb - 1) * img_px + R * P;
        const int tiles = (lin + 127) / 128;
        const int alloc_px = round_up(tiles * 128 + (op->kh - 1) * P + op->kw - 1, px_align);
        if (alloc_px < nb * img_px) continue;
        const long long a_bytes = (long long)n_chunks * alloc_px * row_bytes;
        if (a_bytes > a_budget) continue;
        if ((long long)nb * img_px * row_bytes > 200 * 1024) continue;
        const int bands_y = (op->out_h + R - 1) / R;
        const double valid = (double)op->out_h * op->out_w * nb;
        const double issued = (double)bands_y * tiles * 128;
        const double halo = (double)(R + op->kh - 1) / (R + 0.25 * (op->kh - 1));
        const double share = b_res ? 1.0 : (double)tiles / (double)(mt * ((tiles + mt - 1) / mt));   // weight-stage sharing
        const double score = valid / issued / halo * share * (bufs == 2 ? 1.0 : 0.9);
        if (score > best) { best = score; bR = R; bnb = nb; bbufs = bufs; }
      }
    }
  }
  if (best < 0) return false;
  const int R = bR, nb = bnb;
  q->ck = ck; q->n_chunks = n_chunks; q->taps = taps; q->ksteps = ksteps; q->P = P; q->R = R;
  q->rows_in = R + op->kh - 1; q->nb = nb; q->img_px = q->rows_in * P;
  q->tiles_per_band = ((nb - 1) * q->img_px + R * P + 127) / 128;
  q->bands_y = (op->out_h + R - 1) / R;
  q->n_bands = q->bands_y * ((op->n_img + nb - 1) / nb);
  const int alloc_px = round_up(q->tiles_per_band * 128 + (op->kh - 1) * P + op->kw - 1, px_align);
  q->plane_bytes = alloc_px * row_bytes;
  q->a_buf_bytes = n_chunks * q->plane_bytes;
  q->a_box_bytes = nb * q->img_px * row_bytes;
  q->b_resident = b_res ? 1 : 0;
  q->b_tile_bytes = b_tile;
  q->stages = stages;
  q->mt = mt; q->nbuf = 2 * mt; q->tmem_cols = tmem_cols;
  // Measured on B200: the swizzle XOR is a pure function of the shared-memory ADDRESS bits, so a start address shifted by
  // any number of rows (also odd multiples of 64 bytes in 64B-swizzle mode) reads correctly with base_offset = 0; setting
  // the field to (addr >> 7) & 7 applies the phase twice (tests/test_gpu_encoder.py::test_shifted_view_conv_matches_torch).
  q->use_base_offset = 0;
  // activation part of product j: 3 bf16 parts (weights: 0,1,0,2,1,0) / 2 fp16 parts (weights: 0,1,0)
  static const int split_plane3[6] = {0, 0, 1, 0, 1, 2}, split_plane2[3] = {0, 0, 1};
  const int* split_plane = op->split3 == 2 ? split_plane2 : split_plane3;
  for (int ks = 0; ks < ksteps; ++ks) {
    const int tap = ks / kchunks, j = ks - tap * kchunks;
    const int c = op->split3 ? split_plane[j] : j;
    const int ky = tap / op->kw, kx = tap - ky * op->kw;
    q->koff[ks] = ((uint32_t)c * (uint32_t)q->plane_bytes + (uint32_t)(ky * P + kx) * (uint32_t)row_bytes) >> 4;
  }
  q->a_bufs = bbufs;
  // narrow single-destination tiles without residual: alternate tiles between the two epilogue warp groups
  q->epi_split = narrow ? 1 : 0;
  q->epi_stage_bytes = stage_bytes;
  q->n_img_dev = op->n_img_dev;
  q->order = 0;      // accumulator-major measured at least as fast as rotating per K slice in every layer (profiles/)
  if (getenv("VNFR_SV_ORDER") != nullptr) q->order = atoi(getenv("VNFR_SV_ORDER"));
  return true;
}

size_t sv_smem_bytes(const SvParams& q) {
  const size_t b_bytes = (size_t)(q.b_resident ? q.ksteps : q.stages) * q.b_tile_bytes;
  return 1024 + (size_t)q.a_bufs * q.a_buf_bytes + b_bytes + 1024 + 512 + 40 + 16 * q.stages + 16 * q.nbuf + 16 + 128 + 8 * (size_t)q.epi_stage_bytes;
}

}  // namespace

extern long long g_vnfr_launches;

// debug hook (not part of include/vnfr_b200.h): device buffer of 8 int64 cycle counters, or null to switch off
extern "C" int vnfr_sv_debug(long long* dev_buf) {
  VNFR_CUDA(cudaMemcpyToSymbol(g_sv_dbg, &dev_buf, sizeof(dev_buf)));
  return VNFR_OK;
}

int vnfr_sv_prepare(VnfrConvOp* op) {
  SvParams q;
  memset(&q, 0, sizeof(q));
  if (!sv_plan(op, &q)) return VNFR_ERR_UNSUPPORTED;
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) return VNFR_ERR_CUDA;
  const CUtensorMapDataType dt = op->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = q.ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (q.ck == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  // weights: [cout][k_pad], box {ck, block_n}
  {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)op->k_pad, (cuuint64_t)op->cout_pad};
    const cuuint64_t strides[1] = {(cuuint64_t)op->k_pad * 2};
    const cuuint32_t box[2] = {(cuuint32_t)q.ck, (cuuint32_t)op->block_n};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tm, dt, 2, const_cast<void*>(op->weights), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VNFR_ERR_CUDA;
    memcpy(op->tmap_w, &tm, sizeof(tm));
  }
  // activations: {C, W, H, N}, box {ck, P, rows_in, nb}; out-of-bounds (padding, channel tail, image tail) reads as zero
  {
    CUtensorMap ta;
    const cuuint64_t dims[4] = {(cuuint64_t)op->cin, (cuuint64_t)op->in_w, (cuuint64_t)op->in_h, (cuuint64_t)op->n_img};
    const cuuint64_t strides[3] = {(cuuint64_t)op->in_pitch * 2, (cuuint64_t)op->in_w * op->in_pitch * 2,
                                   (cuuint64_t)op->in_h * op->in_w * op->in_pitch * 2};
    const cuuint32_t box[4] = {(cuuint32_t)q.ck, (cuuint32_t)q.P, (cuuint32_t)q.rows_in, (cuuint32_t)q.nb};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&ta, dt, 4, const_cast<void*>(op->in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VNFR_ERR_CUDA;
    memcpy(op->tmap_a, &ta, sizeof(ta));
  }
  op->a_mode = 3;
  return VNFR_OK;
}

int vnfr_sv_run(const VnfrConvOp* op, void* stream) {
  SvParams q;
  memset(&q, 0, sizeof(q));
  if (!sv_plan(op, &q)) {
    vnfr_set_error(__FILE__, __LINE__, "op prepared for the shifted-view kernel no longer qualifies");
    return VNFR_ERR_ARG;
  }
  static VnfrPerDevice attr_set_once = {};
  if (vnfr_first_on_device(attr_set_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(sv_conv_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(sv_conv_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(sv_conv_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(sv_conv_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(sv_conv_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(sv_conv_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  ConvParams& p = q.p;
  p.in = (const __nv_bfloat16*)op->in;
  p.bias = op->bias;
  p.residual = (const __nv_bfloat16*)op->residual;
  p.out0 = (__nv_bfloat16*)op->out0;
  p.out1 = (__nv_bfloat16*)op->out1;
  p.out_f32 = op->out_f32;
  p.alpha = op->prelu_alpha;
  p.n_img = op->n_img; p.in_h = op->in_h; p.in_w = op->in_w; p.cin = op->cin; p.in_pitch = op->in_pitch;
  p.kh = op->kh; p.kw = op->kw; p.stride = 1; p.pad_h = op->pad_h; p.pad_w = op->pad_w;
  p.out_h = op->out_h; p.out_w = op->out_w;
  p.M = op->n_img * op->out_h * op->out_w;
  p.cout = op->cout; p.block_n = op->block_n;
  p.n_split = op->n_split; p.out0_pitch = op->out0_pitch; p.out1_pitch = op->out1_pitch;
  p.res_pitch = op->res_pitch; p.out_f32_pitch = op->out_f32_pitch;
  p.relu = op->relu; p.dtype = op->dtype; p.a_mode = 3;
  p.tmem_cols = q.tmem_cols; p.stages = q.stages;
  if (p.M <= 0) return VNFR_OK;
  const size_t smem = sv_smem_bytes(q);
  if (smem > 227 * 1024) {
    vnfr_set_error(__FILE__, __LINE__, "shifted-view plan exceeds shared memory");
    return VNFR_ERR_ARG;
  }
  CUtensorMap tm, ta;
  memcpy(&tm, op->tmap_w, sizeof(tm));
  memcpy(&ta, op->tmap_a, sizeof(ta));
  const int grid = q.n_bands < 148 ? q.n_bands : 148;
  cudaStream_t st = (cudaStream_t)stream;
  if (q.ck == 16) {
    if (op->dtype == 1) VNFR_CUDA(launch_pdl(sv_conv_kernel<true, 1>, grid, SV_THREADS, smem, st, tm, ta, q));
    else VNFR_CUDA(launch_pdl(sv_conv_kernel<false, 1>, grid, SV_THREADS, smem, st, tm, ta, q));
  } else if (q.ck == 32) {
    if (op->dtype == 1) VNFR_CUDA(launch_pdl(sv_conv_kernel<true, 2>, grid, SV_THREADS, smem, st, tm, ta, q));
    else VNFR_CUDA(launch_pdl(sv_conv_kernel<false, 2>, grid, SV_THREADS, smem, st, tm, ta, q));
  } else {
    if (op->dtype == 1) VNFR_CUDA(launch_pdl(sv_conv_kernel<true, 4>, grid, SV_THREADS, smem, st, tm, ta, q));
    else VNFR_CUDA(launch_pdl(sv_conv_kernel<false, 4>, grid, SV_THREADS, smem, st, tm, ta, q));
  }
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
