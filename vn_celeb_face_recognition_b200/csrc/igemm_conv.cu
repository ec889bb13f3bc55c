// Implicit-GEMM convolution for the InceptionResnetV1 encoder and the MLP head on 5th-gen tensor cores (sm_100a).
//
//   D[m][n] = sum_k A[m][k] * W[n][k]       m = (img, oy, ox)   n = output channel   k = (ky, kx, c)
//
//   * A (activations, NHWC bf16) is gathered tap by tap straight from HBM/L2 into 128B-swizzled shared memory by
//     128 producer threads (one GEMM row each) with 16-byte cp.async (zero-fill for padding / K tail),
//   * W (bf16 [cout_pad][k_pad], BN scale folded in) arrives through TMA (cp.async.bulk.tensor.2d, 128B swizzle),
//   * one elected thread issues tcgen05.mma (M=128, N=block_n, K=16 per instruction), accumulating fp32 in TMEM,
//   * the producer warps then become the epilogue: tcgen05.ld -> +bias (+residual) (ReLU) -> bf16 -> channel slice
//     of the destination (this is how torch.cat and `out*scale + x` of the reference disappear).
//
// Replaces the cuDNN / cuBLAS call sites of inception_resnet_v1.py:12-33, :56-67, :85-95, :114-126, :296-297 and
// mlp_model.py:10-15.  One CTA = one 128 x block_n output tile; ~96 KB of shared memory and
// <= 256 TMEM columns per CTA so that two CTAs share an SM and one's epilogue overlaps the other's main loop.
#include "tc_common.cuh"

using namespace tc;

namespace {

// ---------------------------------------------------------------------------------------------------------- kernel
// Persistent, warp-specialised: one CTA per SM walks over output tiles (tile = 128 rows x block_n channels).
//   warps 0-7   A producers: two threads per GEMM row (= output pixel), 4 x 16-byte cp.async each per K block; the
//               full barrier is armed by cp.async.mbarrier.arrive (no wait_group in the loop: fully asynchronous)
//   warps 8-15  epilogue: TMEM lane quarter = warp & 3, two warps per quarter take alternate 16-column chunks; they
//               drain accumulator buffer `ab` while the MMA fills the other one
//   warp  16    TMA producer: W always, A too when the convolution is 1x1 (plain 2-D box) (one elected lane)
//   warp  17    TMEM alloc + tcgen05.mma issue (one elected lane)
// Three barrier rings: smem full/empty per stage (global K-block counter runs across tiles, so the load pipeline never
// drains at a tile boundary), TMEM full/empty per accumulator buffer.
// Shared memory: [A stage 0..S) 16 KB each][W stage 0..S) block_n*128 B each][bias 2 x 256 fp32][barriers].
// Optional cycle breakdown of CTA 0 (tools/ig_probe.py): [0] kernel, [1] MMA warp waiting for a full stage, [2] for a
// drained accumulator, [3] epilogue waiting for the accumulator, [4] for the residual / free C buffer, [5] epilogue work,
// [6] store issue + wait, [7] producer waiting for an empty stage, [8] producer waiting for a free C buffer.
__device__ long long* g_ig_dbg = nullptr;
#define IG_T0() (dbg ? clock64() : 0ll)
#define IG_ACC(i, t0) do { if (dbg) dbg[i] += clock64() - (t0); } while (0)

__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_r, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const int KB0 = p.k_blocks;
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + (uint32_t)S * A_STAGE_BYTES;
  const uint32_t smem_c = smem_b + (uint32_t)(p.b_res ? KB0 : S) * b_stage_bytes;   // staged C tile: 64-channel panels of 16 KB
  const uint32_t n_panels = p.epi_mode ? (uint32_t)((p.block_n + 63) >> 6) : 0u;
  const uint32_t c_buf_bytes = n_panels * 16384u;
  const uint32_t smem_bias = smem_c + (uint32_t)p.c_bufs * c_buf_bytes;     // 2 x 256 floats
  const uint32_t bars = smem_bias + 2048u;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * S, bar_tfull = bars + 16u * S, bar_tempty = bar_tfull + 16u,
                 bar_res = bar_tempty + 16u, bar_cfree = bar_res + 16u, bar_bres = bar_cfree + 16u, tmem_slot = bar_bres + 8u,
                 bar_cready = bar_bres + 16u;      // [2]: all epilogue threads have written their rows of C buffer cb
  // TMA-fed A (1x1 convs): warps 0-7 have no gather work, so thread 0 becomes a dedicated C-store thread -- the epilogue
  // warps then never wait for a store to drain and need no CTA-level barrier before it (ncu: "barrier" was the top stall
  // of the projection convs: 5-10 stalled warps per issue)
  const bool store_thread = p.epi_mode && p.a_mode != 0;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (smem_bias - smem_u32(smem_raw)));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  long long* dbg = (blockIdx.x == 0 && lane == 0) ? g_ig_dbg : nullptr;
  const long long t_kernel = IG_T0();
  const int KB = p.k_blocks;
  const int n_tiles_n = p.n_tiles_n;
  const int total_tiles = p.n_tiles_m * n_tiles_n;
  // Tile walk.  Default: tile = blockIdx.x, += gridDim.x.  Resident-weights mode (b_res): the CTA is bound to ONE N tile
  // (blockIdx.x % n_tiles_n) whose K x block_n weight slab stays in shared memory for the whole kernel, and walks the M
  // tiles of that column -- for the small-K projection convolutions the weight slab was 40 % of a tile's L2 traffic.
  const int tile_first = p.b_res ? (blockIdx.x / n_tiles_n) * n_tiles_n + (blockIdx.x % n_tiles_n) : (int)blockIdx.x;
  const int tile_step = p.b_res ? ((int)gridDim.x / n_tiles_n) * n_tiles_n : (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8u * s, p.a_mode == 0 ? NUM_PRODUCER_THREADS + 1 : 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8u * i, 1);
      mbar_init(bar_tempty + 8u * i, NUM_EPILOGUE_THREADS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_res + 8u * i, 1); mbar_init(bar_cfree + 8u * i, 1); mbar_init(bar_cready + 8u * i, NUM_EPILOGUE_THREADS);
    }
    mbar_init(bar_bres, 1);
    fence_barrier_init();
  }
  if (warp == 16 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    if (p.a_mode != 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    if (p.epi_mode) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    if (p.epi_mode && p.residual != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_r) : "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();       // the next kernel may start its prologue on SMs this grid has left
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above touched no global data: now wait for the producer of our inputs (the TMA producer warp waits after it
  // has put the resident weight slab in flight: weights do not depend on the previous kernel)
  if (warp != 16) pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 8) {
    if (p.a_mode == 0) {
    // ================================================= A producers
    // 8 consecutive lanes copy the 8 16-byte chunks (one 128-byte K-block row) of one GEMM row, so every warp-level
    // cp.async touches whole 128-byte lines; each thread serves 4 rows (rg, rg+32, rg+64, rg+96) with the same chunk j,
    // hence the (tap, channel) decode of chunk j is done once per K block per thread.
    const int j = tid & 7;                     // chunk inside the 128-byte row
    const int rg = tid >> 3;                   // 0..31
    const int hw = p.out_h * p.out_w;
    int s = 0;                                 // ring stage / phase parity: wrapping counters, no division in the loop
    uint32_t ph = 1;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int m_base = (tile / n_tiles_n) * BLOCK_M;
      int iy0[4], ix0[4];
      const __nv_bfloat16* img_base[4];
      uint32_t dst_off[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = rg + 32 * q;
        const int m = m_base + r;
        dst_off[q] = (uint32_t)r * 128u + (((uint32_t)j ^ (uint32_t)(r & 7)) << 4);
        if (m < p.M) {
          const int img = m / hw;
          const int rem = m - img * hw;
          const int oy = rem / p.out_w;
          const int ox = rem - oy * p.out_w;
          iy0[q] = oy * p.stride - p.pad_h;
          ix0[q] = ox * p.stride - p.pad_w;
          img_base[q] = p.in + (size_t)img * p.in_h * p.in_w * p.in_pitch;
        } else {
          iy0[q] = -(1 << 28);                 // fails every bounds check -> zero fill
          ix0[q] = 0;
          img_base[q] = p.in;
        }
      }
      // (tap, channel) of chunk j in K block 0, then advanced by 64 elements per block
      int kf = j * 8;
      int c = kf, ky = 0, kx = 0;
      while (c >= p.cin) { c -= p.cin; if (++kx == p.kw) { kx = 0; ++ky; } }
      for (int kb = 0; kb < KB; ++kb) {
        const bool k_ok = kf < p.K;
        mbar_wait(bar_empty + 8u * s, ph);
        const uint32_t stage = smem_a + (uint32_t)s * A_STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int iy = iy0[q] + ky, ix = ix0[q] + kx;
          const bool ok = k_ok && (unsigned)iy < (unsigned)p.in_h && (unsigned)ix < (unsigned)p.in_w;
          const void* src = ok ? (const void*)(img_base[q] + ((size_t)iy * p.in_w + ix) * p.in_pitch + c) : (const void*)p.in;
          cp_async_16(stage + dst_off[q], src, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(bar_full + 8u * s);
        if (++s == S) { s = 0; ph ^= 1u; }
        kf += 64;
        c += 64;
        while (c >= p.cin) { c -= p.cin; if (++kx == p.kw) { kx = 0; ++ky; } }
      }
    }
    cp_async_wait<0>();                        // nothing may be in flight when the CTA retires
    }  // a_mode == 0
    else if (store_thread && tid == 0) {
      // ================================================= dedicated C-store thread (TMA-fed A: these warps are otherwise idle)
      int tcount = 0;
      for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tcount) {
        const int cb = p.c_bufs == 2 ? (tcount & 1) : 0;
        const uint32_t rph = p.c_bufs == 2 ? (uint32_t)((tcount >> 1) & 1) : (uint32_t)(tcount & 1);
        const int n0 = (tile % n_tiles_n) * p.block_n, m0 = (tile / n_tiles_n) * BLOCK_M;
        const int n_valid = min(p.block_n, p.cout - n0);
        const uint32_t smem_cb = smem_c + (uint32_t)cb * c_buf_bytes;
        mbar_wait(bar_cready + 8u * cb, rph);
        for (int j = 0; j * 64 < n_valid; ++j) tma_store_2d(&tmap_c, smem_cb + (uint32_t)j * 16384u, n0 + j * 64, m0);
        tma_store_commit();
        tma_store_wait_read();                 // the panels have been read: hand the buffer back to the producer
        mbar_arrive(bar_cfree + 8u * cb);
      }
      tma_store_wait_all();
    }
  } else if (warp < 16) {
    // ================================================= epilogue: TMEM -> registers -> bias/residual/ReLU -> global
    const int q = warp & 3;                    // TMEM lane quarter of this warp
    const int r = q * 32 + lane;               // row inside the tile
    const int et = tid - NUM_PRODUCER_THREADS; // 0..255
    const int chalf = (warp - 8) >> 2;         // this warp handles 16-column chunks with (chunk & 1) == chalf
    int tcount = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tcount) {
      const int ab = tcount & 1;
      const int m = (tile / n_tiles_n) * BLOCK_M + r;
      const int n0 = (tile % n_tiles_n) * p.block_n;
      const bool row_ok = m < p.M;
      const int n_valid = min(p.block_n, p.cout - n0);
      // stage this tile's bias while the MMAs are still running
      float* sb = s_bias + ab * 256;
      for (int i = et; i < n_valid; i += NUM_EPILOGUE_THREADS) sb[i] = __ldg(p.bias + n0 + i);
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPILOGUE_THREADS) : "memory");      // bias visible to the 8 epilogue warps
      long long* edbg = warp == 8 ? dbg : nullptr;
      const long long e0 = edbg ? clock64() : 0ll;
      mbar_wait(bar_tfull + 8u * ab, (tcount >> 1) & 1);
      if (edbg) edbg[3] += clock64() - e0;
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.tmem_cols);
      if (p.epi_mode) {
        // staged: C buffer `cb` is free (and holds the residual, if any) once its bar_res completes for this tile
        const int cb = p.c_bufs == 2 ? (tcount & 1) : 0;
        const uint32_t cph = p.c_bufs == 2 ? (uint32_t)((tcount >> 1) & 1) : (uint32_t)(tcount & 1);
        const uint32_t smem_cb = smem_c + (uint32_t)cb * c_buf_bytes;
        if (store_thread) {
          mbar_wait(bar_res + 8u * cb, cph);
          epilogue_row_staged<F16>(p, sb, t_row, smem_cb, r, n_valid, chalf, p.residual != nullptr);
          tc_fence_before();
          mbar_arrive(bar_tempty + 8u * ab);     // accumulator drained: the MMA warp may reuse it
          fence_proxy_async_smem();              // generic-proxy writes of this row -> visible to the TMA store
          mbar_arrive(bar_cready + 8u * cb);
          continue;
        }
        if (p.c_bufs == 2 && et == 0 && tcount > 0) {
          // the store of the previous tile (other buffer) was issued a whole tile ago: once it has read its panels the
          // buffer goes back to the producer, which refills it with the NEXT tile's residual while this tile is processed
          tma_store_wait_read();
          mbar_arrive(bar_cfree + 8u * (cb ^ 1));
        }
        const long long e1 = edbg ? clock64() : 0ll;
        mbar_wait(bar_res + 8u * cb, cph);
        const long long e2 = edbg ? clock64() : 0ll;
        epilogue_row_staged<F16>(p, sb, t_row, smem_cb, r, n_valid, chalf, p.residual != nullptr);
        tc_fence_before();
        mbar_arrive(bar_tempty + 8u * ab);     // accumulator drained: the MMA warp may reuse it
        fence_proxy_async_smem();              // generic-proxy writes of the tile -> visible to the TMA store
        asm volatile("bar.sync 2, %0;" ::"n"(NUM_EPILOGUE_THREADS) : "memory");
        const long long e3 = edbg ? clock64() : 0ll;
        if (edbg) { edbg[4] += e2 - e1; edbg[5] += e3 - e2; }
        if (et == 0) {
          const int m0 = (tile / n_tiles_n) * BLOCK_M;
          for (int j = 0; j * 64 < n_valid; ++j) tma_store_2d(&tmap_c, smem_cb + (uint32_t)j * 16384u, n0 + j * 64, m0);
          tma_store_commit();
          if (p.c_bufs != 2) {
            tma_store_wait_read();             // the panels have been read: they may be refilled
            mbar_arrive(bar_cfree);
          }
          if (edbg) edbg[6] += clock64() - e3;
        }
        continue;
      }
      epilogue_row<F16>(p, sb, t_row, m, row_ok, n0, n_valid, chalf);
      tc_fence_before();
      mbar_arrive(bar_tempty + 8u * ab);       // accumulator buffer `ab` may be overwritten
    }
  } else if (warp == 16) {
    // ================================================= W producer: TMA, one elected lane
    if (lane == 0) {
      int s = 0, tn = 0;
      uint32_t ph = 1;
      if (p.b_res && tile_first < total_tiles) {
        // the weight slab of this CTA's N tile: loaded once
        const int n0 = (tile_first % n_tiles_n) * p.block_n;
        mbar_arrive_expect_tx(bar_bres, (uint32_t)KB * b_stage_bytes);
        for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem_b + (uint32_t)kb * b_stage_bytes, &tmap_w, bar_bres, kb * BLOCK_K, n0);
      }
      pdl_wait();
      for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
        const int n0 = (tile % n_tiles_n) * p.block_n;
        const int m0 = (tile / n_tiles_n) * BLOCK_M;
        // base pixel of the tile for the im2col loads: input coordinates of output pixel m0 at tap (0, 0)
        const int chunks = p.cin >> 6;
        const int bn = m0 / (p.out_h * p.out_w), brem = m0 - bn * (p.out_h * p.out_w);
        const int bh = (brem / p.out_w) * p.stride - p.pad_h, bw = (brem % p.out_w) * p.stride - p.pad_w;
        for (int kb = 0; kb < KB; ++kb) {
          const long long p0 = IG_T0();
          mbar_wait(bar_empty + 8u * s, ph);
          IG_ACC(7, p0);
          if (p.a_mode == 1) {
            // 1x1 convolution: the A tile is a plain 2-D box of the [M][pitch] activation matrix
            mbar_arrive_expect_tx(bar_full + 8u * s, (p.b_res ? 0u : b_stage_bytes) + A_STAGE_BYTES);
            tma_load_2d(smem_a + (uint32_t)s * A_STAGE_BYTES, &tmap_a, bar_full + 8u * s, kb * BLOCK_K, m0);
          } else if (p.a_mode == 2) {
            // k x k convolution, cin a multiple of 64: K block kb = (tap, 64-channel chunk) is ONE im2col TMA load of the 128
            // output pixels that follow the tile's first one in (img, oy, ox) order (the unit walks the rows / images itself
            // and zero-fills the padding).  The cp.async gather this replaces moved ~13 B/clk/SM.
            const int tap = kb / chunks, chunk = kb - tap * chunks;
            const int ky = tap / p.kw, kx = tap - ky * p.kw;
            mbar_arrive_expect_tx(bar_full + 8u * s, (p.b_res ? 0u : b_stage_bytes) + A_STAGE_BYTES);
            tma_load_im2col_4d(smem_a + (uint32_t)s * A_STAGE_BYTES, &tmap_a, bar_full + 8u * s, chunk * 64, bw, bh, bn,
                               (uint16_t)kx, (uint16_t)ky);
          } else {
            mbar_arrive_expect_tx(bar_full + 8u * s, b_stage_bytes);
          }
          if (!p.b_res) tma_load_2d(smem_b + (uint32_t)s * b_stage_bytes, &tmap_w, bar_full + 8u * s, kb * BLOCK_K, n0);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (p.epi_mode) {
          // C panels: wait until the previous tile's TMA store has read them, then (re)fill them with this tile's
          // residual.  Issued AFTER the tile's K-block loads so that it never holds up the MMA pipeline; the ring depth
          // gives the load its lead time over the epilogue.
          const int cb = p.c_bufs == 2 ? (tn & 1) : 0;
          const uint32_t cph = p.c_bufs == 2 ? (uint32_t)(((tn >> 1) & 1) ^ 1) : (uint32_t)((tn & 1) ^ 1);
          const uint32_t smem_cb = smem_c + (uint32_t)cb * c_buf_bytes;
          const long long p1 = IG_T0();
          mbar_wait(bar_cfree + 8u * cb, cph);
          IG_ACC(8, p1);
          if (p.residual != nullptr) {
            const int n_valid = min(p.block_n, p.cout - n0);
            const int np = (n_valid + 63) >> 6;
            mbar_arrive_expect_tx(bar_res + 8u * cb, (uint32_t)np * 16384u);
            for (int j = 0; j < np; ++j) tma_load_2d(smem_cb + (uint32_t)j * 16384u, &tmap_r, bar_res + 8u * cb, n0 + j * 64, m0);
          } else {
            mbar_arrive(bar_res + 8u * cb);
          }
          ++tn;
        }
      }
    }
  } else {
    // ================================================= MMA issuer.  The whole warp runs the loop convergently so that
    // stage / phase counters and descriptors live in uniform registers (a lane-0-only loop costs ~25 scalar
    // instructions + a register->uniform waterfall per tcgen05.mma, measured ~300 cycles per instruction); only the
    // tcgen05 instructions themselves are predicated on one elected lane.
    const uint32_t idesc = make_idesc_f16(p.block_n, F16 ? 1 : 0);
    int s = 0, tcount = 0;
    uint32_t ph = 0;
    const uint64_t a_desc0 = make_sw128_desc(smem_a), b_desc0 = make_sw128_desc(smem_b);
    const uint32_t b_stage16 = b_stage_bytes >> 4;
    if (p.b_res && tile_first < total_tiles) mbar_wait(bar_bres, 0);
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++tcount) {
      const int ab = tcount & 1;
      const long long m0 = IG_T0();
      mbar_wait(bar_tempty + 8u * ab, ((tcount >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
      IG_ACC(2, m0);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.tmem_cols);
      for (int kb = 0; kb < KB; ++kb) {
        const long long m1 = IG_T0();
        mbar_wait(bar_full + 8u * s, ph);
        IG_ACC(1, m1);
        if (p.a_mode == 0) fence_proxy_async_smem();   // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
        tc_fence_after();
        const uint64_t a_desc = a_desc0 + (uint64_t)((uint32_t)s * (A_STAGE_BYTES >> 4));
        const uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)(p.b_res ? kb : s) * b_stage16);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
            // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the (>>4) start-address field
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0);
          }
          umma_commit(bar_empty + 8u * s);     // smem stage reusable once these MMAs have read it
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_tfull + 8u * ab);      // accumulator complete
      __syncwarp();
    }
  }

  if (p.epi_mode && !store_thread && tid == NUM_PRODUCER_THREADS) tma_store_wait_all();     // the storing thread: writes have landed
  __syncthreads();
  if (dbg && tid == 0) dbg[0] += clock64() - t_kernel;
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------ host
int g_num_sms = 148;

// C staging buffers: two (epilogue of tile t overlaps the residual load of tile t+1 and the store of tile t-1) when at
// least 3 pipeline stages still fit next to them
int staging_bufs(int block_n, int epi_mode) {
  if (!epi_mode) return 0;
  const int one = ((block_n + 63) / 64) * 16384;
  const int stage_bytes = A_STAGE_BYTES + block_n * 128;
  // a second C staging buffer only when it still leaves 4 ring stages (measured with 3 -> 4: conv2d_4a 383 -> 360 us,
  // mixed_6a 3x3 convolutions 113 -> 104 and 161 -> 153 us; nothing else moves).  VNFR_IG_MIN_STAGES overrides.
  static const int min_stages = getenv("VNFR_IG_MIN_STAGES") ? atoi(getenv("VNFR_IG_MIN_STAGES")) : 4;
  return (220 * 1024 - 2 * one) / stage_bytes >= min_stages ? 2 : 1;
}
int staging_bytes(int block_n, int epi_mode) { return staging_bufs(block_n, epi_mode) * ((block_n + 63) / 64) * 16384; }

int pick_stages(int block_n, int epi_mode, int b_res_kb) {
  // b_res_kb > 0: the weights (b_res_kb K blocks) are resident, the ring carries only A
  const int stage_bytes = A_STAGE_BYTES + (b_res_kb ? 0 : block_n * 128);
  int s = (220 * 1024 - staging_bytes(block_n, epi_mode) - b_res_kb * block_n * 128) / stage_bytes;   // one persistent CTA per SM
  if (s < 2) s = 2;
  if (s > MAX_STAGES) s = MAX_STAGES;
  return s;
}

size_t smem_bytes_for(int block_n, int stages, int epi_mode, int b_res_kb) {
  return 1024 /*alignment slack*/ + (size_t)stages * (A_STAGE_BYTES + (b_res_kb ? 0 : block_n * 128)) + (size_t)b_res_kb * block_n * 128 +
         staging_bytes(block_n, epi_mode) + 2048 /*bias*/ + 16 * stages + 104;
}

// Resident weights pay off for the small-K projection convolutions: TMA-fed 1x1 with a staged epilogue, a weight slab of at
// most 96 KB that still leaves >= 3 A stages, several N tiles and enough M tiles per CTA column.
bool use_resident_weights(const VnfrConvOp* op, int k_blocks, int n_tiles_m, int n_tiles_n) {
  if (op->a_mode != 1 || !op->epi_mode || getenv("VNFR_NO_BRES") != nullptr) return false;
  const int slab = k_blocks * op->block_n * 128;
  if (slab > 96 * 1024 || n_tiles_n < 2 || n_tiles_n > 74) return false;
  if ((220 * 1024 - staging_bytes(op->block_n, op->epi_mode) - slab) / A_STAGE_BYTES < 3) return false;
  return n_tiles_m >= 4 * (148 / n_tiles_n);
}

}  // namespace

extern long long g_vnfr_launches;

// debug hook (not part of include/vnfr_b200.h): device buffer of 16 int64 cycle counters, or null to switch off
extern "C" int vnfr_ig_debug(long long* dev_buf) {
  VNFR_CUDA(cudaMemcpyToSymbol(g_ig_dbg, &dev_buf, sizeof(dev_buf)));
  return VNFR_OK;
}

extern "C" int vnfr_conv_prepare(VnfrConvOp* op) {
  VNFR_REQUIRE(op != nullptr, "op is null");
  VNFR_REQUIRE(op->cin % 8 == 0 && op->in_pitch % 8 == 0, "cin and in_pitch must be multiples of 8 (16-byte gathers)");
  VNFR_REQUIRE(op->block_n % 16 == 0 && op->block_n >= 16 && op->block_n <= 256, "block_n must be a multiple of 16 in [16,256]");
  VNFR_REQUIRE(op->cout % 16 == 0 && op->cout_pad % op->block_n == 0 && op->cout_pad >= op->cout, "bad cout / cout_pad");
  VNFR_REQUIRE(op->k_pad % BLOCK_K == 0 && op->k_pad >= op->kh * op->kw * op->cin, "k_pad must be a multiple of 64 covering K");
  VNFR_REQUIRE(op->n_split % 16 == 0, "n_split must be a multiple of 16");
  VNFR_REQUIRE(op->dtype == 0 || op->dtype == 1, "dtype must be 0 (bf16) or 1 (fp16)");
  VNFR_REQUIRE(op->out_f32 != nullptr || op->out0 != nullptr, "no destination");
  VNFR_REQUIRE(op->out_f32 != nullptr || ((op->out0_pitch % 8 == 0) && (op->n_split >= op->cout || (op->out1 != nullptr && op->out1_pitch % 8 == 0))),
               "bf16 destinations need pitches that are multiples of 8");
  VNFR_REQUIRE(op->residual == nullptr || op->res_pitch % 8 == 0, "res_pitch must be a multiple of 8");
  VNFR_REQUIRE(op->out_h == (op->in_h + 2 * op->pad_h - op->kh) / op->stride + 1 &&
                   op->out_w == (op->in_w + 2 * op->pad_w - op->kw) / op->stride + 1,
               "out_h/out_w inconsistent with the convolution geometry");
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) {
    vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return VNFR_ERR_CUDA;
  }
  op->a_mode = 0;
  if (op->reserved[0] != 0 && getenv("VNFR_NO_SV") == nullptr) {
    // shifted-view kernel requested (reserved[0] = channels per plane); falls back to the generic path when the
    // geometry does not qualify (the caller checks a_mode when its weight packing is not generic-compatible)
    if (vnfr_sv_prepare(op) == VNFR_OK) return VNFR_OK;
  }
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)op->k_pad, (cuuint64_t)op->cout_pad};
  const cuuint64_t strides[1] = {(cuuint64_t)op->k_pad * 2};
  const cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)op->block_n};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, op->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(op->weights), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed");
    return VNFR_ERR_CUDA;
  }
  memcpy(op->tmap_w, &tm, sizeof(tm));
  // staged epilogue (TMA residual load + TMA store): single 16-bit destination whose rows are 16-byte aligned
  op->epi_mode = 0;
  {
    const long long Mo = (long long)op->n_img * op->out_h * op->out_w;
    // every stored 64-channel panel must lie inside its own N tile: block_n a multiple of 64, or a single N tile
    const bool single = op->out_f32 == nullptr && op->n_split >= op->cout && op->out0 != nullptr &&
                        (op->block_n % 64 == 0 || op->cout <= op->block_n);
    if (single && Mo > 0 && ((uintptr_t)op->out0 % 16 == 0) && op->out0_pitch % 8 == 0 &&
        (op->residual == nullptr || (((uintptr_t)op->residual % 16 == 0) && op->res_pitch % 8 == 0)) &&
        getenv("VNFR_NO_TMA_EPILOGUE") == nullptr) {
      const CUtensorMapDataType dt = op->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
      const cuuint64_t cdims[2] = {(cuuint64_t)op->cout, (cuuint64_t)Mo};
      const cuuint32_t cbox[2] = {64, BLOCK_M};
      CUtensorMap tcm, trm;
      const cuuint64_t cstr[1] = {(cuuint64_t)op->out0_pitch * 2};
      bool ok = enc(&tcm, dt, 2, op->out0, cdims, cstr, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      if (ok && op->residual != nullptr) {
        const cuuint64_t rstr[1] = {(cuuint64_t)op->res_pitch * 2};
        ok = enc(&trm, dt, 2, const_cast<void*>(op->residual), cdims, rstr, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        if (ok) memcpy(op->tmap_r, &trm, sizeof(trm));
      }
      if (ok) {
        memcpy(op->tmap_c, &tcm, sizeof(tcm));
        op->epi_mode = 1;
      }
    }
  }
  // A operand: 1x1 / stride 1 / no padding convolutions read a plain [M][in_pitch] matrix -> 2-D tiled TMA
  const long long M = (long long)op->n_img * op->out_h * op->out_w;
  if (op->kh == 1 && op->kw == 1 && op->stride == 1 && op->pad_h == 0 && op->pad_w == 0 && M > 0 &&
      ((uintptr_t)op->in % 16 == 0) && getenv("VNFR_NO_TMA_A") == nullptr) {
    CUtensorMap ta;
    const cuuint64_t adims[2] = {(cuuint64_t)op->cin, (cuuint64_t)M};
    const cuuint64_t astrides[1] = {(cuuint64_t)op->in_pitch * 2};
    const cuuint32_t abox[2] = {BLOCK_K, BLOCK_M};
    CUresult ra = enc(&ta, op->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                      const_cast<void*>(op->in), adims, astrides, abox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (ra == CUDA_SUCCESS) {
      memcpy(op->tmap_a, &ta, sizeof(ta));
      op->a_mode = 1;
    }
  }
  // k x k convolutions whose input channel count is a multiple of 64: the A tile of K block (tap, 64-channel chunk) is one TMA
  // IM2COL load (cuTensorMapEncodeIm2col: NHWC tensor {C, W, H, N}, bounding box corners = -pad / pad - (k - 1), traversal
  // stride = convolution stride, 64 channels x 128 pixels per load); the packed weights' K order (tap-major, cin contiguous)
  // already puts such a block in one 64-wide K block.  VNFR_NO_IM2COL=1 keeps the cp.async gather.
  if (op->a_mode == 0 && op->kh * op->kw > 1 && op->cin % 64 == 0 && op->in_pitch % 8 == 0 && ((uintptr_t)op->in % 16 == 0) && M > 0 &&
      op->stride >= 1 && op->stride <= 8 && op->pad_h <= 127 && op->pad_w <= 127 && op->kh <= 128 && op->kw <= 128 &&
      op->k_pad == op->kh * op->kw * op->cin && getenv("VNFR_NO_IM2COL") == nullptr) {
    typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                       const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeIm2colFn enc_i2c = nullptr;
    if (enc_i2c == nullptr) {
      void* sym = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        enc_i2c = reinterpret_cast<EncodeIm2colFn>(sym);
    }
    if (enc_i2c != nullptr) {
      CUtensorMap ta;
      const cuuint64_t gdim[4] = {(cuuint64_t)op->cin, (cuuint64_t)op->in_w, (cuuint64_t)op->in_h, (cuuint64_t)op->n_img};
      const cuuint64_t gstr[3] = {(cuuint64_t)op->in_pitch * 2, (cuuint64_t)op->in_w * op->in_pitch * 2,
                                  (cuuint64_t)op->in_h * op->in_w * op->in_pitch * 2};
      const int lower[2] = {-op->pad_w, -op->pad_h};                                       // {W, H}
      const int upper[2] = {op->pad_w - (op->kw - 1), op->pad_h - (op->kh - 1)};
      const cuuint32_t estr4[4] = {1, (cuuint32_t)op->stride, (cuuint32_t)op->stride, 1};
      const CUresult ri = enc_i2c(&ta, op->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                  const_cast<void*>(op->in), gdim, gstr, lower, upper, 64, BLOCK_M, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (ri == CUDA_SUCCESS) {
        memcpy(op->tmap_a, &ta, sizeof(ta));
        op->a_mode = 2;
        // TMA-fed k x k convolutions are bound by the depth of the operand ring (tools/ig_probe.py: the MMA warp waits for a
        // full stage 38 % of the time with 3 stages): they give the C staging panels' shared memory to the ring and store
        // rows directly (measured per launch: 192 -> 181, 113 -> 100, 161 -> 150, 52 -> 46, 56 -> 55 us)
        if (getenv("VNFR_IM2COL_STAGED") == nullptr) op->epi_mode = 0;
      }
    }
  }
  return VNFR_OK;
}

extern "C" int vnfr_conv_run(const VnfrConvOp* op, void* stream) {
  VNFR_REQUIRE(op != nullptr, "op is null");
  if (op->a_mode == 3) return vnfr_sv_run(op, stream);
  static VnfrPerDevice attr_set_once = {};
  if (vnfr_first_on_device(attr_set_once)) {
    VNFR_CUDA(cudaFuncSetAttribute(igemm_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(igemm_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  ConvParams p;
  p.in = (const __nv_bfloat16*)op->in;
  p.bias = op->bias;
  p.residual = (const __nv_bfloat16*)op->residual;
  p.out0 = (__nv_bfloat16*)op->out0;
  p.out1 = (__nv_bfloat16*)op->out1;
  p.out_f32 = op->out_f32;
  p.n_img = op->n_img; p.in_h = op->in_h; p.in_w = op->in_w; p.cin = op->cin; p.in_pitch = op->in_pitch;
  p.kh = op->kh; p.kw = op->kw; p.stride = op->stride; p.pad_h = op->pad_h; p.pad_w = op->pad_w;
  p.out_h = op->out_h; p.out_w = op->out_w;
  p.M = op->n_img * op->out_h * op->out_w;
  p.K = op->kh * op->kw * op->cin;
  p.cout = op->cout;
  p.k_blocks = ceil_div(p.K, BLOCK_K);
  p.block_n = op->block_n;
  p.n_split = op->n_split; p.out0_pitch = op->out0_pitch; p.out1_pitch = op->out1_pitch;
  p.res_pitch = op->res_pitch; p.out_f32_pitch = op->out_f32_pitch;
  p.relu = op->relu;
  p.dtype = op->dtype;
  p.a_mode = op->a_mode;
  p.alpha = nullptr;
  p.epi_mode = op->epi_mode;
  p.c_bufs = staging_bufs(op->block_n, op->epi_mode);
  p.n_tiles_m = ceil_div(p.M, BLOCK_M);
  p.n_tiles_n = ceil_div(op->cout, op->block_n);
  p.b_res = use_resident_weights(op, p.k_blocks, p.n_tiles_m, p.n_tiles_n) ? 1 : 0;
  const int b_res_kb = p.b_res ? p.k_blocks : 0;
  p.stages = pick_stages(op->block_n, op->epi_mode, b_res_kb);
  int cols = 32;
  while (cols < op->block_n) cols <<= 1;
  p.tmem_cols = cols;
  if (p.M <= 0) return VNFR_OK;
  CUtensorMap tm, ta, tc_, tr;
  memcpy(&tm, op->tmap_w, sizeof(tm));
  memcpy(&ta, op->a_mode != 0 ? op->tmap_a : op->tmap_w, sizeof(ta));
  memcpy(&tc_, op->epi_mode ? op->tmap_c : op->tmap_w, sizeof(tc_));
  memcpy(&tr, (op->epi_mode && op->residual != nullptr) ? op->tmap_r : op->tmap_w, sizeof(tr));
  const int total_tiles = p.n_tiles_m * p.n_tiles_n;
  dim3 grid(total_tiles < g_num_sms ? total_tiles : g_num_sms);
  if (p.b_res) grid = dim3((g_num_sms / p.n_tiles_n) * p.n_tiles_n);
  if (op->dtype == 1)
    VNFR_CUDA(launch_pdl(igemm_conv_kernel<true>, (int)grid.x, NUM_THREADS, smem_bytes_for(op->block_n, p.stages, op->epi_mode, b_res_kb),
                         (cudaStream_t)stream, tm, ta, tc_, tr, p));
  else
    VNFR_CUDA(launch_pdl(igemm_conv_kernel<false>, (int)grid.x, NUM_THREADS, smem_bytes_for(op->block_n, p.stages, op->epi_mode, b_res_kb),
                         (cudaStream_t)stream, tm, ta, tc_, tr, p));
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_run_ops(const VnfrOp* ops, int n_ops, void* stream) {
  VNFR_REQUIRE(ops != nullptr || n_ops == 0, "ops is null");
  for (int i = 0; i < n_ops; ++i) {
    const VnfrConvOp* c = &ops[i].conv;
    int rc;
    switch (ops[i].kind) {
      case 0: rc = vnfr_conv_run(c, stream); break;
      case 1: rc = vnfr_maxpool3s2_nhwc(c->in, c->n_img, c->in_h, c->in_w, c->cin, c->in_pitch, c->out0, c->out0_pitch, c->dtype, stream); break;
      case 2: rc = vnfr_avgpool_nhwc(c->in, c->n_img, c->in_h * c->in_w, c->cin, c->in_pitch, c->out0, c->dtype, stream); break;
      case 3: rc = vnfr_block17_run(reinterpret_cast<const VnfrBlock17Op*>(ops[i].ext), stream); break;
      default: vnfr_set_error(__FILE__, __LINE__, "unknown op kind"); return VNFR_ERR_ARG;
    }
    if (rc != VNFR_OK) return rc;
  }
  return VNFR_OK;
}
