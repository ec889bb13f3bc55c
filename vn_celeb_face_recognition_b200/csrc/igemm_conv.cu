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
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <cuda_fp16.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // bf16 elements: 128 bytes = one swizzle-128B row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_PRODUCER_THREADS = 256;         // warps 0-7: A gather, two threads per GEMM row
constexpr int NUM_EPILOGUE_THREADS = 256;         // warps 8-15: two warps per TMEM lane quarter, alternating 16-column chunks
constexpr int NUM_THREADS = 576;                  // warp 16: TMA producer, warp 17: TMEM alloc + MMA issue
constexpr int MAX_STAGES = 8;

struct ConvParams {
  const __nv_bfloat16* in;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out0;
  __nv_bfloat16* out1;
  float* out_f32;
  int n_img, in_h, in_w, cin, in_pitch;
  int kh, kw, stride, pad_h, pad_w;
  int out_h, out_w;
  int M, K, cout, k_blocks, block_n, n_tiles_m, n_tiles_n;
  int n_split, out0_pitch, out1_pitch, res_pitch, out_f32_pitch;
  int relu;
  int stages, tmem_cols;
  int a_mode;  // 0: cp.async gather by warps 0-7, 1: TMA tiled 2-D (1x1 convs), 2: TMA im2col
  int dtype;   // 0 = bf16, 1 = fp16 (both: fp32 accumulation in TMEM)
};

// ------------------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must trap (launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// one (non-incrementing) arrival on `bar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset: 8 rows * 128 B  [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16 (format 1) or fp16 (format 0), both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t make_idesc_f16(int n, int is_fp16) {
  const uint32_t fmt = is_fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
// 16-bit float helpers parameterised on the storage type (F16 = true: IEEE half, false: bfloat16)
template <bool F16>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
  if (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x; hi = f.y;
  } else {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xFFFF0000u);
  }
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (F16) {
    // saturate instead of overflowing to inf (fp16 range 65504)
    a = fminf(fmaxf(a, -65504.f), 65504.f); b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// the "+r" ties make every later use of v[] depend on the wait
__device__ __forceinline__ void tmem_ld_wait(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

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
template <bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const uint32_t b_stage_bytes = (uint32_t)p.block_n * 128u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + (uint32_t)S * A_STAGE_BYTES;
  const uint32_t smem_bias = smem_b + (uint32_t)S * b_stage_bytes;          // 2 x 256 floats
  const uint32_t bars = smem_bias + 2048u;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * S, bar_tfull = bars + 16u * S, bar_tempty = bar_tfull + 16u,
                 tmem_slot = bar_tempty + 16u;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (smem_bias - smem_u32(smem_raw)));

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int KB = p.k_blocks;
  const int n_tiles_n = p.n_tiles_n;
  const int total_tiles = p.n_tiles_m * n_tiles_n;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8u * s, p.a_mode == 0 ? NUM_PRODUCER_THREADS + 1 : 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8u * i, 1);
      mbar_init(bar_tempty + 8u * i, NUM_EPILOGUE_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 16 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    if (p.a_mode != 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)(2 * p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
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
    int it = 0;                                // K blocks issued so far (all tiles)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
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
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % S;
        const bool k_ok = kf < p.K;
        mbar_wait(bar_empty + 8u * s, ((it / S) & 1) ^ 1);
        const uint32_t stage = smem_a + (uint32_t)s * A_STAGE_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int iy = iy0[q] + ky, ix = ix0[q] + kx;
          const bool ok = k_ok && (unsigned)iy < (unsigned)p.in_h && (unsigned)ix < (unsigned)p.in_w;
          const void* src = ok ? (const void*)(img_base[q] + ((size_t)iy * p.in_w + ix) * p.in_pitch + c) : (const void*)p.in;
          cp_async_16(stage + dst_off[q], src, ok ? 16u : 0u);
        }
        cp_async_arrive_noinc(bar_full + 8u * s);
        kf += 64;
        c += 64;
        while (c >= p.cin) { c -= p.cin; if (++kx == p.kw) { kx = 0; ++ky; } }
      }
    }
    cp_async_wait<0>();                        // nothing may be in flight when the CTA retires
    }  // a_mode == 0 (with TMA-fed A these warps are idle)
  } else if (warp < 16) {
    // ================================================= epilogue: TMEM -> registers -> bias/residual/ReLU -> global
    const int q = warp & 3;                    // TMEM lane quarter of this warp
    const int r = q * 32 + lane;               // row inside the tile
    const int et = tid - NUM_PRODUCER_THREADS; // 0..255
    const int chalf = (warp - 8) >> 2;         // this warp handles 16-column chunks with (chunk & 1) == chalf
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int ab = tcount & 1;
      const int m = (tile / n_tiles_n) * BLOCK_M + r;
      const int n0 = (tile % n_tiles_n) * p.block_n;
      const bool row_ok = m < p.M;
      const int n_valid = min(p.block_n, p.cout - n0);
      // stage this tile's bias while the MMAs are still running
      float* sb = s_bias + ab * 256;
      for (int i = et; i < n_valid; i += NUM_EPILOGUE_THREADS) sb[i] = __ldg(p.bias + n0 + i);
      const __nv_bfloat16* res_row = p.residual != nullptr && row_ok ? p.residual + (size_t)m * p.res_pitch + n0 : nullptr;
      uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
      if (res_row != nullptr && chalf * 16 < n_valid) {
        rn0 = __ldg(reinterpret_cast<const uint4*>(res_row + chalf * 16));
        rn1 = __ldg(reinterpret_cast<const uint4*>(res_row + chalf * 16) + 1);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPILOGUE_THREADS) : "memory");      // bias visible to the 4 epilogue warps
      mbar_wait(bar_tfull + 8u * ab, (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.tmem_cols);
      for (int c0 = chalf * 16; c0 < n_valid; c0 += 32) {
        float v[16];
        __syncwarp();
        tmem_ld16_issue(t_row + (uint32_t)c0, v);        // warp-collective, also for rows >= M
        const uint4 rc0 = rn0, rc1 = rn1;
        if (res_row != nullptr && c0 + 32 < n_valid) {   // prefetch the next chunk's residual under this chunk's math
          rn0 = __ldg(reinterpret_cast<const uint4*>(res_row + c0 + 32));
          rn1 = __ldg(reinterpret_cast<const uint4*>(res_row + c0 + 32) + 1);
        }
        tmem_ld_wait(v);
        if (row_ok) {
          const int n = n0 + c0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * i);
            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          }
          if (res_row != nullptr) {
            const uint32_t w[8] = {rc0.x, rc0.y, rc0.z, rc0.w, rc1.x, rc1.y, rc1.z, rc1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float lo, hi;
              unpack2<F16>(w[e], lo, hi);
              v[2 * e + 0] += lo;
              v[2 * e + 1] += hi;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          if (p.out_f32 != nullptr) {
            float4* o = reinterpret_cast<float4*>(p.out_f32 + (size_t)m * p.out_f32_pitch + n);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) o[qq] = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
          } else {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack2<F16>(v[2 * i], v[2 * i + 1]);
            __nv_bfloat16* dst = (n < p.n_split) ? p.out0 + (size_t)m * p.out0_pitch + n
                                                 : p.out1 + (size_t)m * p.out1_pitch + (n - p.n_split);
            uint4* o = reinterpret_cast<uint4*>(dst);
            o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8u * ab);       // accumulator buffer `ab` may be overwritten
    }
  } else if (warp == 16) {
    // ================================================= W producer: TMA, one elected lane
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles_n) * p.block_n;
        const int m0 = (tile / n_tiles_n) * BLOCK_M;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(bar_empty + 8u * s, ((it / S) & 1) ^ 1);
          if (p.a_mode == 1) {
            // 1x1 convolution: the A tile is a plain 2-D box of the [M][pitch] activation matrix
            mbar_arrive_expect_tx(bar_full + 8u * s, b_stage_bytes + A_STAGE_BYTES);
            tma_load_2d(smem_a + (uint32_t)s * A_STAGE_BYTES, &tmap_a, bar_full + 8u * s, kb * BLOCK_K, m0);
          } else {
            mbar_arrive_expect_tx(bar_full + 8u * s, b_stage_bytes);
          }
          tma_load_2d(smem_b + (uint32_t)s * b_stage_bytes, &tmap_w, bar_full + 8u * s, kb * BLOCK_K, n0);
        }
      }
    }
  } else {
    // ================================================= MMA issuer: one elected lane
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(p.block_n, F16 ? 1 : 0);
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int ab = tcount & 1;
        mbar_wait(bar_tempty + 8u * ab, ((tcount >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * p.tmem_cols);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(bar_full + 8u * s, (it / S) & 1);
          if (p.a_mode == 0) fence_proxy_async_smem();   // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
          tc_fence_after();
          const uint64_t a_desc = make_sw128_desc(smem_a + (uint32_t)s * A_STAGE_BYTES);
          const uint64_t b_desc = make_sw128_desc(smem_b + (uint32_t)s * b_stage_bytes);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
            // advance 16 elements = 32 bytes along K inside the swizzle atom: +2 in the (>>4) start-address field
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * kk), b_desc + (uint64_t)(2 * kk), idesc, (kb | kk) != 0);
          }
          umma_commit(bar_empty + 8u * s);     // smem stage reusable once these MMAs have read it
        }
        umma_commit(bar_tfull + 8u * ab);      // accumulator complete
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * p.tmem_cols)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int g_num_sms = 148;

int pick_stages(int block_n) {
  const int stage_bytes = A_STAGE_BYTES + block_n * 128;
  int s = (220 * 1024) / stage_bytes;          // one persistent CTA per SM owns (almost) all of its shared memory
  if (s < 2) s = 2;
  if (s > MAX_STAGES) s = MAX_STAGES;
  return s;
}

size_t smem_bytes_for(int block_n, int stages) {
  return 1024 /*alignment slack*/ + (size_t)stages * (A_STAGE_BYTES + block_n * 128) + 2048 /*bias*/ + 16 * stages + 64;
}

}  // namespace

extern long long g_vnfr_launches;

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
  // A operand: 1x1 / stride 1 / no padding convolutions read a plain [M][in_pitch] matrix -> 2-D tiled TMA
  op->a_mode = 0;
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
  return VNFR_OK;
}

extern "C" int vnfr_conv_run(const VnfrConvOp* op, void* stream) {
  VNFR_REQUIRE(op != nullptr, "op is null");
  static bool attr_set = false;
  if (!attr_set) {
    VNFR_CUDA(cudaFuncSetAttribute(igemm_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VNFR_CUDA(cudaFuncSetAttribute(igemm_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
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
  p.stages = pick_stages(op->block_n);
  int cols = 32;
  while (cols < op->block_n) cols <<= 1;
  p.tmem_cols = cols;
  if (p.M <= 0) return VNFR_OK;
  CUtensorMap tm, ta;
  memcpy(&tm, op->tmap_w, sizeof(tm));
  memcpy(&ta, op->a_mode != 0 ? op->tmap_a : op->tmap_w, sizeof(ta));
  p.n_tiles_m = ceil_div(p.M, BLOCK_M);
  p.n_tiles_n = ceil_div(op->cout, op->block_n);
  const int total_tiles = p.n_tiles_m * p.n_tiles_n;
  dim3 grid(total_tiles < g_num_sms ? total_tiles : g_num_sms);
  if (op->dtype == 1)
    igemm_conv_kernel<true><<<grid, NUM_THREADS, smem_bytes_for(op->block_n, p.stages), (cudaStream_t)stream>>>(tm, ta, p);
  else
    igemm_conv_kernel<false><<<grid, NUM_THREADS, smem_bytes_for(op->block_n, p.stages), (cudaStream_t)stream>>>(tm, ta, p);
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
      default: vnfr_set_error(__FILE__, __LINE__, "unknown op kind"); return VNFR_ERR_ARG;
    }
    if (rc != VNFR_OK) return rc;
  }
  return VNFR_OK;
}
