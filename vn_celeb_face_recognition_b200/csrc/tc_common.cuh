// Shared pieces of the tcgen05 convolution kernels (igemm_conv.cu: gathered / 1x1-TMA implicit GEMM; sv_conv.cu:
// shifted-view convolution on a shared-memory-resident input band): PTX wrappers, descriptors, the fused epilogue.
#pragma once
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <cuda_fp16.h>

namespace tc {

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
  int epi_mode;  // 1: shared-memory staged epilogue (TMA residual load, TMA store)
  int c_bufs;    // staging buffers of the staged epilogue (1 or 2)
  int b_res;     // igemm: the CTA's weight slab (all K blocks of one N tile) is resident in shared memory
  const float* alpha;   // nullable: per-channel PReLU slope (applied instead of ReLU)
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
// Programmatic dependent launch: `launch_dependents` lets the NEXT kernel of the stream (launched with the programmatic
// stream-serialisation attribute, see launch_pdl) start its prologue -- barrier init, tensor-map prefetch, TMEM allocation,
// weight loads -- on SMs this grid has already left; `wait` blocks until every prerequisite grid has completed and its
// memory is visible.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
// One lane of a CONVERGED warp (the MMA / TMA issue loops run warp-uniformly so that descriptors and loop counters stay
// in uniform registers; only the issue instruction itself is predicated on the elected lane).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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


// K-major swizzled shared-memory matrix descriptor for rows of `row_bytes` (128 / 64 / 32 = the swizzle span), 8-row
// groups 8*row_bytes apart.  `base_offset` is the 3-bit phase field used when the start address is not aligned to the
// swizzle pattern period (8 rows).
__device__ __forceinline__ uint64_t make_sw_desc(uint32_t smem_addr, uint32_t row_bytes, uint32_t base_offset) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7u) << 49;
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Staged epilogue of one accumulator row (epi_mode 1).  The C tile lives in shared memory as 64-channel panels of
// [128 rows][128 B], 128B-swizzled exactly as TMA writes / reads them: 16-byte chunk k of row r sits at
// r*128 + ((k ^ (r & 7)) << 4), which also makes this one-row-per-thread access pattern bank-conflict free.  When the op
// has a residual the panels already hold it (TMA-loaded while the MMAs ran); the result overwrites it in place and one
// thread then TMA-stores the panels (coalesced, clipped at the tensor bounds) -- no per-thread global access at all.
template <bool F16>
__device__ __forceinline__ void epilogue_row_staged(const ConvParams& p, const float* sb, uint32_t t_row, uint32_t smem_c, int r,
                                                    int n_valid, int chalf, bool has_res) {
  for (int c0 = chalf * 16; c0 < n_valid; c0 += 32) {
    float v[16];
    __syncwarp();
    tmem_ld16_issue(t_row + (uint32_t)c0, v);
    const uint32_t panel = smem_c + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)r * 128u;
    const uint32_t k0 = (uint32_t)(c0 & 63) >> 3;
    const uint32_t a0 = panel + (((k0) ^ (uint32_t)(r & 7)) << 4), a1 = panel + (((k0 + 1) ^ (uint32_t)(r & 7)) << 4);
    uint4 rc0 = make_uint4(0, 0, 0, 0), rc1 = rc0;
    if (has_res) { rc0 = lds128(a0); rc1 = lds128(a1); }
    tmem_ld_wait(v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * i);
      v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
    }
    if (has_res) {
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
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack2<F16>(v[2 * i], v[2 * i + 1]);
    sts128(a0, make_uint4(pk[0], pk[1], pk[2], pk[3]));
    sts128(a1, make_uint4(pk[4], pk[5], pk[6], pk[7]));
  }
}

// Fused epilogue of one accumulator row: TMEM -> registers -> +bias (+residual) (ReLU) -> 16-bit (or fp32) store into the
// channel slice(s) of the destination.  Warp-collective (tcgen05.ld): every lane of the warp must call it, `row_ok`
// masks the stores.  The calling warp handles the 16-column chunks c0 = chalf*16, chalf*16 + 32, ...
//   t_row: TMEM address of this warp's lane quarter / accumulator buffer; m: output pixel index; sb: staged bias.
template <bool F16>
__device__ __forceinline__ void epilogue_row(const ConvParams& p, const float* sb, uint32_t t_row, int m, bool row_ok, int n0,
                                             int n_valid, int chalf) {
  const __nv_bfloat16* res_row = p.residual != nullptr && row_ok ? p.residual + (size_t)m * p.res_pitch + n0 : nullptr;
  uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
  if (res_row != nullptr && chalf * 16 < n_valid) {
    rn0 = __ldg(reinterpret_cast<const uint4*>(res_row + chalf * 16));
    rn1 = __ldg(reinterpret_cast<const uint4*>(res_row + chalf * 16) + 1);
  }
  for (int c0 = chalf * 16; c0 < n_valid; c0 += 32) {
    float v[16];
    __syncwarp();
    tmem_ld16_issue(t_row + (uint32_t)c0, v);        // warp-collective, also for masked rows
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
      if (p.alpha != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * __ldg(p.alpha + n + i);
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
}

// Narrow-tile variant (NC = 32 or 64 columns, one destination, no residual): all tcgen05.ld of the row are issued back to
// back and waited for once, so the TMEM load latency is paid once per tile instead of once per 16-column chunk.
template <bool F16, int NC>
__device__ __forceinline__ void epilogue_row_narrow(const ConvParams& p, const float* sb, uint32_t t_row, int m, bool row_ok) {
  float v[NC];
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld16_issue(t_row + (uint32_t)(16 * c), v + 16 * c);
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld_wait(v + 16 * c);
  if (!row_ok) return;
#pragma unroll
  for (int i = 0; i < NC / 4; ++i) {
    const float4 b4 = *reinterpret_cast<const float4*>(sb + 4 * i);
    v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (p.alpha != nullptr) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * __ldg(p.alpha + i);
  }
  if (p.out_f32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(p.out_f32 + (size_t)m * p.out_f32_pitch);
#pragma unroll
    for (int qq = 0; qq < NC / 4; ++qq) o[qq] = make_float4(v[4 * qq], v[4 * qq + 1], v[4 * qq + 2], v[4 * qq + 3]);
  } else {
    uint4* o = reinterpret_cast<uint4*>(p.out0 + (size_t)m * p.out0_pitch);
#pragma unroll
    for (int qq = 0; qq < NC / 8; ++qq)
      o[qq] = make_uint4(pack2<F16>(v[8 * qq], v[8 * qq + 1]), pack2<F16>(v[8 * qq + 2], v[8 * qq + 3]),
                         pack2<F16>(v[8 * qq + 4], v[8 * qq + 5]), pack2<F16>(v[8 * qq + 6], v[8 * qq + 7]));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
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

// Kernel launch with (pdl = true) or without the programmatic stream-serialisation attribute.  VNFR_NO_PDL=1 disables it.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool pdl = getenv("VNFR_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace tc

// shifted-view convolution (sv_conv.cu)
int vnfr_sv_prepare(VnfrConvOp* op);                       // returns VNFR_OK and sets op->a_mode = 3 when the op qualifies
int vnfr_sv_run(const VnfrConvOp* op, void* stream);
