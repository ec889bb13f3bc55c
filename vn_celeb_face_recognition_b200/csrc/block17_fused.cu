// Block17 of InceptionResnetV1 as ONE persistent tcgen05 kernel (models/inception_resnet_v1.py:70-95):
//
//   x0 = branch0(x)            1x1 896 -> 128 (+BN+ReLU)      \  one GEMM, N = 256
//   t  = branch1.0(x)          1x1 896 -> 128 (+BN+ReLU)      /
//   u  = branch1.1(t)          1x7 128 -> 128, pad (0,3)
//   x1 = branch1.2(u)          7x1 128 -> 128, pad (3,0)
//   x  = relu(x + 0.10 * conv2d(cat(x0, x1)))   1x1 256 -> 896 (+bias), residual scale folded into the weights
//
// The unfused path runs these as four launches that round-trip t / u / cat through L2 and re-stream every operand per
// layer; measured (profiles/r1_encoder_layer_table.txt) they are bound by L2 -> SM operand traffic, not by the tensor pipe.
// Here one CTA owns a tile of TWO 8x8 images (128 GEMM rows) from x to the updated x: t, u and cat never leave shared
// memory, the accumulators live in TMEM, and per tile only x (once as A operand, once as residual), the four weight
// matrices and the output cross the SM boundary.
//
// Tile rows are kept in ONE order everywhere ("v-order": row = (y*2 + img)*8 + x, i.e. an eight-row group = one image row),
// delivered directly by a 4-D tensor map of x with box {64 ch, 8 x, 2 img, 8 y} (the same map stores the result).  A
// convolution along y (7x1) is then a shift by 2*dy whole groups = 2*dy KiB of the K-major 128B-swizzled operand: tap ky is
// the SAME buffer seen from a start address 2*ky KiB further, with 6 zero groups of padding before and after.  A
// convolution along x (1x7) is a shift by dx rows INSIDE every group: its operand keeps each group in 16 rows [4 zero |
// 8 | 4 zero] (stride between groups 2 KiB in the descriptor) and tap kx starts (kx + 1) rows into the buffer -- the UMMA
// swizzle is a pure function of the shared-memory address bits, so a start address that is not a multiple of 1 KiB reads
// correctly (measured; sv_conv.cu relies on the same property).  No transposes, no im2col, no data movement per tap.
//
// Shared memory (224 KB + barriers):  cat [4 planes of 64 ch][128 rows][128 B]  |  tpad 64 KB (t: [2 planes][16 groups][16
// rows][128 B], then u: [2 planes][28 groups][1 KB], later the 4 residual / output staging panels)  |  operand ring of 6 x 16 KB slots (x K-blocks and weight tiles).
// TMEM (512 columns): region 0 = [0,256): GEMM1, projection chunks 0 and 2; region 1 = [256,512): 1x7 (256..383),
// 7x1 (384..511), projection chunks 1 and 3.
// Warps: 0-15 epilogue (four per TMEM lane quarter, a quarter of the columns each), 16 operand producer (TMA), 17 MMA
// issuer + TMEM owner.  The residual "+ x" is added by the tensor core: the 64-channel panels of x ride the operand ring once
// more and are multiplied by a 16 x 16 identity into the projection accumulator (exact in fp32), so the projection epilogue
// is bias + ReLU + pack into four staging panels that TMA stores drain.  (Measured alternatives: residual panels TMA-loaded
// into the staging ring = a load -> epilogue -> store -> read-complete -> reload latency chain per panel, 19 k cycles per tile
// against 11 k cycles of projection MMAs; row-per-thread global loads / stores = 32 cache lines per warp instruction, 32 k.)
#include "tc_common.cuh"

using namespace tc;

namespace {

constexpr int B17_EPI = 512;                // 16 epilogue warps: four per TMEM lane quarter, each a quarter of the columns
constexpr int B17_THREADS = B17_EPI + 64;
constexpr uint32_t SLOT = 16384, NSLOT = 6;
constexpr uint32_t OFF_CAT = 0, OFF_TPAD = 65536, OFF_RING = 131072, OFF_BARS = OFF_RING + NSLOT * SLOT;
constexpr uint32_t TPAD_PLANE = 28 * 1024;
constexpr uint32_t OFF_I16 = OFF_BARS + 256;       // 16 x 16 identity (B operand of the residual MMAs), 32-byte swizzled rows
constexpr uint32_t B17_SMEM = OFF_I16 + 512;
constexpr int KB1 = 14;                    // 896 / 64 K blocks of the first GEMM

struct B17Params {
  const float* b1;   // [256] branch0 | branch1.0 folded-BN bias
  const float* b2;   // [128] 1x7
  const float* b3;   // [128] 7x1
  const float* b4;   // [896] projection bias * scale
  int n_img, n_tiles;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// ReLU + saturation + 16-bit pack of two fp32 values in ONE instruction (F2FP.SATFINITE.RELU)
template <bool F16>
__device__ __forceinline__ uint32_t relu_pack2(float lo, float hi) {
  uint32_t r;
  if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// bias + ReLU + 16-bit pack of 16 accumulator columns -> two 16-byte chunks of one swizzled 128-byte row
template <bool F16>
__device__ __forceinline__ void relu_pack16(float* v, const float* __restrict__ bias, uint4& lo, uint4& hi) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias) + i);
    v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
  }
  lo = make_uint4(relu_pack2<F16>(v[0], v[1]), relu_pack2<F16>(v[2], v[3]), relu_pack2<F16>(v[4], v[5]), relu_pack2<F16>(v[6], v[7]));
  hi = make_uint4(relu_pack2<F16>(v[8], v[9]), relu_pack2<F16>(v[10], v[11]), relu_pack2<F16>(v[12], v[13]), relu_pack2<F16>(v[14], v[15]));
}

// 32 accumulator columns [col0, col0+32) of this thread's TMEM lane -> (+bias, ReLU) -> the 16-byte chunks [chunk0, chunk0+4)
// of one 128-byte swizzled row at `row_addr` (1024-aligned plane base + row*128), swizzle phase `sw` = (row address >> 7) & 7.
// Both TMEM loads are in flight before the first wait.
template <bool F16>
__device__ __forceinline__ void drain32_to_row(uint32_t t_lane, uint32_t col0, const float* __restrict__ bias, uint32_t row_addr, uint32_t sw,
                                               uint32_t chunk0) {
  float v[32];
  __syncwarp();
  tmem_ld16_issue(t_lane + col0, v);
  tmem_ld16_issue(t_lane + col0 + 16u, v + 16);
  tmem_ld_wait(v);
  tmem_ld_wait(v + 16);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint4 lo, hi;
    relu_pack16<F16>(v + 16 * c, bias + 16 * c, lo, hi);
    sts128(row_addr + (((chunk0 + (uint32_t)(2 * c)) ^ sw) << 4), lo);
    sts128(row_addr + (((chunk0 + (uint32_t)(2 * c + 1)) ^ sw) << 4), hi);
  }
}

// K-major SWIZZLE_128B descriptor with an explicit stride between eight-row groups (make_sw128_desc: 1024)
__device__ __forceinline__ uint64_t make_sw128_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Optional cycle breakdown of CTA 0 (tools/b17_probe.py): [0] kernel; MMA warp waiting for [1] a GEMM1 ring step, [2] a
// conv ring step, [3] a projection ring step, [4] tpad (E1 / E2), [5] cat (E3), [6] a drained accumulator; epilogue thread 0:
// [8] waiting for GEMM1, [9] E1 work, [10] waiting for the 1x7, [11] E2 work, [12] waiting for the 7x1, [13] E3 work,
// [14] waiting for a projection chunk, [15] for a residual panel, [16] panel math, [17] panel barrier + store issue.
__device__ long long* g_b17_dbg = nullptr;
#define B17_T0() (dbg ? clock64() : 0ll)
#define B17_ACC(i, t0) do { if (dbg) dbg[i] += clock64() - (t0); } while (0)

template <bool F16>
__global__ void __launch_bounds__(B17_THREADS, 1)
block17_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w3,
                     const __grid_constant__ CUtensorMap tm_w4, const B17Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sm = smem_u32(smem_raw);
  const uint32_t cat = sm + OFF_CAT, tpad = sm + OFF_TPAD, ring = sm + OFF_RING, bars = sm + OFF_BARS;
  const uint32_t bar_full = bars, bar_empty = bars + 48, bar_tfull = bars + 96, bar_tempty = bars + 112,
                 bar_tpad = bars + 128, bar_cat = bars + 136, tmem_slot = bars + 144;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long* dbg = (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 17)) ? g_b17_dbg : nullptr;
  const long long t_kernel = B17_T0();

  if ((sm & 1023u) != 0u) __trap();          // the layout below assumes a 1024-byte aligned dynamic shared-memory base
  if (tid == 0) {
    for (uint32_t s = 0; s < NSLOT; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, B17_EPI); }
    mbar_init(bar_tpad, B17_EPI); mbar_init(bar_cat, B17_EPI);
    fence_barrier_init();
  }
  if (tid < 32) {
    // I16 as a K-major B operand with 32-byte swizzled rows: 16-byte chunk c of row n at n*32 + ((c ^ ((n >> 2) & 1)) << 4).
    // x * I accumulated in fp32 is exact, so "+ x" costs four N = 16 MMAs per 64-channel panel instead of a residual
    // round trip through the epilogue.
    const uint32_t n = (uint32_t)tid >> 1, cch = (uint32_t)tid & 1u;
    uint4 v = make_uint4(0, 0, 0, 0);
    if ((n >> 3) == cch) {
      const uint32_t one = F16 ? 0x3C00u : 0x3F80u, w = (n & 7u) >> 1, sh = (n & 1u) * 16u;
      if (w == 0) v.x = one << sh; else if (w == 1) v.y = one << sh; else if (w == 2) v.z = one << sh; else v.w = one << sh;
    }
    sts128(sm + OFF_I16 + n * 32u + ((cch ^ ((n >> 2) & 1u)) << 4), v);
    fence_proxy_async_smem();
  }
  if (warp == 16 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w3) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w4) : "memory");
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                  // x is the previous kernel's output
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 16) {
    // =========================================================================================== epilogue warps
    const int q = warp & 3, part = warp >> 2, et = tid;       // TMEM lane quarter, column quarter
    const int r = q * 32 + lane;                         // TMEM lane = accumulator row
    const int g = r >> 3, pos = r & 7;                   // eight-row group (y*2 + img) / position x inside it
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    // t (input of the 1x7): v-order rows, every eight-row group padded to 16 rows [4 zero | 8 | 4 zero] so that a shift along
    // x is a shift of the descriptor start by whole rows inside the group (stride between groups: 2 KiB)
    const uint32_t t_addr = (uint32_t)(g * 16 + 4 + pos) * 128u, t_sw = (uint32_t)((4 + pos) & 7);
    // u (input of the 7x1), cat and the staging panels: plain v-order rows (u behind 6 zero groups of padding)
    const uint32_t my_row = (uint32_t)r * 128u, my_sw = (uint32_t)(r & 7);
    const uint32_t u_addr = (uint32_t)(48 + r) * 128u;
    uint32_t ph_r0 = 0, ph_r1 = 0, pair = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int img0 = tile * 2;
      // ---- E1: GEMM1 accumulator -> cat planes 0,1 | t into the padded groups of tpad
      if (et == 0) tma_store_wait_read();                         // the previous tile's output panels have left tpad
      asm volatile("bar.sync 1, %0;" ::"n"(B17_EPI) : "memory");
      for (int i = et; i < 2 * 16 * 64; i += B17_EPI) {            // zero rows 0-3 and 12-15 of every 16-row group, both planes
        const int pl = i >> 10, rem = i & 1023, grp = rem >> 6, w = rem & 63, row = (w >> 3) < 4 ? (w >> 3) : (w >> 3) + 8;
        sts128(tpad + (uint32_t)pl * 32768u + (uint32_t)(grp * 16 + row) * 128u + (uint32_t)(w & 7) * 16u, make_uint4(0, 0, 0, 0));
      }
      long long t0 = B17_T0();
      mbar_wait(bar_tfull, ph_r0); ph_r0 ^= 1u;
      B17_ACC(8, t0); t0 = B17_T0();
      tc_fence_after();
      {
        // 256 columns: parts 0,1 -> t planes 0,1 (columns 128..255), parts 2,3 -> cat planes 0,1 (columns 0..127)
        const uint32_t col = part < 2 ? 128u + 64u * (uint32_t)part : 64u * (uint32_t)(part - 2);
        const uint32_t dst = part < 2 ? tpad + (uint32_t)part * 32768u + t_addr : cat + (uint32_t)(part - 2) * SLOT + my_row;
        const uint32_t sw = part < 2 ? t_sw : my_sw;
        drain32_to_row<F16>(t_lane, col, p.b1 + col, dst, sw, 0);
        drain32_to_row<F16>(t_lane, col + 32u, p.b1 + col + 32, dst, sw, 4);
      }
      tc_fence_before();
      mbar_arrive(bar_tempty);
      fence_proxy_async_smem();
      mbar_arrive(bar_tpad);
      B17_ACC(9, t0); t0 = B17_T0();
      // ---- E2: 1x7 accumulator -> u into tpad (t is dead: its MMAs have completed), behind / before 6 zero groups
      mbar_wait(bar_tfull + 8, ph_r1); ph_r1 ^= 1u;
      B17_ACC(10, t0); t0 = B17_T0();
      tc_fence_after();
      for (int i = et; i < 2 * 12 * 64; i += B17_EPI) {            // zero the 6 + 6 padding groups of both u planes
        const int pl = i / (12 * 64), rem = i - pl * 12 * 64, grp = rem >> 6, w = rem & 63;
        sts128(tpad + (uint32_t)pl * TPAD_PLANE + (uint32_t)(grp < 6 ? grp : grp + 16) * 1024u + (uint32_t)w * 16u, make_uint4(0, 0, 0, 0));
      }
      drain32_to_row<F16>(t_lane, 256u + 32u * (uint32_t)part, p.b2 + 32 * part, tpad + (uint32_t)(part >> 1) * TPAD_PLANE + u_addr, my_sw,
                          4u * (uint32_t)(part & 1));
      tc_fence_before();
      mbar_arrive(bar_tempty + 8);
      fence_proxy_async_smem();
      mbar_arrive(bar_tpad);
      B17_ACC(11, t0); t0 = B17_T0();
      // ---- E3: 7x1 accumulator -> cat planes 2,3
      mbar_wait(bar_tfull + 8, ph_r1); ph_r1 ^= 1u;
      B17_ACC(12, t0); t0 = B17_T0();
      tc_fence_after();
      drain32_to_row<F16>(t_lane, 384u + 32u * (uint32_t)part, p.b3 + 32 * part, cat + (uint32_t)(2 + (part >> 1)) * SLOT + my_row, my_sw,
                          4u * (uint32_t)(part & 1));
      tc_fence_before();
      mbar_arrive(bar_tempty + 8);
      fence_proxy_async_smem();
      mbar_arrive(bar_cat);
      B17_ACC(13, t0);
      // ---- projection chunks: accumulator (W4 products + x, see the MMA warp) + bias -> ReLU -> staging panel -> TMA store.
      // Panels go in pairs: parts 0,1 take the even panel, parts 2,3 the odd one (32 columns of one row per thread); two pairs
      // of staging slots alternate, so a pair's stores read their panels while the next pair is computed.
      for (int ch = 0; ch < 4; ++ch) {
        const int region = ch & 1;
        t0 = B17_T0();
        if (region == 0) { mbar_wait(bar_tfull, ph_r0); ph_r0 ^= 1u; } else { mbar_wait(bar_tfull + 8, ph_r1); ph_r1 ^= 1u; }
        B17_ACC(14, t0);
        tc_fence_after();
        const int npair = ch < 3 ? 2 : 1;
        for (int m = 0; m < npair; ++m, ++pair) {
          const int odd = part >> 1, ch0 = 4 * (part & 1);          // which panel of the pair, first 16-byte chunk of this thread
          const int pnl = 4 * ch + 2 * m + odd;
          const uint32_t slot0 = 2u * (pair & 1u);
          t0 = B17_T0();
          drain32_to_row<F16>(t_lane, (uint32_t)(region * 256 + 64 * (2 * m + odd) + 8 * ch0), p.b4 + pnl * 64 + 8 * ch0,
                              tpad + (slot0 + (uint32_t)odd) * SLOT + my_row, my_sw, (uint32_t)ch0);
          B17_ACC(16, t0); t0 = B17_T0();
          fence_proxy_async_smem();
          if (et == 0) tma_store_wait_read();                       // the previous pair's stores have read the OTHER two slots
          asm volatile("bar.sync 2, %0;" ::"n"(B17_EPI) : "memory");
          if (et == 0) {
            tma_store_4d(&tm_x, tpad + slot0 * SLOT, (4 * ch + 2 * m) * 64, 0, img0, 0);
            tma_store_4d(&tm_x, tpad + (slot0 + 1) * SLOT, (4 * ch + 2 * m + 1) * 64, 0, img0, 0);
            tma_store_commit();
          }
          B17_ACC(17, t0);
        }
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * region);
      }
    }
    if (et == 0) tma_store_wait_all();
  } else if (warp == 16) {
    // =========================================================================================== operand producer
    if (lane == 0) {
      uint32_t c = 0;
      auto wait_slots = [&](uint32_t n) { for (uint32_t i = 0; i < n; ++i) mbar_wait(bar_empty + 8 * ((c + i) % NSLOT), (((c + i) / NSLOT) & 1u) ^ 1u); };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int img0 = tile * 2;
        for (int kb = 0; kb < KB1; ++kb) {                 // GEMM1: x K block (1 slot) + W1 K block (256 rows = 2 slots)
          wait_slots(3);
          const uint32_t s = c % NSLOT, fb = bar_full + 8 * s;
          mbar_arrive_expect_tx(fb, 3 * SLOT);
          tma_load_4d(ring + s * SLOT, &tm_x, fb, kb * 64, 0, img0, 0);
          tma_load_2d(ring + (s + 1) * SLOT, &tm_w1, fb, kb * 64, 0);
          // the step's bytes are tracked on its first slot; the other slots' "full" barriers still complete once per ring
          // revolution so that every barrier's phase parity stays (slot counter / NSLOT) & 1
          mbar_arrive(fb + 8); mbar_arrive(fb + 16);
          c += 3;
        }
        for (int conv = 0; conv < 2; ++conv)               // 1x7 then 7x1: (tap, 64-channel plane) weight tiles, 1 slot each
          for (int ks = 0; ks < 14; ++ks) {
            wait_slots(1);
            const uint32_t s = c % NSLOT, fb = bar_full + 8 * s;
            mbar_arrive_expect_tx(fb, SLOT);
            tma_load_2d(ring + s * SLOT, conv == 0 ? &tm_w2 : &tm_w3, fb, ks * 64, 0);
            c += 1;
          }
        for (int ch = 0; ch < 4; ++ch) {                   // projection: 256 output channels x 64 K per step (2 slots) ...
          for (int kb = 0; kb < 4; ++kb) {
            wait_slots(2);
            const uint32_t s = c % NSLOT, fb = bar_full + 8 * s;
            mbar_arrive_expect_tx(fb, 2 * SLOT);
            tma_load_2d(ring + s * SLOT, &tm_w4, fb, kb * 64, ch * 256);
            mbar_arrive(fb + 8);
            c += 2;
          }
          for (int j = 0; j < (ch < 3 ? 4 : 2); ++j) {     // ... then the chunk's panels of x (the residual), 1 slot each
            wait_slots(1);
            const uint32_t s = c % NSLOT, fb = bar_full + 8 * s;
            mbar_arrive_expect_tx(fb, SLOT);
            tma_load_4d(ring + s * SLOT, &tm_x, fb, (4 * ch + j) * 64, 0, img0, 0);
            c += 1;
          }
        }
        // 42 + 28 + 46 = 116 slots per tile: four idle slots make it 120 = 20 ring revolutions, so every tile starts on slot 0
        // and no multi-slot step ever wraps around the end of the ring
        for (int i = 0; i < 4; ++i) {
          wait_slots(1);
          mbar_arrive(bar_full + 8 * (c % NSLOT));
          c += 1;
        }
      }
    }
  } else if (warp == 17) {
    // =========================================================================================== MMA issuer
    const uint32_t idesc256 = make_idesc_f16(256, F16 ? 1 : 0), idesc128 = make_idesc_f16(128, F16 ? 1 : 0);
    uint32_t c = 0, ph_e0 = 1, ph_e1 = 1, tl = 0;
    const uint64_t ring_desc = make_sw128_desc(ring), cat_desc = make_sw128_desc(cat), tpad_desc = make_sw128_desc(tpad);
    const uint64_t t_desc = make_sw128_desc_sbo(tpad, 2048u);
    const uint64_t i16_desc = make_sw_desc(sm + OFF_I16, 32, 0);
    const uint32_t idesc16 = make_idesc_f16(16, F16 ? 1 : 0);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
      // ---- GEMM1 -> region 0
      long long m0 = B17_T0();
      mbar_wait(bar_tempty, ph_e0); ph_e0 ^= 1u;
      B17_ACC(6, m0);
      tc_fence_after();
      for (int kb = 0; kb < KB1; ++kb) {
        const uint32_t s = c % NSLOT;
        m0 = B17_T0();
        mbar_wait(bar_full + 8 * s, (c / NSLOT) & 1u);
        B17_ACC(1, m0);
        tc_fence_after();
        const uint64_t a = ring_desc + (uint64_t)(s * (SLOT >> 4)), b = ring_desc + (uint64_t)((s + 1) * (SLOT >> 4));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base, a + (uint64_t)(2 * kk), b + (uint64_t)(2 * kk), idesc256, (kb | kk) != 0);
          umma_commit(bar_empty + 8 * s); umma_commit(bar_empty + 8 * (s + 1)); umma_commit(bar_empty + 8 * (s + 2));
        }
        __syncwarp();
        c += 3;
      }
      if (elect_one()) umma_commit(bar_tfull);
      __syncwarp();
      // ---- 1x7 (t in h-order) -> region 1 columns 256..383, then 7x1 (u in v-order) -> columns 384..511
      for (int conv = 0; conv < 2; ++conv) {
        m0 = B17_T0();
        mbar_wait(bar_tpad, (uint32_t)conv);               // E1 / E2 have written tpad (two completions per tile: parities 0, 1)
        B17_ACC(4, m0); m0 = B17_T0();
        mbar_wait(bar_tempty + 8, ph_e1); ph_e1 ^= 1u;
        B17_ACC(6, m0);
        tc_fence_after();
        const uint32_t d = tmem_base + 256u + 128u * (uint32_t)conv;
        for (int ks = 0; ks < 14; ++ks) {
          const uint32_t s = c % NSLOT;
          m0 = B17_T0();
          mbar_wait(bar_full + 8 * s, (c / NSLOT) & 1u);
          B17_ACC(2, m0);
          tc_fence_after();
          const int tap = ks >> 1, pl = ks & 1;              // plane pl = channels [64 pl, 64 pl + 64)
          // 1x7: shift by (tap - 3) rows inside the 16-row groups (stride 2 KiB); 7x1: shift by 2*tap groups of 1 KiB
          const uint64_t a = conv == 0 ? t_desc + (uint64_t)(((uint32_t)pl * 32768u + (uint32_t)(1 + tap) * 128u) >> 4)
                                       : tpad_desc + (uint64_t)(((uint32_t)pl * TPAD_PLANE + (uint32_t)tap * 2048u) >> 4);
          const uint64_t b = ring_desc + (uint64_t)(s * (SLOT >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d, a + (uint64_t)(2 * kk), b + (uint64_t)(2 * kk), idesc128, (ks | kk) != 0);
            umma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
          c += 1;
        }
        if (elect_one()) umma_commit(bar_tfull + 8);
        __syncwarp();
      }
      // ---- projection: cat (4 K planes) x W4 chunk -> region ch & 1
      m0 = B17_T0();
      mbar_wait(bar_cat, tl & 1u);
      B17_ACC(5, m0);
      for (int ch = 0; ch < 4; ++ch) {
        const int region = ch & 1;
        m0 = B17_T0();
        if (region == 0) { mbar_wait(bar_tempty, ph_e0); ph_e0 ^= 1u; } else { mbar_wait(bar_tempty + 8, ph_e1); ph_e1 ^= 1u; }
        B17_ACC(6, m0);
        tc_fence_after();
        const uint32_t d = tmem_base + 256u * (uint32_t)region;
        const uint32_t idesc = ch < 3 ? idesc256 : idesc128;
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t s = c % NSLOT;
          m0 = B17_T0();
          mbar_wait(bar_full + 8 * s, (c / NSLOT) & 1u);
          B17_ACC(3, m0);
          tc_fence_after();
          const uint64_t a = cat_desc + (uint64_t)((uint32_t)kb * (SLOT >> 4)), b = ring_desc + (uint64_t)(s * (SLOT >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d, a + (uint64_t)(2 * kk), b + (uint64_t)(2 * kk), idesc, (kb | kk) != 0);
            umma_commit(bar_empty + 8 * s); umma_commit(bar_empty + 8 * (s + 1));
          }
          __syncwarp();
          c += 2;
        }
        for (int j = 0; j < (ch < 3 ? 4 : 2); ++j) {       // + x: panel j of the chunk times I16, 16 output columns per MMA
          const uint32_t s = c % NSLOT;
          m0 = B17_T0();
          mbar_wait(bar_full + 8 * s, (c / NSLOT) & 1u);
          B17_ACC(3, m0);
          tc_fence_after();
          const uint64_t a = ring_desc + (uint64_t)(s * (SLOT >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d + (uint32_t)(64 * j + 16 * kk), a + (uint64_t)(2 * kk), i16_desc, idesc16, 1u);
            umma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
          c += 1;
        }
        if (elect_one()) umma_commit(bar_tfull + 8 * region);
        __syncwarp();
      }
      for (int i = 0; i < 4; ++i) {                        // the idle slots that realign the ring (see the producer)
        const uint32_t s = c % NSLOT;
        mbar_wait(bar_full + 8 * s, (c / NSLOT) & 1u);
        if (elect_one()) umma_commit(bar_empty + 8 * s);
        __syncwarp();
        c += 1;
      }
    }
  }

  __syncthreads();
  if (dbg && warp == 0) dbg[0] += clock64() - t_kernel;
  if (warp == 17) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool encode_weight_map(EncodeTiledFn enc, CUtensorMap* tm, const void* w, int k_pad, int rows, int box_rows, CUtensorMapDataType dt) {
  const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)k_pad * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, dt, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern long long g_vnfr_launches;

// debug hook (not part of include/vnfr_b200.h): device buffer of 32 int64 cycle counters, or null to switch off
extern "C" int vnfr_b17_debug(long long* dev_buf) {
  VNFR_CUDA(cudaMemcpyToSymbol(g_b17_dbg, &dev_buf, sizeof(dev_buf)));
  return VNFR_OK;
}

extern "C" int vnfr_block17_prepare(VnfrBlock17Op* op) {
  VNFR_REQUIRE(op != nullptr && op->x != nullptr, "op / x is null");
  VNFR_REQUIRE(op->w1 && op->w2 && op->w3 && op->w4 && op->b1 && op->b2 && op->b3 && op->b4, "null weight / bias pointer");
  VNFR_REQUIRE(op->dtype == 0 || op->dtype == 1, "dtype must be 0 (bf16) or 1 (fp16)");
  VNFR_REQUIRE(op->n_img >= 0 && ((uintptr_t)op->x % 16) == 0, "x must be 16-byte aligned");
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) {
    vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return VNFR_ERR_CUDA;
  }
  const CUtensorMapDataType dt = op->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tm[5];
  {
    // x as (c, x, img, y): the box {64, 8, 2, 8} lands in shared memory as 128 rows in v-order (y, img, x)
    const cuuint64_t dims[4] = {896, 8, (cuuint64_t)(op->n_img > 0 ? op->n_img : 1), 8};
    const cuuint64_t strides[3] = {896 * 2, 64 * 896 * 2, 8 * 896 * 2};
    const cuuint32_t box[4] = {64, 8, 2, 8};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tm[0], dt, 4, op->x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed for x");
      return VNFR_ERR_CUDA;
    }
  }
  if (!encode_weight_map(enc, &tm[1], op->w1, 896, 256, 256, dt) || !encode_weight_map(enc, &tm[2], op->w2, 896, 128, 128, dt) ||
      !encode_weight_map(enc, &tm[3], op->w3, 896, 128, 128, dt) || !encode_weight_map(enc, &tm[4], op->w4, 256, 896, 256, dt)) {
    vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed for a weight matrix");
    return VNFR_ERR_CUDA;
  }
  for (int i = 0; i < 5; ++i) memcpy(op->tmap[i], &tm[i], sizeof(CUtensorMap));
  return VNFR_OK;
}

extern "C" int vnfr_block17_run(const VnfrBlock17Op* op, void* stream) {
  VNFR_REQUIRE(op != nullptr, "op is null");
  if (op->n_img <= 0) return VNFR_OK;
  static VnfrPerDevice once = {};
  if (vnfr_first_on_device(once)) {
    VNFR_CUDA(cudaFuncSetAttribute(block17_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B17_SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(block17_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B17_SMEM));
  }
  B17Params p;
  p.b1 = op->b1; p.b2 = op->b2; p.b3 = op->b3; p.b4 = op->b4;
  p.n_img = op->n_img; p.n_tiles = (op->n_img + 1) / 2;
  CUtensorMap tm[5];
  for (int i = 0; i < 5; ++i) memcpy(&tm[i], op->tmap[i], sizeof(CUtensorMap));
  const int grid = p.n_tiles < 148 ? p.n_tiles : 148;
  if (op->dtype == 1)
    VNFR_CUDA(launch_pdl(block17_fused_kernel<true>, grid, B17_THREADS, B17_SMEM, (cudaStream_t)stream, tm[0], tm[1], tm[2], tm[3], tm[4], p));
  else
    VNFR_CUDA(launch_pdl(block17_fused_kernel<false>, grid, B17_THREADS, B17_SMEM, (cudaStream_t)stream, tm[0], tm[1], tm[2], tm[3], tm[4], p));
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
