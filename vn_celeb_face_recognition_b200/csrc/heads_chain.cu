// Back halves of R-Net and O-Net on the tensor cores in split precision (fp32-level accuracy):
//
//   R-Net: maxpool 3/2 -> conv3 2x2 (48 -> 64) + PReLU -> flatten (W,H,C) -> dense4 (576 -> 128) + PReLU
//          -> {dense5_1 softmax, dense5_2}                                       (mtcnn.py:84-99)
//   O-Net: maxpool 2/2 -> conv4 2x2 (64 -> 128) + PReLU -> flatten (W,H,C) -> dense5 (1152 -> 256) + PReLU
//          -> {dense6_1 softmax, dense6_2, dense6_3}                             (mtcnn.py:138-157)
//
// The FMA-pipe versions (rnet_back_kernel / onet_back_kernel, detect_heads.cu) were bound by shared-memory loads at
// 0.34 ms each per 64-frame batch.  Here every layer is one GEMM over ALL crops of the batch:
//
//   pool_split_kernel   fp32 NHWC conv map -> max-pooled 4x4 map as two fp16 planes (x = hi + lo), rows (crop, y*4+x),
//                       channels padded to 64
//   chain_gemm_kernel   conv 2x2 as four taps: tap (ky,kx) is the SAME plane loaded by TMA from row + 4*ky + kx, so an
//                       accumulator row (crop, y*4+x) is the conv output at (y,x) for y,x < 3 (the other 7 rows of a crop
//                       are dropped); the epilogue adds bias, PReLU and writes the split planes of the dense layer
//                       directly in the reference's (W,H,C) flatten order
//   chain_gemm_kernel   dense layer + PReLU -> split planes
//   chain_gemm_kernel   heads: 6 / 16 outputs per crop, softmax of the first two, written at (image, slot)
//
// Each fp32 operand is two fp16 parts and each K block issues hi*hi, hi*lo, lo*hi into one fp32 TMEM accumulator (as
// tail_fused.cu).  The crop count is read on the device (offs[B], written by scan_counts_kernel).
#include "tc_common.cuh"
#include <math_constants.h>
#include <string.h>

using namespace tc;

namespace {

constexpr int HC_THREADS = 192;           // warp 0: TMA producer, warp 1: MMA issuer (+ TMEM alloc), warps 2-5: epilogue
constexpr int HC_STAGES = 3;
constexpr int HC_STAGE_BYTES = 4 * 16384;  // A_hi | A_lo | W_hi | W_lo, each 128 rows x 128 B

struct ChainGemm {
  const int* n_dev; int n_cap;       // items (crops) = min(*n_dev, n_cap)
  int rows_per_item;                 // 16: rows are (crop, cell of the 4x4 pooled map); 1: one row per crop
  int a_rows_pad;                    // rows of one A plane (the lo plane starts there)
  int n_taps, kc_per_tap;            // K blocks of 64: kb = tap * kc_per_tap + kc
  int tap_shift[4];
  int N, N_pad, n_tiles_n, mma_n;
  const float* bias;                 // [N_pad]
  const float* alpha;                // [N_pad] PReLU slopes (nullable)
  int mode;                          // 0: write split planes, 1: detection heads
  int remap;                         // mode 0: row (crop, y*4+x) -> row crop*9 + x*3 + y (y,x < 3; others dropped)
  __half* a_next; long long next_plane;   // output planes; halves between the hi and the lo plane
  int B, cap; const int* offs; const int4* pad; float* prob; float4* reg; float* lmk;
};

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

__device__ __forceinline__ void split8(const float* s, uint4& hi, uint4& lo) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h = __floats2half2_rn(s[2 * e], s[2 * e + 1]);
    const __half2 l = __floats2half2_rn(s[2 * e] - __low2float(h), s[2 * e + 1] - __high2float(h));
    ph[e] = *reinterpret_cast<const uint32_t*>(&h);
    pl[e] = *reinterpret_cast<const uint32_t*>(&l);
  }
  hi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  lo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// max-pool (every window complete: 9 -> 4 with 3/2, 8 -> 4 with 2/2) of the fp32 NHWC map [crop][IN_W*IN_W][C] + split
template <int POOL, int IN_W, int C>
__global__ void __launch_bounds__(256) pool_split_kernel(const float* __restrict__ src, const int* __restrict__ n_dev, int n_cap,
                                                         __half* __restrict__ a_hi, long long plane) {
  const int n = min(__ldg(n_dev), n_cap);
  const long long total = (long long)n * 16 * 8;
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int cg = (int)(idx & 7);
    const long long row = idx >> 3;
    const int pos = (int)(row & 15), crop = (int)(row >> 4);
    const int py = pos >> 2, px = pos & 3;
    float m[8];
    if (cg * 8 < C) {
      const float* base = src + ((size_t)crop * IN_W * IN_W + (2 * py) * IN_W + 2 * px) * C + cg * 8;
      float4 v[POOL * POOL][2];
#pragma unroll
      for (int ky = 0; ky < POOL; ++ky)
#pragma unroll
        for (int kx = 0; kx < POOL; ++kx) {
          const float4* q = reinterpret_cast<const float4*>(base + (ky * IN_W + kx) * C);
          v[ky * POOL + kx][0] = __ldg(q);
          v[ky * POOL + kx][1] = __ldg(q + 1);
        }
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = -CUDART_INF_F;
#pragma unroll
      for (int t = 0; t < POOL * POOL; ++t) {
        m[0] = fmaxf(m[0], v[t][0].x); m[1] = fmaxf(m[1], v[t][0].y); m[2] = fmaxf(m[2], v[t][0].z); m[3] = fmaxf(m[3], v[t][0].w);
        m[4] = fmaxf(m[4], v[t][1].x); m[5] = fmaxf(m[5], v[t][1].y); m[6] = fmaxf(m[6], v[t][1].z); m[7] = fmaxf(m[7], v[t][1].w);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = 0.f;                  // channel padding of the K = 64 block
    }
    uint4 hi, lo;
    split8(m, hi, lo);
    __half* d = a_hi + (size_t)row * 64 + cg * 8;
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + plane) = lo;
  }
}

__global__ void __launch_bounds__(HC_THREADS, 1)
chain_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const ChainGemm p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + HC_STAGES * HC_STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * HC_STAGES, bar_tfull = bars + 16u * HC_STAGES,
                 bar_tempty = bar_tfull + 16u, tmem_slot = bar_tempty + 16u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < HC_STAGES; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, 128); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n = min(__ldg(p.n_dev), p.n_cap);
  const int rows = n * p.rows_per_item;
  const int tiles = ((rows + 127) >> 7) * p.n_tiles_n;
  const int kb_total = p.n_taps * p.kc_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      int ps = 0; uint32_t pph = 1;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int nt = t % p.n_tiles_n, mt = t / p.n_tiles_n;
        for (int kb = 0; kb < kb_total; ++kb) {
          const int tap = kb / p.kc_per_tap, kc = kb - tap * p.kc_per_tap;
          const int arow = mt * 128 + p.tap_shift[tap];
          mbar_wait(bar_empty + 8u * ps, pph);
          const uint32_t st = smem_base + (uint32_t)ps * HC_STAGE_BYTES, fb = bar_full + 8u * ps;
          mbar_arrive_expect_tx(fb, HC_STAGE_BYTES);
          tma_load_2d(st, &tm_a, fb, kc * 64, arow);
          tma_load_2d(st + 16384u, &tm_a, fb, kc * 64, p.a_rows_pad + arow);
          tma_load_2d(st + 32768u, &tm_w, fb, kb * 64, nt * 128);
          tma_load_2d(st + 49152u, &tm_w, fb, kb * 64, p.N_pad + nt * 128);
          if (++ps == HC_STAGES) { ps = 0; pph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(p.mma_n, 1);
    int ms = 0; uint32_t mph = 0;
    int mt_count = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++mt_count) {
      const int ab = mt_count & 1;
      mbar_wait(bar_tempty + 8u * ab, (uint32_t)(((mt_count >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 128);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(bar_full + 8u * ms, mph);
        tc_fence_after();
        const uint32_t st = smem_base + (uint32_t)ms * HC_STAGE_BYTES;
        const uint64_t a_hi = make_sw128_desc(st), a_lo = make_sw128_desc(st + 16384u);
        const uint64_t w_hi = make_sw128_desc(st + 32768u), w_lo = make_sw128_desc(st + 49152u);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + (uint64_t)(2 * kk), w_hi + (uint64_t)(2 * kk), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + (uint64_t)(2 * kk), w_lo + (uint64_t)(2 * kk), idesc, 1u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_lo + (uint64_t)(2 * kk), w_hi + (uint64_t)(2 * kk), idesc, 1u);
          umma_commit(bar_empty + 8u * ms);
        }
        __syncwarp();
        if (++ms == HC_STAGES) { ms = 0; mph ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_tfull + 8u * ab);
      __syncwarp();
    }
  } else {
    const int q = warp & 3, r = q * 32 + lane;
    int et_count = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++et_count) {
      const int nt = t % p.n_tiles_n, mt = t / p.n_tiles_n;
      const int ab = et_count & 1;
      mbar_wait(bar_tfull + 8u * ab, (uint32_t)((et_count >> 1) & 1));
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 128);
      const int R = mt * 128 + r;
      if (p.mode == 0) {
        bool valid;
        size_t drow;
        if (p.remap) {
          const int crop = R >> 4, y = (R >> 2) & 3, x = R & 3;
          valid = crop < n && y < 3 && x < 3;
          drow = (size_t)crop * 9 + x * 3 + y;
        } else {
          valid = R < n;
          drow = (size_t)R;
        }
        __half* d_hi = p.a_next + drow * p.N + nt * 128;
        const int ncols = min(p.N - nt * 128, 128);
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          float v[32];
          __syncwarp();
          tmem_ld16_issue(t_row + (uint32_t)c0, v);
          tmem_ld16_issue(t_row + (uint32_t)(c0 + 16), v + 16);
          tmem_ld_wait(v);
          tmem_ld_wait(v + 16);
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + nt * 128 + c0);
          const float4* a4 = reinterpret_cast<const float4*>(p.alpha + nt * 128 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(b4 + j);
            v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
            if (p.alpha != nullptr) {
              const float4 aa = __ldg(a4 + j);
              v[4 * j] = v[4 * j] > 0.f ? v[4 * j] : v[4 * j] * aa.x;
              v[4 * j + 1] = v[4 * j + 1] > 0.f ? v[4 * j + 1] : v[4 * j + 1] * aa.y;
              v[4 * j + 2] = v[4 * j + 2] > 0.f ? v[4 * j + 2] : v[4 * j + 2] * aa.z;
              v[4 * j + 3] = v[4 * j + 3] > 0.f ? v[4 * j + 3] : v[4 * j + 3] * aa.w;
            }
          }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 hi, lo;
              split8(v + 8 * j, hi, lo);
              *reinterpret_cast<uint4*>(d_hi + c0 + 8 * j) = hi;
              *reinterpret_cast<uint4*>(d_hi + p.next_plane + c0 + 8 * j) = lo;
            }
          }
        }
      } else {
        float v[16];
        __syncwarp();
        tmem_ld16(t_row, v);
        if (R < n) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + j);
          int b, slot;
          locate(p.offs, p.B, R, b, slot);
          const size_t o = (size_t)b * p.cap + slot;
          const int4 pd = p.pad[o];
          const bool empty = !(pd.w > pd.y - 1 && pd.z > pd.x - 1);      // detect_face.py:110 / :138 would skip this crop
          const float mx = fmaxf(v[0], v[1]);
          const float e0 = expf(v[0] - mx), e1 = expf(v[1] - mx);
          p.prob[o] = empty ? 0.f : e1 / (e0 + e1);
          p.reg[o] = make_float4(v[2], v[3], v[4], v[5]);
          if (p.lmk != nullptr) {
#pragma unroll
            for (int j = 0; j < 10; ++j) p.lmk[o * 10 + j] = v[6 + j];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8u * ab);
    }
    if (p.mode == 0) asm volatile("fence.proxy.async.global;" ::: "memory");      // the next GEMM reads these planes through TMA
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

bool encode_rows(EncodeTiledFn enc, CUtensorMap* tm, const void* base, int K, long long rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr size_t HC_SMEM = 1024 + (size_t)HC_STAGES * HC_STAGE_BYTES + 16 * HC_STAGES + 48;

struct ChainDims { int c_in, n1, n2, n3; };
constexpr ChainDims R_DIMS = {48, 64, 128, 6}, O_DIMS = {64, 128, 256, 16};

inline long long round_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

// tensor maps of one (weights, workspace, capacity) combination; a few are cached because the cascade alternates between
// the workspaces of its sub-batches
struct ChainMaps {
  const void* key_w[3]; const void* key_ws; int key_cap; bool onet; bool used;
  CUtensorMap a[3], w[3];
};
ChainMaps g_maps[8];
int g_maps_next = 0;

}  // namespace

extern long long g_vnfr_launches;

extern "C" long long vnfr_heads_back_workspace_bytes(int onet, int crop_cap) {
  const ChainDims d = onet ? O_DIMS : R_DIMS;
  const long long cp = round_up_ll(crop_cap > 0 ? crop_cap : 1, 128);
  return cp * (2LL * 16 * 64 * 2 + 2LL * 9 * d.n1 * 2 + 2LL * d.n2 * 2);
}

// The back half of R-Net (onet = 0; conv_map = conv2 + PReLU output, fp32 [crop][81][48]) or O-Net (onet = 1; conv_map =
// conv3 + PReLU output, fp32 [crop][64][64]) for the crops [0, min(offs[B], crop_cap)).  Internal: called by
// vnfr_rnet_forward_tc / vnfr_onet_forward_tc when a VnfrHeadsBack is passed.
int vnfr_heads_back_run(int onet, const VnfrHeadsBack* hb, const float* conv_map, int B, int cap, const int32_t* offs,
                        const int32_t* pad, float* prob, float* reg, float* lmk, int crop_cap, void* stream) {
  VNFR_REQUIRE(hb != nullptr && conv_map != nullptr && offs != nullptr && pad != nullptr && prob != nullptr && reg != nullptr, "null pointer");
  VNFR_REQUIRE(!onet || lmk != nullptr, "O-Net needs a landmark buffer");
  VNFR_REQUIRE(hb->planes != nullptr && ((uintptr_t)hb->planes % 1024) == 0, "heads-back workspace must be 1024-byte aligned");
  for (int l = 0; l < 3; ++l) VNFR_REQUIRE(hb->w[l] != nullptr && hb->bias[l] != nullptr, "null heads-back weights");
  VNFR_REQUIRE(hb->alpha[0] != nullptr && hb->alpha[1] != nullptr, "null heads-back PReLU slopes");
  const ChainDims d = onet ? O_DIMS : R_DIMS;
  const long long cp = round_up_ll(crop_cap, 128);
  const int np1 = 128, np2 = d.n2 <= 128 ? 128 : 256, np3 = 128;
  __half* a1 = reinterpret_cast<__half*>(hb->planes);
  __half* a2 = a1 + 2 * cp * 16 * 64;
  __half* a3 = a2 + 2 * cp * 9 * d.n1;
  ChainMaps* m = nullptr;
  for (int i = 0; i < 8; ++i) {
    ChainMaps& c = g_maps[i];
    if (c.used && c.onet == (onet != 0) && c.key_ws == hb->planes && c.key_cap == crop_cap && c.key_w[0] == hb->w[0] &&
        c.key_w[1] == hb->w[1] && c.key_w[2] == hb->w[2]) { m = &c; break; }
  }
  if (m == nullptr) {
    EncodeTiledFn enc = get_encode_tiled();
    if (enc == nullptr) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
      return VNFR_ERR_CUDA;
    }
    m = &g_maps[g_maps_next];
    g_maps_next = (g_maps_next + 1) & 7;
    m->used = false;
    const bool ok = encode_rows(enc, &m->a[0], a1, 64, 2 * cp * 16) && encode_rows(enc, &m->w[0], hb->w[0], 256, 2LL * np1) &&
                    encode_rows(enc, &m->a[1], a2, 9 * d.n1, 2 * cp) && encode_rows(enc, &m->w[1], hb->w[1], 9 * d.n1, 2LL * np2) &&
                    encode_rows(enc, &m->a[2], a3, d.n2, 2 * cp) && encode_rows(enc, &m->w[2], hb->w[2], d.n2, 2LL * np3);
    if (!ok) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed");
      return VNFR_ERR_CUDA;
    }
    m->used = true; m->onet = onet != 0; m->key_ws = hb->planes; m->key_cap = crop_cap;
    for (int l = 0; l < 3; ++l) m->key_w[l] = hb->w[l];
  }
  static VnfrPerDevice attr_once = {};
  if (vnfr_first_on_device(attr_once))
    VNFR_CUDA(cudaFuncSetAttribute(chain_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SMEM));
  cudaStream_t st = (cudaStream_t)stream;
  const int* n_dev = offs + B;
  if (onet) pool_split_kernel<2, 8, 64><<<148 * 8, 256, 0, st>>>(conv_map, n_dev, crop_cap, a1, cp * 16 * 64);
  else pool_split_kernel<3, 9, 48><<<148 * 8, 256, 0, st>>>(conv_map, n_dev, crop_cap, a1, cp * 16 * 64);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  ChainGemm g;
  memset(&g, 0, sizeof(g));
  g.n_dev = n_dev; g.n_cap = crop_cap; g.B = B; g.cap = cap; g.offs = offs; g.pad = reinterpret_cast<const int4*>(pad);
  g.prob = prob; g.reg = reinterpret_cast<float4*>(reg); g.lmk = onet ? lmk : nullptr;
  // conv 2x2 over the pooled 4x4 map
  g.rows_per_item = 16; g.a_rows_pad = (int)(cp * 16); g.n_taps = 4; g.kc_per_tap = 1;
  g.tap_shift[0] = 0; g.tap_shift[1] = 1; g.tap_shift[2] = 4; g.tap_shift[3] = 5;
  g.N = d.n1; g.N_pad = np1; g.n_tiles_n = 1; g.mma_n = d.n1; g.bias = hb->bias[0]; g.alpha = hb->alpha[0];
  g.mode = 0; g.remap = 1; g.a_next = a2; g.next_plane = cp * 9 * d.n1;
  chain_gemm_kernel<<<148, HC_THREADS, HC_SMEM, st>>>(m->a[0], m->w[0], g);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  // dense layer
  g.rows_per_item = 1; g.a_rows_pad = (int)cp; g.n_taps = 1; g.kc_per_tap = 9 * d.n1 / 64;
  g.tap_shift[0] = g.tap_shift[1] = g.tap_shift[2] = g.tap_shift[3] = 0;
  g.N = d.n2; g.N_pad = np2; g.n_tiles_n = np2 / 128; g.mma_n = 128; g.bias = hb->bias[1]; g.alpha = hb->alpha[1];
  g.mode = 0; g.remap = 0; g.a_next = a3; g.next_plane = cp * d.n2;
  chain_gemm_kernel<<<148, HC_THREADS, HC_SMEM, st>>>(m->a[1], m->w[1], g);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  // heads
  g.kc_per_tap = d.n2 / 64;
  g.N = d.n3; g.N_pad = np3; g.n_tiles_n = 1; g.mma_n = 16; g.bias = hb->bias[2]; g.alpha = nullptr;
  g.mode = 1; g.a_next = nullptr; g.next_plane = 0;
  chain_gemm_kernel<<<148, HC_THREADS, HC_SMEM, st>>>(m->a[2], m->w[2], g);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
