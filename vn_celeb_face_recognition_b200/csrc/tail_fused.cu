// Fused tail of the embed -> classify path in ONE cooperative kernel (north star: "the bottleneck, L2-normalise and MLP
// classifier, fused in one kernel"):
//
//   AdaptiveAvgPool2d(1) -> last_linear + last_bn -> F.normalize        (inception_resnet_v1.py:294-302)
//   -> dense_1 + ReLU -> dense_2 -> log_softmax                         (mlp_model.py:10-15)
//   -> argmax, exp(max log-prob), threshold -> label                    (demo_image.py:113-137)
//
// and the (embedding, label, probability) rows are written straight into the caller's all-gather send buffer.
//
// Precision: the predicted LABEL must equal the fp32 reference's, and a random-init classifier separates its top two
// classes by ~1e-3 in log-probability, so the three contractions run on the tensor cores in SPLIT PRECISION: every fp32
// operand x is carried as two fp16 parts (hi = fp16(x), lo = fp16(x - hi)), and each K block issues the three products
// hi*hi, hi*lo, lo*hi as separate tcgen05.mma groups into one fp32 TMEM accumulator (dropped lo*lo term: O(2^-22)).
//
// Structure: persistent cooperative grid (one CTA per SM), phases separated by a grid barrier:
//   INPUT  pool (16-bit NHWC -> fp32 mean) or fp32 rows -> split planes A0
//   for each layer l:  GEMM_l (warp-specialised TMA -> tcgen05 -> TMEM -> fp32 split-K partial tiles)
//                      ROW_l  (one warp per row: sum partials in fixed order + bias, then identity / ReLU / L2-normalise /
//                              log-softmax+argmax; writes the fp32 outputs and the split planes of the next layer)
// The GEMM of a 128-face tile is split over N tiles and K ranges so that ~all 148 SMs have a tile; partial sums are
// combined in a fixed order (deterministic: labels do not change from run to run).
#include "tc_common.cuh"
#include <math_constants.h>

using namespace tc;

namespace {

constexpr int TL_THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer (+ TMEM alloc), warps 2-5: epilogue
constexpr int TL_BLOCK_N = 128;
constexpr int TL_STAGES = 3;
constexpr int TL_STAGE_BYTES = 4 * 16384; // A_hi | A_lo | W_hi | W_lo, each 128 rows x 128 B
constexpr int TL_ROWBUF_FLOATS = 8192;    // per-warp row buffer of the row phases (6 x 32 KB = the pipeline stages they alias)
constexpr int TL_MAX_LAYERS = 3;

struct TailLayer {
  int K, N, N_pad, kb, split_k, n_tiles_n;
  int rowop;                 // 0 identity, 1 ReLU, 2 L2-normalise, 3 log-softmax + argmax
  float* partial;            // [split_k][n_pad][N_pad] fp32
  const float* bias;         // [N_pad]
  float* out_vec;            // nullable: fp32 row output of the row phase (embedding / log-probabilities), first N columns
  int out_vec_pitch;
  __half* a_next;            // nullable: split planes [2][n_pad][N] (hi rows, then lo rows) feeding the next layer
};

struct TailParams {
  int n, n_pad, m_tiles, n_layers;
  int in_mode;               // 0: 16-bit NHWC activations, mean over hw pixels; 1: fp32 rows
  const void* x; int hw, x_pitch, x_f16;
  const float* x_f32; int x_f32_pitch, x_f32_cols;
  __half* a0;                // split planes of layer 0's input [2][n_pad][K0]
  TailLayer L[TL_MAX_LAYERS];
  void* emb_half; int emb_half_f16;     // nullable: 16-bit copy of the L2-normalised row
  long long* label; float* prob;         // nullable: log-softmax row phase outputs
  float* label_f; float* prob_f; int lp_pitch;   // nullable: the same as float columns of the send buffer
  const float* thr_class; float thr; int n_classes;
  float* count_cell; int count_value;   // nullable: receives (float)count_value (the count cell of the send buffer)
  unsigned int* bar;         // grid barrier counter, zero at launch
};

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Grid barrier of the cooperative launch.  Every thread's global writes of the finished phase are ordered before the
// arrival (bar.sync + gpu-scope fence of the arriving thread: cumulativity), generic-proxy writes that the next phase
// reads through TMA are published to the async proxy by the writers themselves (fence.proxy.async.global before this).
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++epoch;
    const unsigned int target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    const uint64_t t0 = globaltimer_ns();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v < target && globaltimer_ns() - t0 > 2000000000ull) __trap();      // never hang the GPU on a protocol bug
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Optional phase timestamps of CTA 0 (tools/tail_probe.py): globaltimer ns at kernel start and after every grid barrier.
__device__ unsigned long long* g_tail_dbg = nullptr;
#define TL_STAMP(i) do { if (g_tail_dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) g_tail_dbg[i] = globaltimer_ns(); } while (0)

__global__ void __launch_bounds__(TL_THREADS, 1)
tail_fused_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_w0,
                  const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_w1,
                  const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_w2, const TailParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + TL_STAGES * TL_STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8u * TL_STAGES, bar_tfull = bars + 16u * TL_STAGES,
                 bar_tempty = bar_tfull + 16u, tmem_slot = bar_tempty + 16u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < TL_STAGES; ++s) { mbar_init(bar_full + 8u * s, 1); mbar_init(bar_empty + 8u * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8u * i, 1); mbar_init(bar_tempty + 8u * i, 128); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w0) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2u * TL_BLOCK_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  unsigned int epoch = 0;
  int stamp = 0;
  TL_STAMP(stamp++);
  const int gthreads = gridDim.x * TL_THREADS, gtid = blockIdx.x * TL_THREADS + tid;

  // ================================================================== INPUT phase: rows -> split planes of layer 0
  {
    const int K0 = p.L[0].K, c8 = K0 >> 3;
    __half* a_hi = p.a0;
    __half* a_lo = p.a0 + (size_t)p.n_pad * K0;
    for (int i = gtid; i < p.n * c8; i += gthreads) {
      const int row = i / c8, cg = i - row * c8;
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.in_mode == 0) {
        const uint16_t* xr = reinterpret_cast<const uint16_t*>(p.x) + (size_t)row * p.hw * p.x_pitch + cg * 8;
        for (int px0 = 0; px0 < p.hw; px0 += 9) {            // nine independent 16-byte loads in flight (hw = 9 for 160 px crops), summed in pixel order
          uint4 v3[9];
#pragma unroll
          for (int u = 0; u < 9; ++u)
            v3[u] = px0 + u < p.hw ? __ldg(reinterpret_cast<const uint4*>(xr + (size_t)(px0 + u) * p.x_pitch)) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int u = 0; u < 9; ++u) {
            if (px0 + u >= p.hw) break;
            const uint32_t w[4] = {v3[u].x, v3[u].y, v3[u].z, v3[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float lo, hi;
              if (p.x_f16) unpack2<true>(w[e], lo, hi); else unpack2<false>(w[e], lo, hi);
              s[2 * e] += lo; s[2 * e + 1] += hi;
            }
          }
        }
        const float d = (float)p.hw;
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = s[e] / d;
      } else {
        const float* xr = p.x_f32 + (size_t)row * p.x_f32_pitch;
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = cg * 8 + e < p.x_f32_cols ? __ldg(xr + cg * 8 + e) : 0.f;     // zero K padding
      }
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half h0 = __float2half_rn(s[2 * e]), h1 = __float2half_rn(s[2 * e + 1]);
        const __half l0 = __float2half_rn(s[2 * e] - __half2float(h0)), l1 = __float2half_rn(s[2 * e + 1] - __half2float(h1));
        ph[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        pl[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint4*>(a_hi + (size_t)row * K0 + cg * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      *reinterpret_cast<uint4*>(a_lo + (size_t)row * K0 + cg * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    if (gtid == 0 && p.count_cell != nullptr) *p.count_cell = (float)p.count_value;
    fence_proxy_async_global();
  }
  grid_sync(p.bar, epoch);
  TL_STAMP(stamp++);

  // pipeline state of the three GEMM roles persists across layers (barrier phases keep running)
  int ps = 0; uint32_t pph = 1;            // producer: stage, "slot free" parity
  int ms = 0; uint32_t mph = 0;            // MMA: stage, "slot full" parity
  int mt_count = 0, et_count = 0;          // accumulator tiles issued / drained by this CTA

  for (int l = 0; l < p.n_layers; ++l) {
    const TailLayer& L = p.L[l];
    const CUtensorMap* tm_a = l == 0 ? &tm_a0 : (l == 1 ? &tm_a1 : &tm_a2);
    const CUtensorMap* tm_w = l == 0 ? &tm_w0 : (l == 1 ? &tm_w1 : &tm_w2);
    const int tiles = p.m_tiles * L.n_tiles_n * L.split_k;
    // ================================================================ GEMM phase
    if (warp == 0) {
      if (lane == 0) {
        fence_proxy_async_global();          // A planes were written through the generic proxy by the previous phase
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
          const int sp = t % L.split_k, rest = t / L.split_k;
          const int nt = rest % L.n_tiles_n, mt = rest / L.n_tiles_n;
          const int kb0 = (sp * L.kb) / L.split_k, kb1 = ((sp + 1) * L.kb) / L.split_k;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(bar_empty + 8u * ps, pph);
            const uint32_t st = smem_base + (uint32_t)ps * TL_STAGE_BYTES, fb = bar_full + 8u * ps;
            mbar_arrive_expect_tx(fb, TL_STAGE_BYTES);
            tma_load_2d(st, tm_a, fb, kb * 64, mt * 128);
            tma_load_2d(st + 16384u, tm_a, fb, kb * 64, p.n_pad + mt * 128);
            tma_load_2d(st + 32768u, tm_w, fb, kb * 64, nt * TL_BLOCK_N);
            tma_load_2d(st + 49152u, tm_w, fb, kb * 64, L.N_pad + nt * TL_BLOCK_N);
            if (++ps == TL_STAGES) { ps = 0; pph ^= 1u; }
          }
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc_f16(TL_BLOCK_N, 1);
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++mt_count) {
        const int sp = t % L.split_k;
        const int kb0 = (sp * L.kb) / L.split_k, kb1 = ((sp + 1) * L.kb) / L.split_k;
        const int ab = mt_count & 1;
        mbar_wait(bar_tempty + 8u * ab, (uint32_t)(((mt_count >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * TL_BLOCK_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8u * ms, mph);
          tc_fence_after();
          const uint32_t st = smem_base + (uint32_t)ms * TL_STAGE_BYTES;
          const uint64_t a_hi = make_sw128_desc(st), a_lo = make_sw128_desc(st + 16384u);
          const uint64_t w_hi = make_sw128_desc(st + 32768u), w_lo = make_sw128_desc(st + 49152u);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + (uint64_t)(2 * kk), w_hi + (uint64_t)(2 * kk), idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_hi + (uint64_t)(2 * kk), w_lo + (uint64_t)(2 * kk), idesc, 1u);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(d_tmem, a_lo + (uint64_t)(2 * kk), w_hi + (uint64_t)(2 * kk), idesc, 1u);
            umma_commit(bar_empty + 8u * ms);
          }
          __syncwarp();
          if (++ms == TL_STAGES) { ms = 0; mph ^= 1u; }
        }
        if (elect_one()) umma_commit(bar_tfull + 8u * ab);
        __syncwarp();
      }
    } else {
      const int q = warp & 3, r = q * 32 + lane;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++et_count) {
        const int sp = t % L.split_k, rest = t / L.split_k;
        const int nt = rest % L.n_tiles_n, mt = rest / L.n_tiles_n;
        const int ab = et_count & 1;
        mbar_wait(bar_tfull + 8u * ab, (uint32_t)((et_count >> 1) & 1));
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * TL_BLOCK_N);
        float* dst = L.partial + ((size_t)sp * p.n_pad + (size_t)(mt * 128 + r)) * L.N_pad + nt * TL_BLOCK_N;
#pragma unroll 1
        for (int c0 = 0; c0 < TL_BLOCK_N; c0 += 32) {
          float v[32];
          __syncwarp();
          tmem_ld16_issue(t_row + (uint32_t)c0, v);
          tmem_ld16_issue(t_row + (uint32_t)(c0 + 16), v + 16);
          tmem_ld_wait(v);
          tmem_ld_wait(v + 16);
          float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        tc_fence_before();
        mbar_arrive(bar_tempty + 8u * ab);
      }
    }
    grid_sync(p.bar, epoch);
    TL_STAMP(stamp++);

    // ================================================================ ROW phase: one warp per row.  A lane owns the column quads
    // q = lane + 32 j (columns 4q..4q+3); quads are processed in batches of 4 per lane (512 columns per warp) with the loads of
    // ALL K splits of a batch in flight at once (the first version's one-load-at-a-time loop was pure L2 latency: 26-68 us per
    // phase), partial sums are added in split order (deterministic).
    {
      float* rowbuf = reinterpret_cast<float*>(smem_al) + warp * TL_ROWBUF_FLOATS;
      const int gwarps = gridDim.x * (TL_THREADS / 32), gwarp = blockIdx.x * (TL_THREADS / 32) + warp;
      const size_t plane = (size_t)p.n_pad * L.N_pad;
      const int n_batches = L.N_pad >> 9, tail_quads = (L.N_pad & 511) >> 7;     // batches of 4 quads per lane + 0..3 single quads
      for (int row = gwarp; row < p.n; row += gwarps) {
        const float* pr = L.partial + (size_t)row * L.N_pad;
        float ss = 0.f, mx = -CUDART_INF_F;
        int arg = 0x7fffffff;
        auto consume = [&](float4 v, int col) {            // bias, activation, row statistics, stash in the row buffer
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(L.bias + col));
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          if (L.rowop == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          *reinterpret_cast<float4*>(rowbuf + col) = v;
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          if (col < L.N && v.x > mx) { mx = v.x; arg = col; }
          if (col + 1 < L.N && v.y > mx) { mx = v.y; arg = col + 1; }
          if (col + 2 < L.N && v.z > mx) { mx = v.z; arg = col + 2; }
          if (col + 3 < L.N && v.w > mx) { mx = v.w; arg = col + 3; }
        };
        for (int bt = 0; bt < n_batches; ++bt) {
          const int col0 = bt * 512 + lane * 4;
          float4 acc[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = __ldcg(reinterpret_cast<const float4*>(pr + col0 + 128 * j));
          for (int sp = 1; sp < L.split_k; ++sp) {
            float4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = __ldcg(reinterpret_cast<const float4*>(pr + sp * plane + col0 + 128 * j));
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[j].x += t[j].x; acc[j].y += t[j].y; acc[j].z += t[j].z; acc[j].w += t[j].w; }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) consume(acc[j], col0 + 128 * j);
        }
        for (int j = 0; j < tail_quads; ++j) {
          const int col = n_batches * 512 + 128 * j + lane * 4;
          float4 acc = __ldcg(reinterpret_cast<const float4*>(pr + col));
          for (int sp = 1; sp < L.split_k; ++sp) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(pr + sp * plane + col));
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          consume(acc, col);
        }
        __syncwarp();
        const int n_quads = L.N_pad >> 7;                   // quads per lane
        __half* hi_row = L.a_next != nullptr ? L.a_next + (size_t)row * L.N : nullptr;
        __half* lo_row = L.a_next != nullptr ? L.a_next + ((size_t)p.n_pad + row) * L.N : nullptr;
        float* ov = L.out_vec != nullptr ? L.out_vec + (size_t)row * L.out_vec_pitch : nullptr;
        float scale = 1.f, shift = 0.f;                     // output = v / scale - shift
        if (L.rowop == 2) {
          // F.normalize(p=2, dim=1, eps=1e-12): x / max(||x||, eps)
          ss = warp_sum_f(ss);
          scale = fmaxf(sqrtf(ss), 1e-12f);
        } else if (L.rowop == 3) {
          // log_softmax + argmax (first maximal index) + exp(max log-prob) + identify_person's threshold
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
          }
          float se = 0.f;
          for (int j = 0; j < n_quads; ++j) {
            const int col = 128 * j + lane * 4;
            const float4 v = *reinterpret_cast<const float4*>(rowbuf + col);
            if (col < L.N) se += expf(v.x - mx);
            if (col + 1 < L.N) se += expf(v.y - mx);
            if (col + 2 < L.N) se += expf(v.z - mx);
            if (col + 3 < L.N) se += expf(v.w - mx);
          }
          se = warp_sum_f(se);
          const float lse = logf(se);
          shift = mx + lse;
          if (lane == 0) {
            const float pb = expf(-lse);
            const float th = p.thr_class != nullptr ? __ldg(p.thr_class + arg) : p.thr;
            const int lab = pb >= th ? arg : p.n_classes;
            if (p.label != nullptr) p.label[row] = lab;
            if (p.prob != nullptr) p.prob[row] = pb;
            if (p.label_f != nullptr) p.label_f[(size_t)row * p.lp_pitch] = (float)lab;
            if (p.prob_f != nullptr) p.prob_f[(size_t)row * p.lp_pitch] = pb;
          }
        }
        if (ov != nullptr || hi_row != nullptr || (L.rowop == 2 && p.emb_half != nullptr)) {
          for (int j = 0; j < n_quads; ++j) {
            const int col = 128 * j + lane * 4;
            if (col >= L.N) break;
            float4 v = *reinterpret_cast<const float4*>(rowbuf + col);
            if (L.rowop == 2) { v.x = v.x / scale; v.y = v.y / scale; v.z = v.z / scale; v.w = v.w / scale; }
            else if (L.rowop == 3) { v.x -= shift; v.y -= shift; v.z -= shift; v.w -= shift; }
            if (ov != nullptr) {                            // rows of `ov` need not be 16-byte aligned (pitch = n_classes)
              ov[col] = v.x;
              if (col + 1 < L.N) ov[col + 1] = v.y;
              if (col + 2 < L.N) ov[col + 2] = v.z;
              if (col + 3 < L.N) ov[col + 3] = v.w;
            }
            if (L.rowop == 2 && p.emb_half != nullptr) {    // N is a multiple of 4 here (embedding width)
              uint2 pk;
              if (p.emb_half_f16) { pk.x = pack2<true>(v.x, v.y); pk.y = pack2<true>(v.z, v.w); }
              else { pk.x = pack2<false>(v.x, v.y); pk.y = pack2<false>(v.z, v.w); }
              *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.emb_half) + (size_t)row * L.N + col) = pk;
            }
            if (hi_row != nullptr) {                        // the next layer's K = N is a multiple of 64: whole quads
              const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
              const __half2 l0 = __floats2half2_rn(v.x - __low2float(h0), v.y - __high2float(h0));
              const __half2 l1 = __floats2half2_rn(v.z - __low2float(h1), v.w - __high2float(h1));
              *reinterpret_cast<uint2*>(hi_row + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
              *reinterpret_cast<uint2*>(lo_row + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
            }
          }
        }
        __syncwarp();
      }
      fence_proxy_async_global();            // the next layer's TMA reads the planes written above
      fence_proxy_async_smem();              // and its TMA writes reuse the shared memory of the row buffers
    }
    if (l + 1 < p.n_layers) grid_sync(p.bar, epoch);
    TL_STAMP(stamp++);
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * TL_BLOCK_N) : "memory");
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

size_t tail_smem_bytes() { return 1024 + (size_t)TL_STAGES * TL_STAGE_BYTES + 16 * TL_STAGES + 48; }

}  // namespace

extern long long g_vnfr_launches;

// debug hook (not part of include/vnfr_b200.h): device buffer of 16 uint64 timestamps, or null to switch off
extern "C" int vnfr_tail_debug(unsigned long long* dev_buf) {
  VNFR_CUDA(cudaMemcpyToSymbol(g_tail_dbg, &dev_buf, sizeof(dev_buf)));
  return VNFR_OK;
}

extern "C" int vnfr_tail_prepare(VnfrTailOp* op) {
  VNFR_REQUIRE(op != nullptr, "op is null");
  VNFR_REQUIRE(op->n_layers >= 1 && op->n_layers <= TL_MAX_LAYERS, "n_layers must be 1..3");
  VNFR_REQUIRE(op->n_pad > 0 && op->n_pad % 128 == 0, "n_pad must be a positive multiple of 128");
  VNFR_REQUIRE(op->in_mode == 0 || op->in_mode == 1, "in_mode must be 0 (pooled 16-bit NHWC) or 1 (fp32 rows)");
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) {
    vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return VNFR_ERR_CUDA;
  }
  for (int l = 0; l < op->n_layers; ++l) {
    const VnfrTailLayer& L = op->layer[l];
    VNFR_REQUIRE(L.K > 0 && L.K % 64 == 0, "layer K must be a multiple of 64");
    VNFR_REQUIRE(L.N > 0 && L.N_pad % TL_BLOCK_N == 0 && L.N_pad >= L.N && L.N <= TL_ROWBUF_FLOATS, "bad layer N / N_pad");
    VNFR_REQUIRE(L.split_k >= 1 && L.split_k <= L.K / 64, "split_k must be in [1, K/64]");
    VNFR_REQUIRE(L.rowop >= 0 && L.rowop <= 3, "rowop must be 0..3");
    VNFR_REQUIRE(L.weights != nullptr && L.bias != nullptr && L.partial != nullptr && L.a_in != nullptr, "null layer buffer");
    VNFR_REQUIRE(l == 0 || (op->layer[l - 1].N == L.K && op->layer[l - 1].a_next == L.a_in), "layer l must read what layer l-1 writes");
    CUtensorMap ta, tw;
    if (!encode_rows(enc, &ta, L.a_in, L.K, 2LL * op->n_pad) || !encode_rows(enc, &tw, L.weights, L.K, 2LL * L.N_pad)) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed");
      return VNFR_ERR_CUDA;
    }
    memcpy(op->tmap_a[l], &ta, sizeof(ta));
    memcpy(op->tmap_w[l], &tw, sizeof(tw));
  }
  return VNFR_OK;
}

extern "C" int vnfr_tail_run(const VnfrTailOp* op, int n, void* stream) {
  VNFR_REQUIRE(op != nullptr, "op is null");
  VNFR_REQUIRE(n >= 0 && n <= op->n_pad, "n exceeds the planned capacity n_pad");
  if (n == 0) return VNFR_OK;
  int dev = 0;
  VNFR_CUDA(cudaGetDevice(&dev));
  static int grid_for_dev[64] = {0};
  VNFR_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  const size_t smem = tail_smem_bytes();
  if (grid_for_dev[dev] == 0) {
    int coop = 0, sms = 0, per_sm = 0;
    VNFR_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    VNFR_REQUIRE(coop != 0, "device does not support cooperative launches");
    VNFR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VNFR_CUDA(cudaFuncSetAttribute(tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VNFR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tail_fused_kernel, TL_THREADS, smem));
    VNFR_REQUIRE(per_sm >= 1, "tail kernel does not fit on an SM");
    grid_for_dev[dev] = sms;
  }
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.n_pad = op->n_pad; p.m_tiles = ceil_div(n, 128); p.n_layers = op->n_layers;
  p.in_mode = op->in_mode;
  p.x = op->x; p.hw = op->hw; p.x_pitch = op->x_pitch; p.x_f16 = op->x_dtype == 1;
  p.x_f32 = op->x_f32; p.x_f32_pitch = op->x_f32_pitch; p.x_f32_cols = op->x_f32_cols;
  VNFR_REQUIRE(p.in_mode == 1 ? (p.x_f32 != nullptr && p.x_f32_cols > 0 && p.x_f32_cols <= op->layer[0].K) : (p.x != nullptr && p.hw > 0 && p.x_pitch % 8 == 0), "bad input");
  p.a0 = (__half*)op->layer[0].a_in;
  int max_tiles = 1;
  for (int l = 0; l < op->n_layers; ++l) {
    const VnfrTailLayer& S = op->layer[l];
    TailLayer& L = p.L[l];
    L.K = S.K; L.N = S.N; L.N_pad = S.N_pad; L.kb = S.K / 64; L.split_k = S.split_k; L.n_tiles_n = S.N_pad / TL_BLOCK_N;
    L.rowop = S.rowop; L.partial = S.partial; L.bias = S.bias; L.out_vec = S.out_vec; L.out_vec_pitch = S.out_vec_pitch;
    L.a_next = (__half*)S.a_next;
    const int tiles = p.m_tiles * L.n_tiles_n * L.split_k;
    if (tiles > max_tiles) max_tiles = tiles;
  }
  p.emb_half = op->emb_half; p.emb_half_f16 = op->emb_half_dtype == 1;
  p.label = (long long*)op->label; p.prob = op->prob;
  p.label_f = op->label_f; p.prob_f = op->prob_f; p.lp_pitch = op->lp_pitch;
  p.thr_class = op->thr_class; p.thr = op->thr; p.n_classes = op->n_classes;
  p.count_cell = op->count_cell; p.count_value = op->count_value;
  p.bar = op->grid_barrier;
  VNFR_REQUIRE(p.bar != nullptr, "grid_barrier is null");
  cudaStream_t st = (cudaStream_t)stream;
  VNFR_CUDA(cudaMemsetAsync(p.bar, 0, sizeof(unsigned int), st));
  CUtensorMap tm[6];
  for (int l = 0; l < TL_MAX_LAYERS; ++l) {
    const int src = l < op->n_layers ? l : 0;
    memcpy(&tm[2 * l], op->tmap_a[src], sizeof(CUtensorMap));
    memcpy(&tm[2 * l + 1], op->tmap_w[src], sizeof(CUtensorMap));
  }
  int grid = grid_for_dev[dev];
  // the row phases need ceil(n / 6) CTAs, the GEMM phases max_tiles: a smaller grid makes the barriers cheaper
  const int want = max_tiles > ceil_div(n, TL_THREADS / 32) ? max_tiles : ceil_div(n, TL_THREADS / 32);
  if (want < grid) grid = want;
  void* args[] = {&tm[0], &tm[1], &tm[2], &tm[3], &tm[4], &tm[5], &p};
  VNFR_CUDA(cudaLaunchCooperativeKernel((const void*)tail_fused_kernel, dim3((unsigned)grid), dim3(TL_THREADS), args, smem, st));
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
