// Cosine top-k of query embeddings against a gallery shard with the top-k FUSED into the score GEMM (BASELINE.json config
// 5 / north star "optional cosine top-k against a sharded gallery"; the reference itself has no gallery search).
//
//   S = Q G^T   (Q: m x 512 unit vectors, G: g x 512 unit vectors, 16-bit, fp32 accumulation)   top-8 per row of S
//
// The score matrix never exists in memory: a CTA keeps its 128-query tile of Q resident in shared memory (128 KB), streams
// gallery tiles of 256 rows through a TMA ring, accumulates 128 x 256 scores in TMEM (two buffers), and the epilogue warps
// scan each accumulator straight from TMEM into a per-row running top-8 held in registers (values + global row indices,
// ties towards the lower index).  Work item = (query tile, gallery split): with few query tiles the gallery is split across
// CTAs so that all SMs are busy; every (split, query) pair writes one top-8 list, merged afterwards (gallery.py).
//
// Warps: 0-7 epilogue (TMEM lane quarter = warp & 3; warps 0-3 scan columns 0..127 of a tile, warps 4-7 columns 128..255),
// 8 TMA producer, 9 MMA issuer + TMEM owner.
#include "tc_common.cuh"
#include <math_constants.h>

using namespace tc;

namespace {

constexpr int GT_THREADS = 320, GT_EPI = 256, GT_K = 8, GT_N = 256, GT_STAGES = 3;
constexpr uint32_t GT_Q_BYTES = 8 * 16384, GT_STAGE_BYTES = GT_N * 128;
constexpr uint32_t GT_OFF_RING = GT_Q_BYTES, GT_OFF_BARS = GT_OFF_RING + GT_STAGES * GT_STAGE_BYTES, GT_OFF_MERGE = GT_OFF_BARS + 128;
constexpr uint32_t GT_SMEM = GT_OFF_MERGE;      // the merge scratch aliases the ring (used after an item's last tile)

struct GtParams {
  int m, g_valid, g_tiles, splits, m_tiles, index_offset;
  float* out_val;        // [splits][m][GT_K]
  int* out_idx;          // [splits][m][GT_K]
};

struct TopK {
  float v[GT_K];
  int ix[GT_K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < GT_K; ++j) { v[j] = -CUDART_INF_F; ix[j] = 0x7fffffff; }
  }
  // candidates arrive in increasing index order per thread, so "strictly greater" keeps the lower index on ties
  __device__ __forceinline__ void push(float x, int i) {
    if (x > v[GT_K - 1]) {
      v[GT_K - 1] = x; ix[GT_K - 1] = i;
#pragma unroll
      for (int j = GT_K - 1; j > 0; --j) {
        if (v[j] > v[j - 1]) { const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv; const int ti = ix[j]; ix[j] = ix[j - 1]; ix[j - 1] = ti; }
      }
    }
  }
  // general insert (merge of two lists: ties broken by index)
  __device__ __forceinline__ void push_tie(float x, int i) {
    if (x > v[GT_K - 1] || (x == v[GT_K - 1] && i < ix[GT_K - 1])) {
      v[GT_K - 1] = x; ix[GT_K - 1] = i;
#pragma unroll
      for (int j = GT_K - 1; j > 0; --j) {
        if (v[j] > v[j - 1] || (v[j] == v[j - 1] && ix[j] < ix[j - 1])) {
          const float tv = v[j]; v[j] = v[j - 1]; v[j - 1] = tv; const int ti = ix[j]; ix[j] = ix[j - 1]; ix[j - 1] = ti;
        }
      }
    }
  }
};

// 32 consecutive scores of one row against the row's running top-8.  The common case -- none of them beats the current 8th
// best -- is decided by ONE comparison of the chunk maximum (a max tree: ~1 instruction per score; the first version's
// compare-and-branch per score cost 7 instructions per score and made the EPILOGUE the bound of the whole search: ncu 3.4 G
// warp instructions per 123 k x 125 k search, the tensor pipe waiting for accumulators to be drained).
__device__ __forceinline__ void scan_chunk(TopK& tk, const float (&v)[32], int gi, int g_valid) {
  if (gi + 32 <= g_valid) {
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3]));
    const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
    if (mx > tk.v[GT_K - 1]) {
#pragma unroll
      for (int j = 0; j < 32; ++j) tk.push(v[j], gi + j);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (gi + j < g_valid) tk.push(v[j], gi + j);
  }
}

template <bool F16>
__global__ void __launch_bounds__(GT_THREADS, 1)
gallery_topk_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_g, const GtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sm = smem_u32(smem_raw);
  const uint32_t q_smem = sm, ring = sm + GT_OFF_RING, bars = sm + GT_OFF_BARS;
  const uint32_t bar_full = bars, bar_empty = bars + 24, bar_tfull = bars + 48, bar_tempty = bars + 64, bar_qfull = bars + 80,
                 bar_qempty = bars + 88, tmem_slot = bars + 96;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((sm & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, GT_EPI); }
    mbar_init(bar_qfull, 1); mbar_init(bar_qempty, 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_g) : "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const int n_items = p.m_tiles * p.splits;

  if (warp < 8) {
    // ================================================= epilogue: TMEM -> running top-8 per (row, column half)
    const int q = warp & 3, half = warp >> 2, r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* mv = reinterpret_cast<float*>(smem_raw + GT_OFF_RING);                 // merge scratch [128][GT_K] values, then indices
    int* mi = reinterpret_cast<int*>(smem_raw + GT_OFF_RING + 128 * GT_K * 4);
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int sp = item / p.m_tiles, mt = item - sp * p.m_tiles;      // split-major: see the producer
      const int t0 = (sp * p.g_tiles) / p.splits, t1 = ((sp + 1) * p.g_tiles) / p.splits;
      TopK tk;
      tk.init();
      for (int t = t0; t < t1; ++t, ++tcount) {
        const uint32_t ab = tcount & 1u;
        mbar_wait(bar_tfull + 8 * ab, (tcount >> 1) & 1u);
        tc_fence_after();
        const int col_base = t * GT_N + half * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float v[32];
          __syncwarp();
          tmem_ld16_issue(t_lane + ab * GT_N + (uint32_t)(half * 128 + c0), v);
          tmem_ld16_issue(t_lane + ab * GT_N + (uint32_t)(half * 128 + c0 + 16), v + 16);
          tmem_ld_wait(v);
          tmem_ld_wait(v + 16);
          scan_chunk(tk, v, col_base + c0, p.g_valid);
        }
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * ab);
      }
      // merge the two column halves of each row through shared memory (the ring is idle: this item's MMAs are complete) and
      // write the row's list
      asm volatile("bar.sync 1, %0;" ::"n"(GT_EPI) : "memory");
      if (half == 1) {
#pragma unroll
        for (int j = 0; j < GT_K; ++j) { mv[r * GT_K + j] = tk.v[j]; mi[r * GT_K + j] = tk.ix[j]; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GT_EPI) : "memory");
      if (half == 0) {
#pragma unroll
        for (int j = 0; j < GT_K; ++j) tk.push_tie(mv[r * GT_K + j], mi[r * GT_K + j]);
        const int row = mt * 128 + r;
        if (row < p.m) {
          float* ov = p.out_val + ((size_t)sp * p.m + row) * GT_K;
          int* oi = p.out_idx + ((size_t)sp * p.m + row) * GT_K;
#pragma unroll
          for (int j = 0; j < GT_K; ++j) { ov[j] = tk.v[j]; oi[j] = tk.ix[j] == 0x7fffffff ? -1 : tk.ix[j] + p.index_offset; }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GT_EPI) : "memory");          // scratch reads done before the producer refills the ring
      if (tid == 0) mbar_arrive(bar_qempty);                               // ... which it may only do after this point
    }
  } else if (warp == 8) {
    // ================================================= TMA producer
    if (lane == 0) {
      uint32_t c = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        // items are numbered split-major: at any time all CTAs stream the SAME gallery range, which then stays in L2 (with
        // one split a 128 MB shard is re-read from HBM by every query tile: the search ran at HBM speed, 21 ms for 123 k x
        // 125 k, whatever the MMA did)
        const int sp = item / p.m_tiles, mt = item - sp * p.m_tiles;
        const int t0 = (sp * p.g_tiles) / p.splits, t1 = ((sp + 1) * p.g_tiles) / p.splits;
        mbar_wait(bar_qempty, (it & 1u) ^ 1u);              // the previous item is finished (Q tile and merge scratch free)
        mbar_arrive_expect_tx(bar_qfull, GT_Q_BYTES);
        for (int kb = 0; kb < 8; ++kb) tma_load_2d(q_smem + kb * 16384, &tm_q, bar_qfull, kb * 64, mt * 128);
        for (int t = t0; t < t1; ++t)
          for (int kb = 0; kb < 8; ++kb, ++c) {
            const uint32_t s = c % GT_STAGES;
            mbar_wait(bar_empty + 8 * s, ((c / GT_STAGES) & 1u) ^ 1u);
            mbar_arrive_expect_tx(bar_full + 8 * s, GT_STAGE_BYTES);
            tma_load_2d(ring + s * GT_STAGE_BYTES, &tm_g, bar_full + 8 * s, kb * 64, t * GT_N);
          }
      }
    }
  } else {
    // ================================================= MMA issuer
    const uint32_t idesc = make_idesc_f16(GT_N, F16 ? 1 : 0);
    const uint64_t q_desc = make_sw128_desc(q_smem), r_desc = make_sw128_desc(ring);
    uint32_t c = 0, tcount = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int sp = item / p.m_tiles;
      const int t0 = (sp * p.g_tiles) / p.splits, t1 = ((sp + 1) * p.g_tiles) / p.splits;
      mbar_wait(bar_qfull, it & 1u);
      tc_fence_after();
      for (int t = t0; t < t1; ++t, ++tcount) {
        const uint32_t ab = tcount & 1u;
        mbar_wait(bar_tempty + 8 * ab, ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb, ++c) {
          const uint32_t s = c % GT_STAGES;
          mbar_wait(bar_full + 8 * s, (c / GT_STAGES) & 1u);
          tc_fence_after();
          const uint64_t a = q_desc + (uint64_t)((uint32_t)kb * 1024u), b = r_desc + (uint64_t)(s * (GT_STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_base + ab * GT_N, a + (uint64_t)(2 * kk), b + (uint64_t)(2 * kk), idesc, (kb | kk) != 0);
            umma_commit(bar_empty + 8 * s);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(bar_tfull + 8 * ab);
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// Two-CTA variant (tcgen05.mma.cta_group::2): a pair of CTAs on the two SMs of a TPC computes a 256 x 256 score tile.  Each CTA
// keeps ITS 128-query tile resident and loads only HALF of every gallery tile (128 rows); the leader CTA (cluster rank 0) issues
// one M = 256 instruction per K step that reads A from both CTAs' shared memory and each half of B from the CTA that holds it,
// so every SM STAGES half the gallery bytes per score (opt-in: see vnfr_gallery_topk for the measurement).
//   * both CTAs' TMA loads signal the LEADER's "full" barrier (cp.async.bulk.tensor ... .cta_group::2 with the barrier address
//     of CTA 0: bit 24 of a shared::cluster address selects the CTA of the pair),
//   * the leader's tcgen05.commit multicasts "stage free" / "accumulator ready" to the barriers of both CTAs,
//   * both CTAs' epilogue threads release an accumulator buffer by arriving on the leader's barrier (shared::cluster arrive).
constexpr uint32_t GT2_STAGE_BYTES = 128 * 128, GT2_STAGES = 6;
constexpr uint32_t GT2_OFF_RING = GT_Q_BYTES, GT2_OFF_BARS = GT2_OFF_RING + GT2_STAGES * GT2_STAGE_BYTES, GT2_SMEM = GT2_OFF_BARS + 256;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-of-the-pair bit of a shared::cluster address: CTA 0 = the leader

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  // (not .aligned: the producer warp's lanes re-converge just before this, the other roles run warp-uniformly)
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {      // arrives on `bar` of BOTH CTAs when the MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // bar: own shared::cta address; the arrival lands in CTA 0
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {     // acquire at cluster scope (peer arrivals)
  const uint64_t t0 = globaltimer_ns();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (globaltimer_ns() - t0 > 2000000000ull) __trap();
  }
}

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT_THREADS, 1)
gallery_topk2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_g, const GtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sm = smem_u32(smem_raw);
  const uint32_t q_smem = sm, ring = sm + GT2_OFF_RING, bars = sm + GT2_OFF_BARS;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * GT2_STAGES, bar_tfull = bars + 16 * GT2_STAGES, bar_tempty = bar_tfull + 16,
                 bar_qfull = bar_tempty + 16, bar_qempty = bar_qfull + 8, tmem_slot = bar_qempty + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if ((sm & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (uint32_t s = 0; s < GT2_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    // accumulator released by the epilogue threads of BOTH CTAs (only the leader's copy of this barrier is used)
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 2 * GT_EPI); }
    mbar_init(bar_qfull, 1); mbar_init(bar_qempty, 1);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_g) : "memory");
  }
  cluster_sync_all();                           // barrier inits of both CTAs are visible before any remote arrive / multicast
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const int m_pairs = (p.m + 255) / 256;
  const int n_items = m_pairs * p.splits;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp < 8) {
    // ================================================= epilogue of this CTA's 128 queries (its own TMEM lanes)
    const int q = warp & 3, half = warp >> 2, r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* mv = reinterpret_cast<float*>(smem_raw + GT2_OFF_RING);
    int* mi = reinterpret_cast<int*>(smem_raw + GT2_OFF_RING + 128 * GT_K * 4);
    uint32_t tcount = 0;
    for (int item = pair; item < n_items; item += n_pairs) {
      const int sp = item / m_pairs, mp = item - sp * m_pairs;          // split-major (all pairs stream the same gallery range)
      const int t0 = (sp * p.g_tiles) / p.splits, t1 = ((sp + 1) * p.g_tiles) / p.splits;
      TopK tk;
      tk.init();
      for (int t = t0; t < t1; ++t, ++tcount) {
        const uint32_t ab = tcount & 1u;
        mbar_wait(bar_tfull + 8 * ab, (tcount >> 1) & 1u);
        tc_fence_after();
        const int col_base = t * GT_N + half * 128;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float v[32];
          __syncwarp();
          tmem_ld16_issue(t_lane + ab * GT_N + (uint32_t)(half * 128 + c0), v);
          tmem_ld16_issue(t_lane + ab * GT_N + (uint32_t)(half * 128 + c0 + 16), v + 16);
          tmem_ld_wait(v);
          tmem_ld_wait(v + 16);
          scan_chunk(tk, v, col_base + c0, p.g_valid);
        }
        tc_fence_before();
        mbar_arrive_leader(bar_tempty + 8 * ab);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GT_EPI) : "memory");
      if (half == 1) {
#pragma unroll
        for (int j = 0; j < GT_K; ++j) { mv[r * GT_K + j] = tk.v[j]; mi[r * GT_K + j] = tk.ix[j]; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GT_EPI) : "memory");
      if (half == 0) {
#pragma unroll
        for (int j = 0; j < GT_K; ++j) tk.push_tie(mv[r * GT_K + j], mi[r * GT_K + j]);
        const int row = mp * 256 + (int)rank * 128 + r;
        if (row < p.m) {
          float* ov = p.out_val + ((size_t)sp * p.m + row) * GT_K;
          int* oi = p.out_idx + ((size_t)sp * p.m + row) * GT_K;
#pragma unroll
          for (int j = 0; j < GT_K; ++j) { ov[j] = tk.v[j]; oi[j] = tk.ix[j] == 0x7fffffff ? -1 : tk.ix[j] + p.index_offset; }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(GT_EPI) : "memory");
      if (tid == 0) mbar_arrive(bar_qempty);             // this CTA's Q tile and merge scratch are free
    }
  } else if (warp == 8) {
    // ================================================= TMA producer: this CTA's query tile and ITS half of every gallery tile
    if (lane == 0) {
      uint32_t c = 0, it = 0;
      const uint32_t lead_full = bar_full & PEER_MASK, lead_qfull = bar_qfull & PEER_MASK;
      for (int item = pair; item < n_items; item += n_pairs, ++it) {
        const int sp = item / m_pairs, mp = item - sp * m_pairs;
        const int t0 = (sp * p.g_tiles) / p.splits, t1 = ((sp + 1) * p.g_tiles) / p.splits;
        mbar_wait(bar_qempty, (it & 1u) ^ 1u);
        if (leader) mbar_arrive_expect_tx(bar_qfull, 2 * GT_Q_BYTES);          // both CTAs' query tiles
        for (int kb = 0; kb < 8; ++kb) tma_load_2d_2sm(q_smem + kb * 16384, &tm_q, lead_qfull, kb * 64, mp * 256 + (int)rank * 128);
        for (int t = t0; t < t1; ++t)
          for (int kb = 0; kb < 8; ++kb, ++c) {
            const uint32_t s = c % GT2_STAGES;
            mbar_wait(bar_empty + 8 * s, ((c / GT2_STAGES) & 1u) ^ 1u);
            if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * GT2_STAGE_BYTES);
            tma_load_2d_2sm(ring + s * GT2_STAGE_BYTES, &tm_g, lead_full + 8 * s, kb * 64, t * GT_N + (int)rank * 128);
          }
      }
    }
    __syncwarp();
  } else if (leader) {
    // ================================================= MMA issuer (leader CTA only): M = 256 over the pair, N = 256
    const uint32_t fmt = F16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(GT_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t q_desc = make_sw128_desc(q_smem), r_desc = make_sw128_desc(ring);
    uint32_t c = 0, tcount = 0, it = 0;
    for (int item = pair; item < n_items; item += n_pairs, ++it) {
      const int sp = item / m_pairs;
      const int t0 = (sp * p.g_tiles) / p.splits, t1 = ((sp + 1) * p.g_tiles) / p.splits;
      mbar_wait(bar_qfull, it & 1u);
      tc_fence_after();
      for (int t = t0; t < t1; ++t, ++tcount) {
        const uint32_t ab = tcount & 1u;
        mbar_wait_cluster(bar_tempty + 8 * ab, ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        for (int kb = 0; kb < 8; ++kb, ++c) {
          const uint32_t s = c % GT2_STAGES;
          mbar_wait(bar_full + 8 * s, (c / GT2_STAGES) & 1u);
          tc_fence_after();
          const uint64_t a = q_desc + (uint64_t)((uint32_t)kb * 1024u), b = r_desc + (uint64_t)(s * (GT2_STAGE_BYTES >> 4));
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_2sm(tmem_base + ab * GT_N, a + (uint64_t)(2 * kk), b + (uint64_t)(2 * kk), idesc, (kb | kk) != 0);
            umma_commit_2sm(bar_empty + 8 * s);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_2sm(bar_tfull + 8 * ab);
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                           // no CTA leaves (or frees TMEM) while its peer may still signal it
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

extern long long g_vnfr_launches;

extern "C" int vnfr_gallery_topk(const void* q, int m, const void* gallery, int g_valid, int g_pad, int dtype, int splits, int index_offset,
                                 float* out_val, int32_t* out_idx, void* stream) {
  VNFR_REQUIRE(q != nullptr && gallery != nullptr && out_val != nullptr && out_idx != nullptr, "null pointer");
  VNFR_REQUIRE(g_pad % GT_N == 0 && g_valid >= 0 && g_valid <= g_pad, "g_pad must be a multiple of 256 covering g_valid");
  VNFR_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp16)");
  VNFR_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)gallery % 16) == 0, "operands must be 16-byte aligned");
  if (m <= 0) return VNFR_OK;
  const int g_tiles = g_pad / GT_N;
  VNFR_REQUIRE(splits >= 1 && (g_tiles == 0 || splits <= g_tiles), "splits must be in [1, gallery tiles]");
  EncodeTiledFn enc = get_encode_tiled();
  if (enc == nullptr) {
    vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return VNFR_ERR_CUDA;
  }
  const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tq, tg;
  const cuuint32_t estr[2] = {1, 1};
  {
    const cuuint64_t dims[2] = {512, (cuuint64_t)m};
    const cuuint64_t strides[1] = {1024};
    const cuuint32_t box[2] = {64, 128};
    if (enc(&tq, dt, 2, const_cast<void*>(q), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed for the queries");
      return VNFR_ERR_CUDA;
    }
  }
  {
    const cuuint64_t dims[2] = {512, (cuuint64_t)(g_pad > 0 ? g_pad : GT_N)};
    const cuuint64_t strides[1] = {1024};
    const cuuint32_t box[2] = {64, GT_N};
    if (enc(&tg, dt, 2, const_cast<void*>(gallery), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed for the gallery");
      return VNFR_ERR_CUDA;
    }
  }
  static VnfrPerDevice once = {};
  if (vnfr_first_on_device(once)) {
    VNFR_CUDA(cudaFuncSetAttribute(gallery_topk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
    VNFR_CUDA(cudaFuncSetAttribute(gallery_topk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM));
  }
  GtParams p;
  p.m = m; p.g_valid = g_valid; p.g_tiles = g_tiles; p.splits = splits; p.m_tiles = ceil_div(m, 128); p.index_offset = index_offset;
  p.out_val = out_val; p.out_idx = out_idx;
  // VNFR_GALLERY_2CTA=1 selects the two-CTA kernel.  Measured on 122 880 queries x 125 000 rows: 14.06 ms against 13.87 ms
  // for the one-CTA kernel (1 120 / 1 136 TFLOP/s) -- the pair stages half the gallery bytes per SM, but each SM's tensor core
  // still reads A and ALL of B (its half locally, the other half from the peer), and this kernel is bound by exactly that
  // operand traffic plus the TMA writes (~80 KB per K step through ~85 B/clk of shared memory), so it stays opt-in.
  static const bool two_cta = getenv("VNFR_GALLERY_2CTA") != nullptr;
  if (two_cta && p.m_tiles >= 2) {
    // two-CTA MMA: pairs of CTAs take 256 queries, each CTA loads half of every gallery tile (box of 128 rows)
    CUtensorMap tg2;
    const cuuint64_t dims[2] = {512, (cuuint64_t)(g_pad > 0 ? g_pad : GT_N)};
    const cuuint64_t strides[1] = {1024};
    const cuuint32_t box[2] = {64, 128};
    if (enc(&tg2, dt, 2, const_cast<void*>(gallery), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      vnfr_set_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed for the gallery (two-CTA box)");
      return VNFR_ERR_CUDA;
    }
    static VnfrPerDevice once2 = {};
    if (vnfr_first_on_device(once2)) {
      VNFR_CUDA(cudaFuncSetAttribute(gallery_topk2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT2_SMEM));
      VNFR_CUDA(cudaFuncSetAttribute(gallery_topk2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT2_SMEM));
    }
    const int items2 = ceil_div(m, 256) * splits;
    const int pairs = items2 < 74 ? items2 : 74;
    if (dtype == 1) gallery_topk2_kernel<true><<<2 * pairs, GT_THREADS, GT2_SMEM, (cudaStream_t)stream>>>(tq, tg2, p);
    else gallery_topk2_kernel<false><<<2 * pairs, GT_THREADS, GT2_SMEM, (cudaStream_t)stream>>>(tq, tg2, p);
    ++g_vnfr_launches;
    VNFR_CHECK_LAUNCH();
    return VNFR_OK;
  }
  const int items = p.m_tiles * splits;
  const int grid = items < 148 ? items : 148;
  if (dtype == 1) gallery_topk_kernel<true><<<grid, GT_THREADS, GT_SMEM, (cudaStream_t)stream>>>(tq, tg, p);
  else gallery_topk_kernel<false><<<grid, GT_THREADS, GT_SMEM, (cudaStream_t)stream>>>(tq, tg, p);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
