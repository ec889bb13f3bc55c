// Block-level building blocks shared by every NMS stage of the detector: a bitonic key/value sort in shared memory and
// a greedy NMS over the sorted boxes that walks the list in chunks of 64 with 64-bit suppression masks in shared
// memory.  Visit order and arithmetic follow torchvision.ops.nms (mode 0; detect_face.py:79, :93, :128) and the
// reference's nms_numpy(..., "Min") (mode 1; detect_face.py:221-257) exactly, so that with identical scores and boxes the
// keep list is bit-identical.  One documented difference: torchvision.ops.batched_nms / batched_nms_numpy (detect_face.py:
// 259-274) separate the images of a batch by ADDING idx*(max_coord+1) to every box in fp32 before the IoU arithmetic, which
// rounds the fractional stage-2/3 coordinates of images idx > 0 to the ulp of ~1e5 (2^-7); here every image is its own
// segment and its boxes are used un-offset (= what the reference computes for image 0 / un-batched calls).  A decision
// sitting within 2^-7 px of the 0.7 IoU / "Min" threshold can therefore differ for the non-first images of a batch.
//   mode 0: area (x2-x1)*(y2-y1), inter = max(0,dx)*max(0,dy), suppress iff inter/(a_i + a_j - inter) > thr
//   mode 1: area (x2-x1+1)*(y2-y1+1), inter with +1, suppress iff inter/min(a_i, a_j) > thr
#pragma once
#include "common.cuh"

// Monotone map float -> uint32 (ascending), valid for negative values too.
__device__ __forceinline__ uint32_t float_sortable(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// Sort key: score descending, then `tiebreak` ascending.
__device__ __forceinline__ unsigned long long nms_key(float score, uint32_t tiebreak) {
  return ((unsigned long long)(~float_sortable(score)) << 32) | tiebreak;
}

// Ascending bitonic sort of np2 (power of two) keys with a 32-bit payload, all threads of the block participate.
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* key, uint32_t* val, int np2) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // lower index of the pair
        const int l = i | j;
        const bool up = (i & k) == 0;
        const unsigned long long a = key[i], b = key[l];
        if ((a > b) == up) {
          key[i] = b; key[l] = a;
          const uint32_t va = val[i]; val[i] = val[l]; val[l] = va;
        }
      }
      __syncthreads();
    }
  }
}

template <int MODE>
__device__ __forceinline__ float nms_area(const float4 b) {
  if (MODE == 0) return mul_rn(sub_rn(b.z, b.x), sub_rn(b.w, b.y));
  return mul_rn(add_rn(sub_rn(b.z, b.x), 1.0f), add_rn(sub_rn(b.w, b.y), 1.0f));
}

template <int MODE>
__device__ __forceinline__ bool nms_suppresses(const float4 a, const float area_a, const float4 b, const float area_b,
                                               const float thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  if (MODE == 0) {
    const float w = fmaxf(0.0f, sub_rn(xx2, xx1)), h = fmaxf(0.0f, sub_rn(yy2, yy1));
    const float inter = mul_rn(w, h);
    return div_rn(inter, sub_rn(add_rn(area_a, area_b), inter)) > thr;
  } else {
    const float w = fmaxf(0.0f, add_rn(sub_rn(xx2, xx1), 1.0f)), h = fmaxf(0.0f, add_rn(sub_rn(yy2, yy1), 1.0f));
    const float inter = mul_rn(w, h);
    return div_rn(inter, fminf(area_a, area_b)) > thr;
  }
}

struct NmsScratch {
  unsigned long long mask[64];
  unsigned long long dead;
  int nkept;
};

// Greedy NMS over boxes sb[0..n) already in visit order (areas in sa[]).  kept[] receives the positions (into sb) of the
// kept boxes in visit order; returns their number (valid on every thread).  All threads of the block must call this.
template <int MODE>
__device__ __forceinline__ int block_nms_sorted(const float4* sb, const float* sa, int n, float thr, int* kept, NmsScratch* sc) {
  if (threadIdx.x == 0) sc->nkept = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += 64) {
    const int m = min(64, n - c0);
    if (threadIdx.x < 64) sc->mask[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) sc->dead = 0ull;
    __syncthreads();
    const int nk = sc->nkept;
    // A: suppression of this chunk by boxes kept in earlier chunks
    for (int idx = threadIdx.x; idx < 64 * nk; idx += blockDim.x) {
      const int j = idx & 63, k = idx >> 6;
      if (j < m) {
        const int pk = kept[k];
        if (nms_suppresses<MODE>(sb[pk], sa[pk], sb[c0 + j], sa[c0 + j], thr)) atomicOr(&sc->dead, 1ull << j);
      }
    }
    // B: pairwise mask inside the chunk (row i suppresses columns j > i)
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
      const int i = idx >> 6, j = idx & 63;
      if (i < j && j < m) {
        if (nms_suppresses<MODE>(sb[c0 + i], sa[c0 + i], sb[c0 + j], sa[c0 + j], thr)) atomicOr(&sc->mask[i], 1ull << j);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long alive = ~sc->dead;
      if (m < 64) alive &= (1ull << m) - 1ull;
      int k = nk;
      for (int j = 0; j < m; ++j) {
        if ((alive >> j) & 1ull) {
          kept[k++] = c0 + j;
          alive &= ~sc->mask[j];
        }
      }
      sc->nkept = k;
    }
    __syncthreads();
  }
  return sc->nkept;
}

__host__ __device__ __forceinline__ int next_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}
