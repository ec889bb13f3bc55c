// The box bookkeeping between the three MTCNN networks, one CTA per image (or per image x pyramid level), entirely on
// the device: score filters, sorts, the four NMS passes, box regression, square-up, truncation/clamping and landmark
// mapping.  Replaces detect_face.py:75-104 (stage 1 tail), :119-131 (stage 2 tail), :148-169 (stage 3 tail) including
// the host-side NumPy NMS (:221-274) and every .cpu()/nonzero() synchronisation in between.
//
// Ordering contract: every NMS visits boxes in score-descending order and ties are broken by
// the position the box had in the previous stage's output, exactly as the reference's stable sorts do; the final
// "Min" NMS breaks ties towards the later candidate (ascending stable argsort read from the end).  Box arithmetic uses
// individually rounded fp32 operations in the reference's order (no FMA contraction).
#include "nms.cuh"

extern long long g_vnfr_launches;

namespace {

constexpr int NTH = 512;

struct LevelTable {
  int n_levels;
  float scale[VNFR_MAX_LEVELS];
};

// generateBoundingBox, detect_face.py:212-217: q1 = floor((2*idx + 1)/scale), q2 = floor((2*idx + 12)/scale); the
// division is an fp32 division by (float)scale, which is what torch's CPU kernel computes for `tensor / python_float`.
__device__ __forceinline__ float4 cell_box(uint32_t cell, float scale) {
  const float x = (float)(cell & 0xFFFFu), y = (float)(cell >> 16);
  const float bx = mul_rn(2.0f, x), by = mul_rn(2.0f, y);
  return make_float4(floorf(div_rn(add_rn(bx, 1.0f), scale)), floorf(div_rn(add_rn(by, 1.0f), scale)),
                     floorf(div_rn(add_rn(bx, 12.0f), scale)), floorf(div_rn(add_rn(by, 12.0f), scale)));
}

// rerec, detect_face.py:292-301
__device__ __forceinline__ float4 rerec(float4 b) {
  const float h = sub_rn(b.w, b.y), w = sub_rn(b.z, b.x);
  const float l = fmaxf(w, h);
  const float x1 = sub_rn(add_rn(b.x, mul_rn(w, 0.5f)), mul_rn(l, 0.5f));
  const float y1 = sub_rn(add_rn(b.y, mul_rn(h, 0.5f)), mul_rn(l, 0.5f));
  return make_float4(x1, y1, add_rn(x1, l), add_rn(y1, l));
}

// pad, detect_face.py:277-289: trunc -> int32, clamp.  Returns (x, y, ex, ey).
__device__ __forceinline__ int4 pad_box(float4 b, int W, int H) {
  int x = (int)truncf(b.x), y = (int)truncf(b.y), ex = (int)truncf(b.z), ey = (int)truncf(b.w);
  if (x < 1) x = 1;
  if (y < 1) y = 1;
  if (ex > W) ex = W;
  if (ey > H) ey = H;
  return make_int4(x, y, ex, ey);
}

// bbreg, detect_face.py:188-200 (+1 widths)
__device__ __forceinline__ float4 bbreg(float4 b, float4 r) {
  const float w = add_rn(sub_rn(b.z, b.x), 1.0f), h = add_rn(sub_rn(b.w, b.y), 1.0f);
  return make_float4(add_rn(b.x, mul_rn(r.x, w)), add_rn(b.y, mul_rn(r.y, h)), add_rn(b.z, mul_rn(r.z, w)),
                     add_rn(b.w, mul_rn(r.w, h)));
}

struct SortSmem {
  unsigned long long* key;
  uint32_t* val;
  float4* sb;
  float* sa;
  int* kept;
};
__device__ __forceinline__ SortSmem carve(unsigned char* smem, int np2) {
  SortSmem s;
  s.key = reinterpret_cast<unsigned long long*>(smem);
  s.sb = reinterpret_cast<float4*>(s.key + np2);
  s.sa = reinterpret_cast<float*>(s.sb + np2);
  s.val = reinterpret_cast<uint32_t*>(s.sa + np2);
  s.kept = reinterpret_cast<int*>(s.val + np2);
  return s;
}
inline size_t sort_smem_bytes(int cap) { return (size_t)next_pow2(cap < 2 ? 2 : cap) * (8 + 16 + 4 + 4 + 4); }

// ---- stage 1a: NMS(0.5) inside each (image, level) segment, detect_face.py:79.  Candidates were appended in arbitrary
// order by the P-Net kernel, so ties are broken by raster cell order (= the reference's nonzero() order).
__global__ void __launch_bounds__(NTH) stage1_level_nms_kernel(const LevelTable lt, int cap1, const int* __restrict__ cand_count,
                                                               const uint32_t* __restrict__ cand_cell,
                                                               const float* __restrict__ cand_score, int* __restrict__ keep_count,
                                                               int* __restrict__ keep, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ NmsScratch sc;
  const int seg = blockIdx.x;
  const int l = seg % lt.n_levels;
  const int raw = cand_count[seg];
  const int n = min(raw, cap1);
  if (raw > cap1 && threadIdx.x == 0) atomicOr(status, 1);
  const int np2 = next_pow2(max(n, 2));
  SortSmem s = carve(smem, np2);
  const size_t base = (size_t)seg * cap1;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    if (i < n) { s.key[i] = nms_key(cand_score[base + i], cand_cell[base + i]); s.val[i] = (uint32_t)i; }
    else { s.key[i] = ~0ull; s.val[i] = 0xFFFFFFFFu; }
  }
  __syncthreads();
  block_bitonic_sort(s.key, s.val, np2);
  const float scale = lt.scale[l];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 b = cell_box(cand_cell[base + s.val[i]], scale);
    s.sb[i] = b;
    s.sa[i] = nms_area<0>(b);
  }
  __syncthreads();
  const int nk = block_nms_sorted<0>(s.sb, s.sa, n, 0.5f, s.kept, &sc);
  for (int i = threadIdx.x; i < nk; i += blockDim.x) keep[base + i] = (int)s.val[s.kept[i]];
  if (threadIdx.x == 0) keep_count[seg] = nk;
}

// ---- stage 1b: per image, concatenate the per-level survivors (level-major, each level in score order), NMS(0.7)
// (detect_face.py:83-94), then regression without +1 (:96-102), rerec (:103) and pad (:104).
__global__ void __launch_bounds__(NTH) stage1_image_kernel(const LevelTable lt, int W, int H, int cap1,
                                                           const uint32_t* __restrict__ cand_cell,
                                                           const float* __restrict__ cand_score, const float4* __restrict__ cand_reg,
                                                           const int* __restrict__ keep_count, const int* __restrict__ keep, int cap2,
                                                           int* __restrict__ s2_count, float4* __restrict__ s2_box,
                                                           int4* __restrict__ s2_pad, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ NmsScratch sc;
  __shared__ int s_off[VNFR_MAX_LEVELS + 1];
  const int b = blockIdx.x;
  const int L = lt.n_levels;
  if (threadIdx.x == 0) {
    int t = 0;
    for (int l = 0; l < L; ++l) { s_off[l] = t; t += keep_count[b * L + l]; }
    s_off[L] = t;
    if (t > cap2) atomicOr(status, 2);
  }
  __syncthreads();
  const int n = min(s_off[L], cap2);
  const int np2 = next_pow2(max(n, 2));
  SortSmem s = carve(smem, np2);
  // val = (level << 24 | slot) is not enough for cap1 up to 2^24?  cap1 <= 65536 is enforced on the host: 16 bits slot.
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    if (i < n) {
      int l = 0;
      while (i >= s_off[l + 1]) ++l;
      const int slot = keep[(size_t)(b * L + l) * cap1 + (i - s_off[l])];
      s.key[i] = nms_key(cand_score[(size_t)(b * L + l) * cap1 + slot], (uint32_t)i);
      s.val[i] = ((uint32_t)l << 16) | (uint32_t)slot;
    } else { s.key[i] = ~0ull; s.val[i] = 0xFFFFFFFFu; }
  }
  __syncthreads();
  block_bitonic_sort(s.key, s.val, np2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int l = s.val[i] >> 16, slot = s.val[i] & 0xFFFF;
    const float4 bx = cell_box(cand_cell[(size_t)(b * L + l) * cap1 + slot], lt.scale[l]);
    s.sb[i] = bx;
    s.sa[i] = nms_area<0>(bx);
  }
  __syncthreads();
  const int nk = block_nms_sorted<0>(s.sb, s.sa, n, 0.7f, s.kept, &sc);
  for (int i = threadIdx.x; i < nk; i += blockDim.x) {
    const int p = s.kept[i];
    const int l = s.val[p] >> 16, slot = s.val[p] & 0xFFFF;
    const float4 q = s.sb[p];
    const float4 r = cand_reg[(size_t)(b * L + l) * cap1 + slot];
    const float regw = sub_rn(q.z, q.x), regh = sub_rn(q.w, q.y);
    float4 bx = make_float4(add_rn(q.x, mul_rn(r.x, regw)), add_rn(q.y, mul_rn(r.y, regh)), add_rn(q.z, mul_rn(r.z, regw)),
                            add_rn(q.w, mul_rn(r.w, regh)));
    bx = rerec(bx);
    s2_box[(size_t)b * cap2 + i] = bx;
    s2_pad[(size_t)b * cap2 + i] = pad_box(bx, W, H);
  }
  if (threadIdx.x == 0) s2_count[b] = nk;
}

// ---- stage 2 tail: score > t1 (detect_face.py:119-125), NMS(0.7) (:128), bbreg (:130), rerec (:131), pad (:136).
__global__ void __launch_bounds__(NTH) stage2_image_kernel(int W, int H, int cap2, const int* __restrict__ s2_count,
                                                           const float4* __restrict__ s2_box, const float* __restrict__ prob,
                                                           const float4* __restrict__ reg, float thr, int cap3,
                                                           int* __restrict__ s3_count, float4* __restrict__ s3_box,
                                                           int4* __restrict__ s3_pad, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ NmsScratch sc;
  const int b = blockIdx.x;
  const int n_in = min(s2_count[b], cap2);
  const int np2 = next_pow2(max(n_in, 2));
  SortSmem s = carve(smem, np2);
  const size_t base = (size_t)b * cap2;
  // failing candidates get the padding key: they sort behind every passing one; n = number of passing candidates
  __shared__ int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    const bool pass = i < n_in && prob[base + i] > thr;
    if (pass) { s.key[i] = nms_key(prob[base + i], (uint32_t)i); s.val[i] = (uint32_t)i; ++local; }
    else { s.key[i] = ~0ull; s.val[i] = 0xFFFFFFFFu; }
  }
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  const int n = s_n;
  block_bitonic_sort(s.key, s.val, np2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 bx = s2_box[base + s.val[i]];
    s.sb[i] = bx;
    s.sa[i] = nms_area<0>(bx);
  }
  __syncthreads();
  const int nk_all = block_nms_sorted<0>(s.sb, s.sa, n, 0.7f, s.kept, &sc);
  if (nk_all > cap3 && threadIdx.x == 0) atomicOr(status, 4);
  const int nk = min(nk_all, cap3);
  for (int i = threadIdx.x; i < nk; i += blockDim.x) {
    const int p = s.kept[i];
    float4 bx = bbreg(s.sb[p], reg[base + s.val[p]]);
    bx = rerec(bx);
    s3_box[(size_t)b * cap3 + i] = bx;
    s3_pad[(size_t)b * cap3 + i] = pad_box(bx, W, H);
  }
  if (threadIdx.x == 0) s3_count[b] = nk;
}

// ---- stage 3 tail: score > t2 (:151-157), landmarks to image coordinates (:159-163, from the PRE-bbreg box), bbreg
// (:164), "Min" NMS 0.7 (:168, :221-257), optional area-descending reorder of MTCNN.detect (mtcnn.py:334-340).
__global__ void __launch_bounds__(NTH) stage3_image_kernel(int cap3, const int* __restrict__ s3_count,
                                                           const float4* __restrict__ s3_box, const float* __restrict__ prob,
                                                           const float4* __restrict__ reg, const float* __restrict__ lmk, float thr,
                                                           int select_largest, int capf, int* __restrict__ out_count,
                                                           float* __restrict__ out_box, float* __restrict__ out_pts,
                                                           int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ NmsScratch sc;
  __shared__ int s_n;
  const int b = blockIdx.x;
  const int n_in = min(s3_count[b], cap3);
  const int np2 = next_pow2(max(n_in, 2));
  SortSmem s = carve(smem, np2);
  const size_t base = (size_t)b * cap3;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  int local = 0;
  for (int i = threadIdx.x; i < np2; i += blockDim.x) {
    const bool pass = i < n_in && prob[base + i] > thr;
    // "Min" NMS visits by ascending stable argsort from the END: ties -> later candidate first
    if (pass) { s.key[i] = nms_key(prob[base + i], 0xFFFFFFFFu - (uint32_t)i); s.val[i] = (uint32_t)i; ++local; }
    else { s.key[i] = ~0ull; s.val[i] = 0xFFFFFFFFu; }
  }
  if (local) atomicAdd(&s_n, local);
  __syncthreads();
  const int n = s_n;
  block_bitonic_sort(s.key, s.val, np2);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 bx = bbreg(s3_box[base + s.val[i]], reg[base + s.val[i]]);
    s.sb[i] = bx;
    s.sa[i] = nms_area<1>(bx);
  }
  __syncthreads();
  const int nk_all = block_nms_sorted<1>(s.sb, s.sa, n, 0.7f, s.kept, &sc);
  if (nk_all > capf && threadIdx.x == 0) atomicOr(status, 8);
  const int nk = min(nk_all, capf);
  // output order: NMS pick order (score descending), or area descending when select_largest.  np.argsort(area)[::-1]
  // is an ascending sort read backwards: equal areas -> later pick first.  Reuse the key/val arrays for that sort.
  __syncthreads();
  const int mp2 = next_pow2(max(nk, 2));
  unsigned long long* okey = s.key;         // safe: sorted keys are no longer needed
  uint32_t* oval = reinterpret_cast<uint32_t*>(s.sa);      // areas no longer needed
  for (int i = threadIdx.x; i < mp2; i += blockDim.x) {
    if (i < nk) {
      const float4 bx = s.sb[s.kept[i]];
      const float area = mul_rn(sub_rn(bx.z, bx.x), sub_rn(bx.w, bx.y));
      okey[i] = select_largest ? nms_key(area, 0xFFFFFFFFu - (uint32_t)i) : (unsigned long long)i;
      oval[i] = (uint32_t)i;
    } else { okey[i] = ~0ull; oval[i] = 0xFFFFFFFFu; }
  }
  __syncthreads();
  block_bitonic_sort(okey, oval, mp2);
  for (int o = threadIdx.x; o < nk; o += blockDim.x) {
    const int p = s.kept[oval[o]];
    const int src = (int)s.val[p];
    const float4 bx = s.sb[p];
    float* ob = out_box + ((size_t)b * capf + o) * 5;
    ob[0] = bx.x; ob[1] = bx.y; ob[2] = bx.z; ob[3] = bx.w; ob[4] = prob[base + src];
    const float4 pre = s3_box[base + src];
    const float w = add_rn(sub_rn(pre.z, pre.x), 1.0f), h = add_rn(sub_rn(pre.w, pre.y), 1.0f);
    float* op = out_pts + ((size_t)b * capf + o) * 10;
    const float* lm = lmk + (base + src) * 10;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      op[2 * j] = sub_rn(add_rn(mul_rn(w, lm[j]), pre.x), 1.0f);
      op[2 * j + 1] = sub_rn(add_rn(mul_rn(h, lm[5 + j]), pre.y), 1.0f);
    }
  }
  if (threadIdx.x == 0) out_count[b] = nk;
}

template <class K>
int set_smem(K kern, size_t bytes) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : 1;
}

}  // namespace

extern "C" int vnfr_stage1_boxes(const VnfrPyramid* pyr, int cap1, const int32_t* cand_count, const uint32_t* cand_cell,
                                 const float* cand_score, const float* cand_reg, int32_t* keep1_count, int32_t* keep1, int cap2,
                                 int32_t* s2_count, float* s2_box, int32_t* s2_pad, int32_t* status, void* stream) {
  VNFR_REQUIRE(pyr != nullptr && cap1 >= 1 && cap1 <= 4096 && cap2 >= 1 && cap2 <= 4096, "caps must be in [1, 4096]");
  if (pyr->B == 0) return VNFR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LevelTable lt;
  lt.n_levels = pyr->n_levels;
  for (int l = 0; l < pyr->n_levels; ++l) lt.scale[l] = pyr->scale[l];
  if (pyr->n_levels > 0) {
    VNFR_REQUIRE(set_smem(stage1_level_nms_kernel, sort_smem_bytes(cap1)) == 0, "cannot set shared memory size");
    stage1_level_nms_kernel<<<pyr->B * pyr->n_levels, NTH, sort_smem_bytes(cap1), st>>>(lt, cap1, cand_count, cand_cell, cand_score,
                                                                                       keep1_count, keep1, status);
    ++g_vnfr_launches;
  }
  VNFR_REQUIRE(set_smem(stage1_image_kernel, sort_smem_bytes(cap2)) == 0, "cannot set shared memory size");
  stage1_image_kernel<<<pyr->B, NTH, sort_smem_bytes(cap2), st>>>(lt, pyr->W, pyr->H, cap1, cand_cell, cand_score,
                                                                  reinterpret_cast<const float4*>(cand_reg), keep1_count, keep1, cap2,
                                                                  s2_count, reinterpret_cast<float4*>(s2_box),
                                                                  reinterpret_cast<int4*>(s2_pad), status);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_stage2_boxes(int B, int H, int W, int cap2, const int32_t* s2_count, const float* s2_box, const float* s2_prob,
                                 const float* s2_reg, float threshold, int cap3, int32_t* s3_count, float* s3_box, int32_t* s3_pad,
                                 int32_t* status, void* stream) {
  VNFR_REQUIRE(cap2 >= 1 && cap2 <= 4096 && cap3 >= 1 && cap3 <= 4096, "caps must be in [1, 4096]");
  if (B == 0) return VNFR_OK;
  VNFR_REQUIRE(set_smem(stage2_image_kernel, sort_smem_bytes(cap2)) == 0, "cannot set shared memory size");
  stage2_image_kernel<<<B, NTH, sort_smem_bytes(cap2), (cudaStream_t)stream>>>(
      W, H, cap2, s2_count, reinterpret_cast<const float4*>(s2_box), s2_prob, reinterpret_cast<const float4*>(s2_reg), threshold, cap3,
      s3_count, reinterpret_cast<float4*>(s3_box), reinterpret_cast<int4*>(s3_pad), status);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}

extern "C" int vnfr_stage3_faces(int B, int cap3, const int32_t* s3_count, const float* s3_box, const float* s3_prob,
                                 const float* s3_reg, const float* s3_lmk, float threshold, int select_largest, int capf,
                                 int32_t* out_count, float* out_box, float* out_pts, int32_t* status, void* stream) {
  VNFR_REQUIRE(cap3 >= 1 && cap3 <= 4096 && capf >= 1, "caps out of range");
  if (B == 0) return VNFR_OK;
  VNFR_REQUIRE(set_smem(stage3_image_kernel, sort_smem_bytes(cap3)) == 0, "cannot set shared memory size");
  stage3_image_kernel<<<B, NTH, sort_smem_bytes(cap3), (cudaStream_t)stream>>>(
      cap3, s3_count, reinterpret_cast<const float4*>(s3_box), s3_prob, reinterpret_cast<const float4*>(s3_reg), s3_lmk, threshold,
      select_largest, capf, out_count, out_box, out_pts, status);
  ++g_vnfr_launches;
  VNFR_CHECK_LAUNCH();
  return VNFR_OK;
}
