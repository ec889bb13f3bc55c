/*
 * vnfr_b200.h -- C-ABI of the B200-native (sm_100a) face detect -> align -> embed -> classify hot path.
 *
 * Drop-in boundary: the reference (votnhan/VN_celeb_face_recognition) is pure Python and has
 * no FFI layer; its boundary is the Python class API  models.MTCNN.detect / forward / inference,
 * models.InceptionResnetV1.forward, models.MLPModel.forward  plus the glue functions of demo_image.py /
 * find_embedding.py.  The host-side mirror of that API lives in vn_celeb_face_recognition_b200/models/*.py and calls
 * ONLY the entry points declared here (via ctypes).  Every entry point replaces one or more library call sites of
 * the reference, cited as file:line relative to the reference root.
 *
 * Conventions: plain pointers and sizes, no torch / C++ types; all pointers are DEVICE pointers unless the name ends
 * in _host; `stream` is a cudaStream_t passed as void*; every function returns 0 on success or a VNFR_ERR_* code
 * (message via vnfr_last_error()); no function allocates device memory -- workspaces are passed in; no function
 * synchronises the stream.
 */
#ifndef VNFR_B200_H
#define VNFR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNFR_OK 0
#define VNFR_ERR_ARG 1
#define VNFR_ERR_CUDA 2
#define VNFR_ERR_UNSUPPORTED 3

#define VNFR_MAX_LEVELS 24

/* ---- library ---------------------------------------------------------------------------------------------------- */
const char* vnfr_last_error(void);
int vnfr_version(void);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches`). */
long long vnfr_launch_count(void);
/* Adds n to the counter: kernels replayed by a CUDA graph captured around vnfr_run_ops. */
int vnfr_count_launches(long long n);

/* ---- detection: pyramid plan ------------------------------------------------------------------------------------ */
/* Scale pyramid of detect_face (models/mtcnn_utils/detect_face.py:48-60, :71) and the derived P-Net map geometry
 * (models/mtcnn.py:38-49: conv3 -> pool2(ceil) -> conv3 -> conv3).  Host-side, fp64 like the reference. */
typedef struct {
  int32_t B, H, W;                       /* frames in the batch, frame size                                      */
  int32_t n_levels;
  float scale[VNFR_MAX_LEVELS];          /* (float) of the Python-double scale                                   */
  double scale_d[VNFR_MAX_LEVELS];
  int32_t lh[VNFR_MAX_LEVELS], lw[VNFR_MAX_LEVELS];   /* level size  int(H*s+1), int(W*s+1)                      */
  int32_t oh[VNFR_MAX_LEVELS], ow[VNFR_MAX_LEVELS];   /* P-Net output map size                                   */
  int64_t level_off[VNFR_MAX_LEVELS + 1];/* float offset of level l in the pyramid buffer, layout [l][b][3][lh][lw] */
  int64_t map_off[VNFR_MAX_LEVELS + 1];  /* cell offset of level l in dense P-Net maps, layout [l][b][oh][ow]    */
  int32_t tiles_x[VNFR_MAX_LEVELS], tiles_y[VNFR_MAX_LEVELS];
  int32_t tile_off[VNFR_MAX_LEVELS + 1]; /* first P-Net tile of level l (per image)                              */
  int64_t px_off[VNFR_MAX_LEVELS + 1];   /* pixel offset (per image) of level l, for the resize kernel grid      */
} VnfrPyramid;

int vnfr_pyramid_plan(int B, int H, int W, int min_face_size, double factor, VnfrPyramid* out_host);

/* ---- detection stage kernels -------------------------------------------------------------------------------------
 * frames: u8 [B][H][W][3] (RGB, the layout detect_face receives, detect_face.py:26-41).                            */

/* imresample(area) + (x-127.5)*0.0078125 for EVERY pyramid level in one launch (detect_face.py:46, :71-72, :304-306).
 * levels: fp32, layout given by VnfrPyramid.level_off.                                                             */
int vnfr_pyramid_resize_norm(const VnfrPyramid* pyr_host, const uint8_t* frames, float* levels, void* stream);

/* P-Net weights: `packed_host` (6632 floats, layout produced by models/mtcnn.py:_pack_pnet) -> the kernel's layout, written
 * to the HOST buffer out_host (vnfr_pnet_packed_bytes() bytes).  The caller uploads it to a 16-byte aligned device buffer it
 * owns and passes that to vnfr_pnet_sweep_compact: the library keeps no weights of its own (no hidden state).            */
int vnfr_pnet_packed_bytes(void);
int vnfr_pnet_pack_weights(const float* packed_host, int n_floats, void* out_host, int out_bytes);

/* PNet.forward over every level + generateBoundingBox's threshold, fused (mtcnn.py:38-49; detect_face.py:73-75,
 * :203-218).  Candidates of image b / level l go to segment seg = b*n_levels + l with capacity `cap`:
 *   cand_count[seg]; cand_cell[seg][i] = (y<<16)|x; cand_score[seg][i]; cand_reg[seg][i][4].
 * dense_prob / dense_reg (nullable, layout VnfrPyramid.map_off: prob [cell], reg [4][cell] per (l,b)) are for parity
 * tests only.                                                                                                      */
int vnfr_pnet_sweep_compact(const VnfrPyramid* pyr_host, const float* levels, const void* pnet_weights, float threshold, int cap,
                            int32_t* cand_count, uint32_t* cand_cell, float* cand_score, float* cand_reg,
                            float* dense_prob, float* dense_reg, void* stream);

/* Generic segmented NMS (torchvision.ops.nms semantics for mode 0: detect_face.py:79, :93, :128; nms_numpy "Min"
 * semantics for mode 1: detect_face.py:221-257).  Segment s holds n = min(count[s], cap) boxes at boxes[s][i][4],
 * scores[s][i]; visit order = score descending, ties by ascending i (mode 0) or descending i (mode 1).
 * keep[s][0..keep_count[s]) = kept indices i in visit order.                                                       */
int vnfr_nms_segments(int n_segments, int cap, const int32_t* count, const float* boxes, const float* scores,
                      float threshold, int mode, int32_t* keep_count, int32_t* keep, void* stream);

/* Stage-1 tail on the device (detect_face.py:79-104): NMS(0.5) per (image, level), NMS(0.7) per image across levels,
 * regression without +1, rerec, pad.  keep1_count [B*L] / keep1 [B*L][cap1] are scratch.  Outputs per image b:
 * s2_count[b], s2_box[b][i][4] (squared-up float boxes), s2_pad[b][i] = (x, y, ex, ey) int32 (detect_face.py:277-289).
 * status: bit 0/1/2/3/4 set when cap1/cap2/cap3/capf/max_faces overflowed, bit 5 when an R-/O-Net crop workspace
 * overflowed (results are then truncated).                                                                         */
int vnfr_stage1_boxes(const VnfrPyramid* pyr_host, int cap1, const int32_t* cand_count, const uint32_t* cand_cell,
                      const float* cand_score, const float* cand_reg, int32_t* keep1_count, int32_t* keep1, int cap2,
                      int32_t* s2_count, float* s2_box, int32_t* s2_pad, int32_t* status, void* stream);

/* R-Net / O-Net over every candidate box of every frame: crop imgs[b,:,y-1:ey,x-1:ex], area-resize to 24 / 48,
 * normalise, network forward (detect_face.py:108-117, :136-146, :16-23; mtcnn.py:84-99, :138-157).
 * weights: packed fp32 device buffer (vnfr_rnet_weight_floats() / vnfr_onet_weight_floats() floats; layout in
 * csrc/detect_heads.cu, produced by models/mtcnn.py).  prob[b][i] = softmax[:,1], reg[b][i][4], lmk[b][i][10].
 * offs [B+1] is scratch (exclusive scan of counts).  crops [crop_cap][3][S][S] fp32 (S = 24 / 48, 16-byte aligned) is the
 * workspace the crop kernel fills, image-major in candidate order, and the network kernel reads; candidates beyond
 * crop_cap are dropped and bit 5 of *status is set.                                                                 */
int vnfr_rnet_weight_floats(void);
int vnfr_onet_weight_floats(void);
int vnfr_rnet_forward(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                      const float* weights, float* prob, float* reg, int32_t* offs, float* crops, int crop_cap,
                      int32_t* status, void* stream);
int vnfr_onet_forward(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                      const float* weights, float* prob, float* reg, float* lmk, int32_t* offs, float* crops, int crop_cap,
                      int32_t* status, void* stream);

/* The layers after the last tensor-core convolution of R-Net (maxpool 3/2 -> conv3 2x2 + PReLU -> dense4 + PReLU -> heads,
 * mtcnn.py:84-99) and O-Net (maxpool 2/2 -> conv4 2x2 + PReLU -> dense5 + PReLU -> heads, mtcnn.py:138-157) as three
 * split-precision tensor-core GEMMs over all crops of the batch (heads_chain.cu) instead of the per-crop FMA kernels.
 * Every fp32 weight matrix is passed as two fp16 planes: rows [0, N_pad) = fp16(w), rows [N_pad, 2 N_pad) = fp16(w - hi),
 * N_pad = N rounded up to 128, row length K:
 *   w[0]: the 2x2 convolution, K = 256: column (ky*2 + kx)*64 + c = weight[n][c][ky][kx], channels c >= C_in zero;
 *   w[1]: the dense layer exactly as torch stores it (K = 576 / 1152, the reference's (W,H,C) flatten order);
 *   w[2]: the heads stacked: R-Net [dense5_1 (2); dense5_2 (4)], O-Net [dense6_1 (2); dense6_2 (4); dense6_3 (10)].
 * bias[l]: fp32 [N_pad]; alpha[0..1]: PReLU slopes fp32 [N_pad] of the convolution and the dense layer.
 * planes: workspace of vnfr_heads_back_workspace_bytes(onet, crop_cap) bytes, 1024-byte aligned.                        */
typedef struct VnfrHeadsBack {
  const void* w[3];
  const float* bias[3];
  const float* alpha[2];
  void* planes;
} VnfrHeadsBack;
long long vnfr_heads_back_workspace_bytes(int onet, int crop_cap);

/* R-Net with conv2 (28 -> 48, 3x3; 64 % of its FLOPs) on the tensor cores in split precision (two fp16 parts per fp32 operand,
 * three products, fp32 accumulation; vnfr_conv_run with VnfrConvOp.split3 = 2) over all crops in one launch; the other layers
 * stay on the fp32 FMA path.  w2_split: fp16 [48][896] (split2 layout, sv_ck 32); p1: fp16 [crop_cap][11][11][64] and
 * c2: fp32 [crop_cap][81][48] workspaces.  back (nullable): run the layers after conv2 on the tensor cores as well
 * (VnfrHeadsBack above).  Same outputs and semantics as vnfr_rnet_forward.                                              */
int vnfr_rnet_forward_tc(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                         const float* weights, const void* w2_split, float* prob, float* reg, int32_t* offs, float* crops,
                         void* p1, float* c2, int crop_cap, int32_t* status, const VnfrHeadsBack* back, void* stream);

/* O-Net with conv2 (63 % of its FLOPs) on the tensor cores in split precision, fp32 accumulation (vnfr_conv_run with
 * VnfrConvOp.split3 = split_mode) over all crops in one launch; the other layers stay on the fp32 FMA path.
 *   split_mode 1: three bf16 parts per fp32 operand, six products.  w2_split: bf16 [64][1728], p1: bf16 [crop_cap][23][23][96]
 *   split_mode 2: two fp16 parts, three products (half the tensor work). w2_split: fp16 [64][896], p1: fp16 [crop_cap][23][23][64]
 * c2: fp32 [crop_cap][441][64] workspace.
 * w3_split (nullable): conv3 (64 -> 64, 3x3; 18 % of the FLOPs) on the tensor cores as well, always as two fp16 parts:
 * fp16 [64][1728] (pack_conv_split2, sv_ck 64) with workspaces p3: fp16 [crop_cap][10][10][128] and c3: fp32 [crop_cap][64][64];
 * NULL keeps conv3 on the FMA path (p3 / c3 are then ignored).  back (nullable, needs w3_split): the layers after conv3 on
 * the tensor cores as well (VnfrHeadsBack above).  Same outputs and semantics as vnfr_onet_forward.                     */
int vnfr_onet_forward_tc(const uint8_t* frames, int B, int H, int W, int cap, const int32_t* count, const int32_t* pad,
                         const float* weights, const void* w2_split, int split_mode, float* prob, float* reg, float* lmk,
                         int32_t* offs, float* crops, void* p1, float* c2, const void* w3_split, void* p3, float* c3,
                         int crop_cap, int32_t* status, const VnfrHeadsBack* back, void* stream);

/* Stage-2 tail (detect_face.py:119-136): score > threshold, NMS(0.7), bbreg, rerec, pad. */
int vnfr_stage2_boxes(int B, int H, int W, int cap2, const int32_t* s2_count, const float* s2_box, const float* s2_prob,
                      const float* s2_reg, float threshold, int cap3, int32_t* s3_count, float* s3_box, int32_t* s3_pad,
                      int32_t* status, void* stream);

/* Stage-3 tail (detect_face.py:148-169 incl. the host NumPy "Min" NMS :221-274; mtcnn.py:334-340 when select_largest):
 * out_box[b][i] = (x1,y1,x2,y2,score), out_pts[b][i] = (x0,y0,...,x4,y4), in the order MTCNN.detect returns them.  */
int vnfr_stage3_faces(int B, int cap3, const int32_t* s3_count, const float* s3_box, const float* s3_prob,
                      const float* s3_reg, const float* s3_lmk, float threshold, int select_largest, int capf,
                      int32_t* out_count, float* out_box, float* out_pts, int32_t* status, void* stream);

/* Faces -> encoder inputs, one gather kernel.  mode 0: MTCNN.extract semantics for tensor frames (mtcnn.py:458-518,
 * detect_face.py:317-322, :342-378); mode 1: demo_video alignment (demo_image.py:174-199, :236-239;
 * align_face.py:51-57: 5-point similarity + cv2.warpAffine) with template_host[10] = center points (x,y)*5.
 * Faces are numbered image-major in detection order; offs [B+1] scratch.  face_u8 [max_faces][S][S][3] (nullable),
 * face_half (standardised, dtype 0 = bf16 / 1 = fp16): half_layout 0 = [max_faces][S][S][8] (3 real + 5 zero channels),
 * 1 = space-to-depth [max_faces][ceil(S/2)][ceil(S/2)][16], channel ((y&1)*2 + (x&1))*4 + c (the buffer must be zeroed
 * once by the caller when S is odd: pad sub-pixels are never written); face_img [max_faces] (nullable).            */
int vnfr_face_crops(const uint8_t* frames, int B, int H, int W, int capf, const int32_t* count, const float* box,
                    const float* pts, int mode, int image_size, int margin, const float* template_host, int dtype,
                    int max_faces, int32_t* offs, uint8_t* face_u8, void* face_half, int32_t* face_img, int32_t* status,
                    int half_layout, void* stream);

/* ---- encoder: implicit-GEMM convolution on tcgen05 / TMEM ---------------------------------------------------------
 * One fused op = conv (no bias) + folded-BN bias + optional residual add + optional ReLU, NHWC bf16 (or fp16, see `dtype`) in/out, writing
 * into a channel slice of a (possibly wider) destination: replaces BasicConv2d (inception_resnet_v1.py:12-33), the
 * block projections + `out*scale + x` + ReLU (:56-67, :85-95, :114-126), torch.cat (:63, :91, :120, :147, :179),
 * last_linear + last_bn (:296-297) and both MLP layers (mlp_model.py:10-15).                                        */
typedef struct {
  unsigned char tmap_w[128];      /* CUtensorMap of the packed weights, filled by vnfr_conv_prepare                */
  unsigned char tmap_a[128];      /* CUtensorMap of the activations (a_mode 1 / 2), filled by vnfr_conv_prepare   */
  const void* in;                 /* bf16 NHWC; pixel pitch in_pitch elements; already offset to first channel     */
  const void* weights;            /* bf16 [cout_pad][k_pad], k = (ky*KW + kx)*cin + c, zero padded                 */
  const float* bias;              /* fp32 [cout_pad]                                                               */
  const void* residual;           /* bf16, nullable: added before ReLU; pixel pitch res_pitch                      */
  void* out0;                     /* bf16 destination for output channels [0, n_split)                             */
  void* out1;                     /* bf16 destination for output channels [n_split, cout), nullable                */
  float* out_f32;                 /* nullable: fp32 destination [M][out_f32_pitch] (used instead of out0/out1)     */
  int32_t n_img, in_h, in_w, cin, in_pitch;
  int32_t kh, kw, stride, pad_h, pad_w;
  int32_t out_h, out_w;
  int32_t cout, cout_pad, k_pad, block_n;
  int32_t n_split, out0_pitch, out1_pitch, res_pitch, out_f32_pitch;
  int32_t relu;
  int32_t dtype;                  /* storage type of in/weights/residual/out0/out1: 0 = bf16, 1 = fp16             */
  int32_t a_mode;                 /* set by vnfr_conv_prepare: 0 = cp.async gather, 1 = TMA tiled (1x1), 3 = shifted view */
  int32_t epi_mode;               /* set by vnfr_conv_prepare: 0 = per-row register epilogue, 1 = shared-memory staged tile with
                                     TMA residual load + TMA store (single 16-bit destination)                          */
  unsigned char tmap_c[128];      /* CUtensorMap of the destination (epi_mode 1)                                       */
  unsigned char tmap_r[128];      /* CUtensorMap of the residual (epi_mode 1, residual != NULL)                        */
  const float* prelu_alpha;       /* nullable (shifted-view kernel only): per-channel PReLU slope applied instead of ReLU      */
  const int32_t* n_img_dev;       /* nullable (shifted-view kernel only): DEVICE count of valid images (<= n_img)              */
  int32_t split3;                 /* shifted-view kernel only: 1 = split-precision convolution.  `in` holds 3 planes of `cin/3`
                                     channels (hi | mid | lo bf16 parts of an fp32 activation), the weights are packed per tap
                                     as the 6 products (a0b0, a0b1, a1b0, a0b2, a1b1, a2b0): k = (tap*6 + j)*sv_ck + c; the
                                     fp32 accumulator then carries ~fp32 accuracy (dropped terms are O(2^-24)).
                                     2 = the same with 2 fp16 parts (hi | lo, dtype must be 1): `in` holds 2 planes of `cin/2`
                                     channels, 3 products (a0b0, a0b1, a1b0): k = (tap*3 + j)*sv_ck + c; dropped term O(2^-22),
                                     half the tensor-core work                                                               */
  int32_t reserved[1];            /* [0] = sv_ck: 0, or 32 / 64 = request the shifted-view kernel (stride-1 k x k convs, cout <=
                                     256) with weights packed k = (tap*ceil(cin/sv_ck) + chunk)*sv_ck + c               */
} VnfrConvOp;

/* Fills op->tmap_w (cuTensorMapEncodeTiled on op->weights, box {64, block_n}, 128B swizzle) and validates the op. */
int vnfr_conv_prepare(VnfrConvOp* op_host);
int vnfr_conv_run(const VnfrConvOp* op_host, void* stream);

/* ---- encoder: a whole Block17 in one persistent kernel ---------------------------------------------------------------
 * models/inception_resnet_v1.py:70-95: branch0 / branch1.0 (1x1, one GEMM with N = 256) -> 1x7 -> 7x1 -> 1x1 projection
 * (+ bias, residual scale folded in) + x -> ReLU, IN PLACE on x: NHWC 16-bit [n_img][8][8][896] (the 8x8 map of a 160x160
 * crop).  One CTA owns two images from x to the updated x; the branch activations stay in shared memory / TMEM.
 * Weights in the layout of VnfrConvOp.weights: w1 [256][896] (branch0 rows 0..127, branch1.0 rows 128..255), w2 [128][896]
 * (k = kx*128 + c), w3 [128][896] (k = ky*128 + c), w4 [896][256]; folded-BN biases b1 [256], b2 [128], b3 [128], b4 [896]. */
typedef struct {
  unsigned char tmap[5][128];     /* x (4-D, v-order box), w1, w2, w3, w4: filled by vnfr_block17_prepare              */
  void* x;
  const void *w1, *w2, *w3, *w4;
  const float *b1, *b2, *b3, *b4;
  int32_t n_img;
  int32_t dtype;                  /* 0 = bf16, 1 = fp16                                                              */
} VnfrBlock17Op;
int vnfr_block17_prepare(VnfrBlock17Op* op_host);
int vnfr_block17_run(const VnfrBlock17Op* op_host, void* stream);

/* A flat op list = one encoder / classifier forward.  kind 0: convolution (all fields of `conv`); kind 1:
 * MaxPool2d(3,2) and kind 2: AdaptiveAvgPool2d(1) reuse conv.{in, out0, n_img, in_h, in_w, cin, in_pitch, out0_pitch};
 * kind 3: fused Block17, `ext` points at a host VnfrBlock17Op (conv is unused).
 * vnfr_run_ops launches them in order on `stream` (one host call per forward instead of one per layer). */
typedef struct {
  int32_t kind;
  int32_t reserved;
  VnfrConvOp conv;
  const void* ext;
} VnfrOp;
int vnfr_run_ops(const VnfrOp* ops_host, int n_ops, void* stream);

/* MaxPool2d(3, stride 2) on NHWC bf16 (inception_resnet_v1.py:147 `branch2`, :179 `branch3`, :224 `maxpool_3a`). */
int vnfr_maxpool3s2_nhwc(const void* in, int n_img, int in_h, int in_w, int c, int in_pitch, void* out, int out_pitch,
                         int dtype, void* stream);
/* AdaptiveAvgPool2d(1) on NHWC bf16 -> bf16 [n_img][c] (inception_resnet_v1.py:294). */
int vnfr_avgpool_nhwc(const void* in, int n_img, int hw, int c, int in_pitch, void* out, int dtype, void* stream);
/* fp32 NCHW (3 channels) -> bf16 NHWC with 8 channels (3 real + 5 zero): input adapter of InceptionResnetV1.forward. */
int vnfr_nchw3_to_nhwc8(const float* in, int n_img, int h, int w, void* out, int dtype, void* stream);
/* Same input -> the space-to-depth layout [n][ceil(h/2)][ceil(w/2)][16] (see vnfr_face_crops half_layout 1). */
int vnfr_nchw3_to_s2d16(const float* in, int n_img, int h, int w, void* out, int dtype, void* stream);
/* transforms_default on the device (data_loader/__init__.py:27-34, 52-56: np.float32 -> (x-127.5)/128 -> CHW) fused with the
 * layout change: u8 HWC faces [n][h][w][3] -> the space-to-depth encoder input (see vnfr_face_crops half_layout 1).        */
int vnfr_u8hwc_to_s2d16(const uint8_t* in, int n_img, int h, int w, void* out, int dtype, void* stream);
/* F.normalize(p=2, dim=1) (inception_resnet_v1.py:302): x fp32 [n][d] -> emb fp32 [n][d] and bf16 copy (nullable). */
int vnfr_l2norm_rows(const float* x, int n, int d, int x_pitch, float* emb, void* emb_half, int dtype, void* stream);
/* F.log_softmax(dim=1) + argmax + exp(max log-prob) (mlp_model.py:14; demo_image.py:126-130).
 * logits fp32 [n][pitch] (first c valid) -> logp fp32 [n][c] (nullable), label int64 [n], prob fp32 [n].           */
int vnfr_logsoftmax_argmax(const float* logits, int n, int c, int pitch, float* logp, int64_t* label, float* prob,
                           void* stream);
/* Row-wise top-k (k <= 8; largest first, ties towards the lower index) of scores fp32 [n][pitch] (first g columns valid):
 * out_val [n][k], out_idx [n][k] = col_offset + column.  accumulate != 0 merges with the values already in out_val /
 * out_idx (tile-by-tile search of a gallery larger than one score buffer).  The scores are cosine similarities produced
 * by vnfr_conv_run (queries as a [n][1][1][512] activation, a gallery tile as the [g][512] "weights"): the north star's
 * cosine top-k against a gallery shard (BASELINE.json config 5; the reference itself has no gallery search). */
int vnfr_topk_rows(const float* scores, int n, int g, int pitch, int k, int col_offset, int accumulate, float* out_val,
                   int32_t* out_idx, void* stream);

/* Cosine top-8 against a gallery shard with the top-k fused into the score GEMM: q 16-bit [m][512], gallery 16-bit
 * [g_pad][512] (g_pad a multiple of 256, rows >= g_valid ignored) -> out_val / out_idx [splits][m][8], best first, ties
 * towards the lower row, idx = index_offset + gallery row (-1 = no entry).  The m x g score matrix is never written: scores
 * go from TMEM straight into per-row running top-8 lists.  `splits` (1..g_pad/256) divides the gallery among CTAs when there
 * are fewer than ~148 query tiles; the caller merges the `splits` lists per query.                                       */
int vnfr_gallery_topk(const void* q, int m, const void* gallery, int g_valid, int g_pad, int dtype, int splits, int index_offset,
                      float* out_val, int32_t* out_idx, void* stream);

/* cv2.cvtColor(frame, COLOR_BGR2RGB) of demo_video.py:107-110 on the device: swaps bytes 0 and 2 of every packed u8 pixel
 * (n_pixels a multiple of 4; in == out allowed), so BGR video frames can be uploaded as they are decoded.               */
int vnfr_swap_rb_u8(const uint8_t* in, uint8_t* out, long long n_pixels, void* stream);

/* NV12 video frames [n][h*3/2][w] u8 (luma plane, then interleaved U,V at half resolution) -> packed RGB [n][h][w][3], BT.601
 * limited range, bit-identical to cv2.cvtColor(COLOR_YUV2RGB_NV12): the decode-side colour conversion of the ingest path
 * (demo_video.py:78-110) on the device, at half the host -> device bytes of RGB frames.                                     */
int vnfr_nv12_to_rgb_u8(const uint8_t* in, uint8_t* out, int n_img, int h, int w, void* stream);

/* ---- fused tail: pool -> bottleneck -> L2-normalise -> MLP -> log-softmax/argmax in ONE cooperative kernel --------------
 * Replaces avgpool_1a + last_linear + last_bn + F.normalize (models/inception_resnet_v1.py:294-302; `logits` +
 * log_softmax :298-300 when classify), MLPModel.forward (models/mlp_model.py:10-15) and the argmax / exp / threshold of
 * identify_person (demo_image.py:113-137).  Up to three linear layers, each followed by a row operation; every
 * contraction runs on the tensor cores in split precision (two fp16 parts per fp32 operand, three products, fp32
 * accumulation: fp32-level accuracy, so the predicted label equals the fp32 reference's).
 * Weights of a layer: fp16 [2*N_pad][K] = rows [0,N_pad) the hi parts fp16(W), rows [N_pad,2*N_pad) the lo parts
 * fp16(W - hi) (rows >= N zero); bias fp32 [N_pad].  a_in: fp16 [2*n_pad][K] scratch (hi rows, then lo rows), written by
 * the kernel; layer l+1's a_in must be layer l's a_next.  partial: fp32 [split_k][n_pad][N_pad] scratch.               */
typedef struct {
  int32_t K, N, N_pad;            /* K multiple of 64; N_pad multiple of 128, N <= 8192                              */
  int32_t split_k;                /* K ranges per output tile (partials are summed in a fixed order)                 */
  int32_t rowop;                  /* after the layer: 0 identity, 1 ReLU, 2 L2-normalise (F.normalize), 3 log_softmax + argmax */
  int32_t out_vec_pitch;
  const void* weights;
  const float* bias;
  float* partial;
  void* a_in;
  void* a_next;                   /* nullable (last layer)                                                           */
  float* out_vec;                 /* nullable: fp32 [n][out_vec_pitch] row-operation output (embedding / log-probs)   */
} VnfrTailLayer;

typedef struct {
  unsigned char tmap_a[3][128];   /* filled by vnfr_tail_prepare                                                     */
  unsigned char tmap_w[3][128];
  int32_t n_layers, n_pad;        /* n_pad: row capacity of the scratch buffers, multiple of 128                     */
  int32_t in_mode;                /* 0: x = 16-bit NHWC [n][hw][x_pitch] (first layer[0].K channels), mean over hw;
                                     1: x_f32 = fp32 [n][x_f32_pitch], first x_f32_cols columns                      */
  int32_t hw, x_pitch, x_dtype;   /* x_dtype: 0 = bf16, 1 = fp16                                                     */
  int32_t x_f32_pitch;
  int32_t x_f32_cols;             /* valid columns of x_f32 (<= layer[0].K; the rest of K is zero padding)           */
  int32_t emb_half_dtype;
  int32_t reserved;
  const void* x;
  const float* x_f32;
  VnfrTailLayer layer[3];
  void* emb_half;                 /* nullable: 16-bit [n][N] copy of the L2-normalised row (rowop 2)                 */
  int64_t* label;                 /* nullable outputs of rowop 3: argmax (or n_classes when prob < threshold)        */
  float* prob;                    /*   exp(max log-prob)                                                             */
  float* label_f;                 /* nullable: the same two values as floats with row pitch lp_pitch -- the label /  */
  float* prob_f;                  /*   probability columns of the all-gather send buffer (dist.py payload)           */
  int32_t lp_pitch;
  int32_t n_classes;
  const float* thr_class;         /* nullable: per-class thresholds [n_classes] (demo_image.py:118-124)              */
  float thr;                      /* scalar threshold otherwise (0 = keep every label)                               */
  int32_t count_value;            /* value written to *count_cell (the caller's total row count when it runs chunks) */
  float* count_cell;              /* nullable: receives (float)count_value -- the count cell of the send buffer     */
  unsigned int* grid_barrier;     /* one zero-initialisable device word                                              */
} VnfrTailOp;

int vnfr_tail_prepare(VnfrTailOp* op_host);
/* n <= n_pad rows; launches ONE cooperative kernel (plus a 4-byte memset of the barrier word) on `stream`.         */
int vnfr_tail_run(const VnfrTailOp* op_host, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VNFR_B200_H */
