"""CPU restatement of the MTCNN detection algorithm (the reference's ``detect_face`` and helpers).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Arithmetic is fp32 on the CPU in the reference's order of
operations; every function cites the reference lines it follows.  Third-party arithmetic that
is not under /root/reference is restated here from its published behaviour and pinned against the installed library in
tests/test_oracle_*.py:
  * torchvision.ops.nms / batched_nms (unpinned by the reference; 0.26.0 in this image) -> ``nms_iou``, ``batched_nms``
  * torch.nn.functional.interpolate(mode="area") == adaptive average pooling -> ``area_resize``
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import nets

f32 = np.float32


# --------------------------------------------------------------------------------------------------------------------
# pyramid / resize
# --------------------------------------------------------------------------------------------------------------------
def scale_pyramid(h, w, minsize, factor):
    """detect_face.py:48-60 (Python doubles)."""
    m = 12.0 / minsize
    minl = min(h, w) * m
    scale_i = m
    scales = []
    while minl >= 12:
        scales.append(scale_i)
        scale_i = scale_i * factor
        minl = minl * factor
    return scales


def level_size(h, w, scale):
    """detect_face.py:71."""
    return int(h * scale + 1), int(w * scale + 1)


def area_resize(x, size):
    """``imresample`` (detect_face.py:304-306) = F.interpolate(mode='area') = adaptive average pooling:
    out[i, j] = mean of in[floor(i*H/oh) : ceil((i+1)*H/oh), floor(j*W/ow) : ceil((j+1)*W/ow)] in fp32.
    x: torch (N,C,H,W) float32."""
    return F.interpolate(x, size=size, mode="area")


def area_resize_np(img_chw, oh, ow):
    """Plain-numpy restatement of adaptive average pooling for integer-valued inputs (u8 pixels): the window sum is
    exact in fp32 (< 2^24), then two fp32 divisions (sum / kh / kw, bit-identical to torch's CPU adaptive_avg_pool2d) --
    this is the arithmetic the CUDA kernels use.
    img_chw: (C,H,W) any dtype holding integers 0..255."""
    C, H, W = img_chw.shape
    integ = np.zeros((C, H + 1, W + 1), dtype=np.int64)
    integ[:, 1:, 1:] = np.cumsum(np.cumsum(img_chw.astype(np.int64), axis=1), axis=2)
    i = np.arange(oh)
    j = np.arange(ow)
    y0 = (i * H) // oh
    y1 = -((-(i + 1) * H) // oh)
    x0 = (j * W) // ow
    x1 = -((-(j + 1) * W) // ow)
    s = (integ[:, y1][:, :, x1] - integ[:, y0][:, :, x1] - integ[:, y1][:, :, x0] + integ[:, y0][:, :, x0])
    kh = (y1 - y0).astype(f32)[None, :, None]
    kw = (x1 - x0).astype(f32)[None, None, :]
    return s.astype(f32) / kh / kw          # torch's CPU kernel divides by kh, then by kw (pinned in tests)


def normalize(x):
    """detect_face.py:72 / :114 / :143."""
    return (x - 127.5) * 0.0078125


# --------------------------------------------------------------------------------------------------------------------
# box generation and box math
# --------------------------------------------------------------------------------------------------------------------
def generate_bounding_box(reg, prob1, scale, thresh):
    """detect_face.py:203-218.  reg (B,4,h,w), prob1 (B,h,w) torch fp32.  Returns boxes (N,9) fp32, image_inds (N,)
    in (b, y, x) raster order."""
    stride, cellsize = 2, 12
    reg = reg.permute(1, 0, 2, 3)
    mask = prob1 >= thresh
    mask_inds = mask.nonzero()
    image_inds = mask_inds[:, 0]
    score = prob1[mask]
    reg = reg[:, mask].permute(1, 0)
    bb = mask_inds[:, 1:].type(reg.dtype).flip(1)
    q1 = ((stride * bb + 1) / scale).floor()
    q2 = ((stride * bb + cellsize - 1 + 1) / scale).floor()
    return torch.cat([q1, q2, score.unsqueeze(1), reg], dim=1), image_inds


def stage1_regress(boxes):
    """detect_face.py:96-102 (no +1).  boxes (N,9) -> (N,5)."""
    regw = boxes[:, 2] - boxes[:, 0]
    regh = boxes[:, 3] - boxes[:, 1]
    qq1 = boxes[:, 0] + boxes[:, 5] * regw
    qq2 = boxes[:, 1] + boxes[:, 6] * regh
    qq3 = boxes[:, 2] + boxes[:, 7] * regw
    qq4 = boxes[:, 3] + boxes[:, 8] * regh
    return torch.stack([qq1, qq2, qq3, qq4, boxes[:, 4]]).permute(1, 0).contiguous()


def bbreg(boxes, reg):
    """detect_face.py:188-200 (+1 widths).  Returns a new tensor."""
    boxes = boxes.clone()
    w = boxes[:, 2] - boxes[:, 0] + 1
    h = boxes[:, 3] - boxes[:, 1] + 1
    b1 = boxes[:, 0] + reg[:, 0] * w
    b2 = boxes[:, 1] + reg[:, 1] * h
    b3 = boxes[:, 2] + reg[:, 2] * w
    b4 = boxes[:, 3] + reg[:, 3] * h
    boxes[:, :4] = torch.stack([b1, b2, b3, b4]).permute(1, 0)
    return boxes


def rerec(boxes):
    """detect_face.py:292-301: square-up about the centre.  Returns a new tensor."""
    boxes = boxes.clone()
    h = boxes[:, 3] - boxes[:, 1]
    w = boxes[:, 2] - boxes[:, 0]
    l = torch.max(w, h)
    boxes[:, 0] = boxes[:, 0] + w * 0.5 - l * 0.5
    boxes[:, 1] = boxes[:, 1] + h * 0.5 - l * 0.5
    boxes[:, 2:4] = boxes[:, :2] + l.repeat(2, 1).permute(1, 0)
    return boxes


def pad(boxes, w, h):
    """detect_face.py:277-289: trunc -> int32, clamp x,y >= 1, ex <= w, ey <= h.  Returns y, ey, x, ex (numpy int32)."""
    b = boxes.trunc().int().cpu().numpy()
    x = b[:, 0].copy(); y = b[:, 1].copy(); ex = b[:, 2].copy(); ey = b[:, 3].copy()
    x[x < 1] = 1
    y[y < 1] = 1
    ex[ex > w] = w
    ey[ey > h] = h
    return y, ey, x, ex


# --------------------------------------------------------------------------------------------------------------------
# NMS
# --------------------------------------------------------------------------------------------------------------------
def nms_iou(boxes, scores, thr):
    """Restatement of torchvision.ops.nms (CPU kernel, torchvision/csrc/ops/cpu/nms_kernel.cpp -- third-party,
    unpinned; behaviour pinned against the installed 0.26.0 in tests/test_oracle_nms.py): stable score-descending
    visit order, area (x2-x1)*(y2-y1), inter clamped at 0 without +1, suppress iff inter/(a_i+a_j-inter) > thr, all fp32.
    boxes (N,4) fp32 numpy, scores (N,) fp32.  Returns kept indices in visit order (int64)."""
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores, kind="stable")
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr32 = f32(thr)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(f32(0), xx2 - xx1); h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr32]] = True
    return np.asarray(keep, dtype=np.int64)


def batched_nms(boxes, scores, idxs, thr, faithful=True):
    """torchvision.ops.boxes.batched_nms as called at detect_face.py:79, :93, :128.

    faithful=True follows torchvision 0.26: the fp32 coordinate-offset trick when boxes.numel() <= 4000 on the CPU,
    else per-class NMS followed by a (non-stable) score-descending sort.  faithful=False is the mathematically intended
    per-image NMS, returned in stable score-descending order -- what the CUDA kernels implement."""
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    idxs = np.asarray(idxs)
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    if faithful and boxes.size <= 4000:
        max_coordinate = boxes.max()
        offsets = idxs.astype(f32) * (max_coordinate + f32(1))
        return nms_iou(boxes + offsets[:, None], scores, thr)
    keep_mask = np.zeros(scores.shape[0], dtype=bool)
    for c in np.unique(idxs):
        cur = np.where(idxs == c)[0]
        keep_mask[cur[nms_iou(boxes[cur], scores[cur], thr)]] = True
    keep = np.where(keep_mask)[0]
    return keep[np.argsort(-scores[keep], kind="stable")]


def nms_min(boxes, scores, thr, method="Min", tie="numpy"):
    """``nms_numpy`` (detect_face.py:221-257): +1 areas, overlap = inter/min(area) ("Min") or IoU ("Union"), keep iff
    o <= thr, visit by ascending argsort from the end.  The reference's ``np.argsort`` is not stable; tie="numpy" calls
    np.argsort exactly like the reference, tie="stable" is the documented convention of the CUDA kernel (ascending
    stable sort taken from the end: among equal scores the LATER candidate is visited first -- identical to numpy for
    N <= 16 where numpy's introsort is an insertion sort)."""
    boxes = np.asarray(boxes, dtype=f32)
    s = np.asarray(scores, dtype=f32)
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0].copy(), boxes[:, 1].copy(), boxes[:, 2].copy(), boxes[:, 3].copy()
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    I = np.argsort(s) if tie == "numpy" else np.argsort(s, kind="stable")
    pick = []
    while I.size > 0:
        i = I[-1]
        pick.append(i)
        idx = I[0:-1]
        xx1 = np.maximum(x1[i], x1[idx]); yy1 = np.maximum(y1[i], y1[idx])
        xx2 = np.minimum(x2[i], x2[idx]); yy2 = np.minimum(y2[i], y2[idx])
        w = np.maximum(f32(0.0), xx2 - xx1 + 1); h = np.maximum(f32(0.0), yy2 - yy1 + 1)
        inter = w * h
        if method == "Min":
            o = inter / np.minimum(area[i], area[idx])
        else:
            o = inter / (area[i] + area[idx] - inter)
        I = I[np.where(o <= f32(thr))]
    return np.asarray(pick, dtype=np.int64)


def batched_nms_min(boxes, scores, idxs, thr, method="Min", faithful=True):
    """``batched_nms_numpy`` (detect_face.py:260-274).  faithful=True: fp32 coordinate-offset trick over the whole
    batch; False: per image, concatenated image-major (what the CUDA kernel does; the per-image ORDER is the same)."""
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    idxs = np.asarray(idxs)
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    if faithful:
        max_coordinate = boxes.max()
        offsets = idxs.astype(f32) * (max_coordinate + f32(1))
        return nms_min(boxes + offsets[:, None], scores, thr, method, tie="numpy")
    out = []
    for c in np.unique(idxs):
        cur = np.where(idxs == c)[0]
        out.append(cur[nms_min(boxes[cur], scores[cur], thr, method, tie="stable")])
    return np.concatenate(out) if out else np.zeros((0,), dtype=np.int64)


# --------------------------------------------------------------------------------------------------------------------
# the three stages
# --------------------------------------------------------------------------------------------------------------------
def crop_resize_batch(imgs, image_inds, y, ey, x, ex, size):
    """detect_face.py:108-114 / :136-143: per box slice imgs[b, :, y-1:ey, x-1:ex] -> area resize to size^2 ->
    normalise.  imgs torch (B,3,H,W) fp32."""
    out = []
    for k in range(len(y)):
        if ey[k] > (y[k] - 1) and ex[k] > (x[k] - 1):
            img_k = imgs[int(image_inds[k]), :, (y[k] - 1):ey[k], (x[k] - 1):ex[k]].unsqueeze(0)
            out.append(area_resize(img_k, (size, size)))
    if not out:
        return torch.zeros(0, 3, size, size)
    return normalize(torch.cat(out, dim=0))


def _batched(fn, sd, x, bs=512):
    """fixed_batch_process, detect_face.py:16-23."""
    outs = [fn(sd, x[i:i + bs]) for i in range(0, len(x), bs)]
    return tuple(torch.cat(v, dim=0) for v in zip(*outs))


def detect_face(imgs_u8, minsize, pnet_sd, rnet_sd, onet_sd, threshold, factor, faithful=True, taps=None):
    """detect_face.py:25-185.  imgs_u8: numpy/torch uint8 (B,H,W,3).  Returns (list of (n_i,5) boxes, list of (n_i,5,2)
    points) as fp32 numpy, per image, in final-NMS pick order.  ``taps`` (dict) collects every intermediate that a
    kernel-level parity test needs.  ``faithful`` selects the torchvision/numpy batched-NMS behaviour (see
    ``batched_nms``)."""
    def tap(k, v):
        if taps is not None:
            taps[k] = v
        return v

    if not isinstance(imgs_u8, torch.Tensor):
        imgs_u8 = torch.as_tensor(np.array(imgs_u8, copy=True))     # detect_face.py:28 copies too (fresh, aligned buffer)
    if imgs_u8.dim() == 3:
        imgs_u8 = imgs_u8.unsqueeze(0)
    with torch.no_grad():
        imgs = imgs_u8.permute(0, 3, 1, 2).float()
        B = imgs.shape[0]
        h, w = imgs.shape[2:4]
        scales = tap("scales", scale_pyramid(h, w, minsize, factor))

        # ---- stage 1 (detect_face.py:70-104)
        boxes, image_inds, scale_picks, level_inds = [], [], [], []
        offset = 0
        for li, scale in enumerate(scales):
            im_data = normalize(area_resize(imgs, level_size(h, w, scale)))
            tap("level%d" % li, im_data)
            reg, probs = nets.pnet_forward(pnet_sd, im_data)
            tap("pnet_reg%d" % li, reg); tap("pnet_prob%d" % li, probs)
            b_s, i_s = generate_bounding_box(reg, probs[:, 1], scale, threshold[0])
            tap("cand%d" % li, (b_s, i_s))
            boxes.append(b_s); image_inds.append(i_s)
            pick = batched_nms(b_s[:, :4].numpy(), b_s[:, 4].numpy(), i_s.numpy(), 0.5, faithful)
            tap("pick%d" % li, pick)
            scale_picks.append(torch.as_tensor(pick, dtype=torch.long) + offset)
            level_inds.append(torch.full((b_s.shape[0],), li, dtype=torch.long))
            offset += b_s.shape[0]
        if scales:
            boxes = torch.cat(boxes, dim=0); image_inds = torch.cat(image_inds, dim=0)
            scale_picks = torch.cat(scale_picks, dim=0)
        else:
            boxes = torch.zeros(0, 9); image_inds = torch.zeros(0, dtype=torch.long)
            scale_picks = torch.zeros(0, dtype=torch.long)
        boxes, image_inds = boxes[scale_picks], image_inds[scale_picks]
        pick = torch.as_tensor(batched_nms(boxes[:, :4].numpy(), boxes[:, 4].numpy(), image_inds.numpy(), 0.7, faithful),
                               dtype=torch.long)
        boxes, image_inds = boxes[pick], image_inds[pick]
        tap("stage1_nms", (boxes.clone(), image_inds.clone()))
        boxes = rerec(stage1_regress(boxes))
        tap("stage1_boxes", (boxes.clone(), image_inds.clone()))
        y, ey, x, ex = pad(boxes, w, h)

        # ---- stage 2 (detect_face.py:107-131)
        if len(boxes) > 0:
            im_data = crop_resize_batch(imgs, image_inds, y, ey, x, ex, 24)
            tap("rnet_in", im_data)
            out0, out1 = _batched(nets.rnet_forward, rnet_sd, im_data)
            tap("rnet_out", (out0, out1))
            score = out1[:, 1]
            ipass = score > threshold[1]
            boxes = torch.cat((boxes[ipass, :4], score[ipass].unsqueeze(1)), dim=1)
            image_inds = image_inds[ipass]
            mv = out0[ipass]
            pick = torch.as_tensor(batched_nms(boxes[:, :4].numpy(), boxes[:, 4].numpy(), image_inds.numpy(), 0.7,
                                               faithful), dtype=torch.long)
            boxes, image_inds, mv = boxes[pick], image_inds[pick], mv[pick]
            boxes = rerec(bbreg(boxes, mv))
        tap("stage2_boxes", (boxes.clone(), image_inds.clone()))

        # ---- stage 3 (detect_face.py:134-169)
        points = torch.zeros(0, 5, 2)
        if len(boxes) > 0:
            y, ey, x, ex = pad(boxes, w, h)
            im_data = crop_resize_batch(imgs, image_inds, y, ey, x, ex, 48)
            tap("onet_in", im_data)
            out0, out1, out2 = _batched(nets.onet_forward, onet_sd, im_data)
            tap("onet_out", (out0, out1, out2))
            score = out2[:, 1]
            ipass = score > threshold[2]
            pts = out1[ipass]
            boxes = torch.cat((boxes[ipass, :4], score[ipass].unsqueeze(1)), dim=1)
            image_inds = image_inds[ipass]
            mv = out0[ipass]
            w_i = boxes[:, 2] - boxes[:, 0] + 1
            h_i = boxes[:, 3] - boxes[:, 1] + 1
            points_x = w_i[:, None] * pts[:, 0:5] + boxes[:, 0:1] - 1
            points_y = h_i[:, None] * pts[:, 5:10] + boxes[:, 1:2] - 1
            points = torch.stack((points_x, points_y), dim=2)
            boxes = bbreg(boxes, mv)
            tap("stage3_pre_nms", (boxes.clone(), image_inds.clone(), points.clone()))
            pick = torch.as_tensor(batched_nms_min(boxes[:, :4].numpy(), boxes[:, 4].numpy(), image_inds.numpy(), 0.7,
                                                   "Min", faithful), dtype=torch.long)
            boxes, image_inds, points = boxes[pick], image_inds[pick], points[pick]

        boxes = boxes.numpy(); points = points.numpy(); image_inds = image_inds.numpy()
        batch_boxes, batch_points = [], []
        for b in range(B):
            sel = np.where(image_inds == b)
            batch_boxes.append(boxes[sel].copy())
            batch_points.append(points[sel].copy())
    return batch_boxes, batch_points


def mtcnn_detect(imgs_u8, sds, min_face_size=20, thresholds=(0.6, 0.7, 0.7), factor=0.709, select_largest=True,
                 faithful=True, taps=None):
    """MTCNN.detect post-processing (mtcnn.py:326-347) on a 4-D batch: returns lists (boxes (n,4), probs (n,),
    points (n,5,2)) per image; area-descending if select_largest else score-descending."""
    bb, pp = detect_face(imgs_u8, min_face_size, sds["pnet"], sds["rnet"], sds["onet"], list(thresholds), factor,
                         faithful, taps)
    boxes, probs, points = [], [], []
    for box, point in zip(bb, pp):
        if len(box) == 0:
            boxes.append(np.zeros((0, 4), f32)); probs.append(np.zeros((0,), f32)); points.append(np.zeros((0, 5, 2), f32))
            continue
        if select_largest:
            order = np.argsort((box[:, 2] - box[:, 0]) * (box[:, 3] - box[:, 1]))[::-1]
            box = box[order]; point = point[order]
        boxes.append(box[:, :4]); probs.append(box[:, 4]); points.append(point)
    return boxes, probs, points


# --------------------------------------------------------------------------------------------------------------------
# face extraction (MTCNN.extract path)
# --------------------------------------------------------------------------------------------------------------------
def extract_box(box, image_size, margin, raw_w, raw_h):
    """extract_face, detect_face.py:358-368: margin expansion, int-truncate, clamp."""
    m = [margin * (box[2] - box[0]) / (image_size - margin), margin * (box[3] - box[1]) / (image_size - margin)]
    return [int(max(box[0] - m[0] / 2, 0)), int(max(box[1] - m[1] / 2, 0)),
            int(min(box[2] + m[0] / 2, raw_w)), int(min(box[3] + m[1] / 2, raw_h))]


def extract_face_tensor(img_hwc_u8, box, image_size=160, margin=0, post_process=True):
    """extract_face for a torch.Tensor image (detect_face.py:317-322, :376) + fixed_image_standardization
    (mtcnn.py:516-518): crop -> area resize -> .byte() (truncation) -> float CHW -> (x-127.5)/128."""
    img = torch.as_tensor(img_hwc_u8)
    H, W = img.shape[:2]
    b = extract_box(box, image_size, margin, W, H)
    crop = img[b[1]:b[3], b[0]:b[2]]
    out = area_resize(crop.permute(2, 0, 1).unsqueeze(0).float(), (image_size, image_size)).byte().squeeze(0)
    face = out.float()
    if post_process:
        face = (face - 127.5) / 128.0
    return face


def extract_face_ndarray(img_hwc_u8, box, image_size=160, margin=0, post_process=True):
    """extract_face for a numpy image (detect_face.py:310-316): cv2.resize(INTER_AREA)."""
    import cv2
    H, W = img_hwc_u8.shape[:2]
    b = extract_box(box, image_size, margin, W, H)
    crop = img_hwc_u8[b[1]:b[3], b[0]:b[2]]
    out = cv2.resize(crop, (image_size, image_size), interpolation=cv2.INTER_AREA).copy()
    face = torch.from_numpy(np.float32(out)).permute(2, 0, 1).contiguous()
    if post_process:
        face = (face - 127.5) / 128.0
    return face
