"""CPU restatement of the 5-point similarity alignment used by the demo_video path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  * align_face.py:12-48  -- ``center_point_dict`` landmark templates,
  * align_face.py:51-57  -- ``alignment``: SimilarityTransform().estimate(dst, src); cv2.warpAffine(img, M[0:2], (w, h),
                            borderValue=0.0) (cv2 default flags = INTER_LINEAR, BORDER_CONSTANT),
  * demo_image.py:174-199 -- ``get_face_from_boxes`` integer crop,
  * demo_image.py:236-239 -- ``move_landmark_to_box`` (uses the UNclamped float corner).

``skimage.transform.SimilarityTransform`` is a third-party dependency that is absent from /root/reference and from this
image (no version pin anywhere in the reference).  Its published algorithm is Umeyama 1991, "Least-squares estimation
of transformation parameters between two point patterns", with scale estimation (skimage ``_umeyama(src, dst,
estimate_scale=True)``); that is what ``umeyama`` below restates.  Parity for this row is anchored on the reference's
call site (align_face.py:52-54) -- "parity unpinned" for the skimage part, see DESIGN.md.
"""
import numpy as np

CENTER_POINTS = {
    (96, 112): np.array([[30.2946, 51.6963], [65.5318, 51.5014], [48.0252, 71.7366], [33.5493, 92.3655],
                         [62.7299, 92.2041]], dtype=np.float32),
    (112, 112): np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366], [41.5493, 92.3655],
                          [70.7299, 92.2041]], dtype=np.float32),
    (150, 150): np.array([[51.287415, 69.23612], [98.48009, 68.97509], [75.03375, 96.075806],
                          [55.646385, 123.7038], [94.72754, 123.48763]], dtype=np.float32),
    (160, 160): np.array([[54.706573, 73.85186], [105.045425, 73.573425], [80.036, 102.48086],
                          [59.356144, 131.95071], [101.04271, 131.72014]], dtype=np.float32),
    (224, 224): np.array([[76.589195, 103.3926], [147.0636, 103.0028], [112.0504, 143.4732],
                          [83.098595, 184.731], [141.4598, 184.4082]], dtype=np.float32),
}


def umeyama(src, dst, estimate_scale=True):
    """Umeyama 1991 eq. 34-43.  src, dst: (N, 2).  Returns the 3x3 homogeneous matrix T with dst ~= T @ [src, 1]."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    num, dim = src.shape
    src_mean = src.mean(axis=0)
    dst_mean = dst.mean(axis=0)
    src_demean = src - src_mean
    dst_demean = dst - dst_mean
    A = dst_demean.T @ src_demean / num
    d = np.ones((dim,), dtype=np.float64)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1, dtype=np.float64)
    U, S, V = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    elif rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V
    if estimate_scale:
        scale = 1.0 / src_demean.var(axis=0).sum() * (S @ d)
    else:
        scale = 1.0
    T[:dim, dim] = dst_mean - scale * (T[:dim, :dim] @ src_mean.T)
    T[:dim, :dim] *= scale
    return T


class SimilarityTransform:
    """Only the surface align_face.py:52-54 uses: ``estimate(src, dst)`` and ``params`` (3x3)."""

    def __init__(self):
        self.params = np.eye(3)

    def estimate(self, src, dst):
        self.params = umeyama(src, dst, True)
        return True


def alignment(cv_img, template, landmarks, dst_w, dst_h):
    """align_face.py:51-57.  ``template`` = center points (the reference's ``src``), ``landmarks`` = its ``dst``."""
    import cv2
    M = umeyama(landmarks, template, True)[0:2, :]
    return cv2.warpAffine(cv_img, M, (dst_w, dst_h), borderValue=0.0)


def get_face_from_boxes(image, boxes):
    """demo_image.py:174-199 with box_requirements=None."""
    faces, idx = [], []
    ori_h, ori_w = image.shape[:2]
    for i, box in enumerate(boxes):
        x1 = max(int(box[0]), 0)
        y1 = max(int(box[1]), 0)
        x2 = min(int(box[2] + 1), ori_w)
        y2 = min(int(box[3] + 1), ori_h)
        faces.append(image[y1:y2, x1:x2, :])
        idx.append(i)
    return faces, idx


def warp_affine_u8(src, M, dst_w, dst_h):
    """Restatement of cv2.warpAffine(src u8 HxWxC, M 2x3 double, (dst_w, dst_h), INTER_LINEAR, BORDER_CONSTANT 0).

    OpenCV (unpinned by the reference; 4.13 in this image) computes, per destination pixel, fixed-point source
    coordinates with AB_BITS=10 and INTER_BITS=5, then blends the 2x2 neighbourhood with int16 weights that sum to
    2^15 (INTER_REMAP_COEF_BITS).  This restatement is validated bit-for-bit against cv2 in tests/test_oracle_align.py
    and is what the CUDA warp kernel mirrors.
    """
    AB_BITS, INTER_BITS = 10, 5
    AB_SCALE = 1 << AB_BITS
    INTER_TAB = 1 << INTER_BITS
    src = np.asarray(src)
    H, W = src.shape[:2]
    C = src.shape[2] if src.ndim == 3 else 1
    s3 = src.reshape(H, W, C).astype(np.int64)
    M = np.asarray(M, dtype=np.float64)
    # invertAffineTransform
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    A12, A21 = -M[0, 1] * D, -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    Mi = np.array([[A11, A12, b1], [A21, A22, b2]], dtype=np.float64)

    xs = np.arange(dst_w, dtype=np.float64)
    ys = np.arange(dst_h, dtype=np.float64)
    adelta = np.rint(Mi[0, 0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(Mi[1, 0] * xs * AB_SCALE).astype(np.int64)
    round_delta = AB_SCALE // INTER_TAB // 2
    X0 = np.rint((Mi[0, 1] * ys + Mi[0, 2]) * AB_SCALE).astype(np.int64) + round_delta
    Y0 = np.rint((Mi[1, 1] * ys + Mi[1, 2]) * AB_SCALE).astype(np.int64) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = X & (INTER_TAB - 1)
    fy = Y & (INTER_TAB - 1)

    wtab = bilinear_tab_s16()
    w = wtab[fy, fx]  # (h, w, 4) int: [ (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx ]

    def px(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = s3[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None], v, 0)

    acc = (px(sy, sx) * w[..., 0:1] + px(sy, sx + 1) * w[..., 1:2] +
           px(sy + 1, sx) * w[..., 2:3] + px(sy + 1, sx + 1) * w[..., 3:4])
    out = (acc + (1 << 14)) >> 15
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out.reshape(dst_h, dst_w, C) if src.ndim == 3 else out.reshape(dst_h, dst_w)


_TAB = None


def bilinear_tab_s16():
    """OpenCV imgwarp.cpp initInterTab2D(INTER_LINEAR, fixpt=True): 32x32 table of 4 int16 weights summing to 32768."""
    global _TAB
    if _TAB is not None:
        return _TAB
    N, SCALE = 32, 1 << 15
    t1 = np.zeros((N, 2), dtype=np.float32)
    for i in range(N):
        x = np.float32(i) * np.float32(1.0 / N)
        t1[i, 0] = np.float32(1.0) - x
        t1[i, 1] = x
    tab = np.zeros((N, N, 4), dtype=np.int64)
    for i in range(N):
        for j in range(N):
            f = np.array([t1[i, 0] * t1[j, 0], t1[i, 0] * t1[j, 1], t1[i, 1] * t1[j, 0], t1[i, 1] * t1[j, 1]],
                         dtype=np.float32)
            iv = np.clip(np.rint(f * np.float32(SCALE)), -32768, 32767).astype(np.int64)
            isum = int(iv.sum())
            if isum != SCALE:
                diff = isum - SCALE
                # OpenCV adjusts the largest (diff<0) or smallest... it walks the central 2x2 of the ksize window:
                # for ksize=2 that is all four taps; it adds to the max when diff < 0, subtracts from the min... see
                # imgwarp.cpp: "if (diff < 0) itab[ksize2*ksize+ksize2] ... " -> restated as: pick argmax/argmin.
                k = 2
                mk1 = mk2 = 0  # argmin position
                Mk1 = Mk2 = 0  # argmax position
                m2 = iv.reshape(2, 2)
                for k1 in range(0, 2):
                    for k2 in range(0, 2):
                        if m2[k1, k2] < m2[mk1, mk2]:
                            mk1, mk2 = k1, k2
                        elif m2[k1, k2] > m2[Mk1, Mk2]:
                            Mk1, Mk2 = k1, k2
                if diff < 0:
                    m2[Mk1, Mk2] -= diff
                else:
                    m2[mk1, mk2] -= diff
                iv = m2.reshape(4)
            tab[i, j] = iv
    _TAB = tab
    return tab
