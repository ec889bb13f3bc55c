"""Seeded synthetic inputs of the BASELINE.json configs and weight loading helpers: a DATA
generator shared by bench.py and the tests, no algorithm of the path lives here.  Random-noise frames produce zero
detections, so frames are built by pasting the 20 bundled face crops (data assets under assets/faces) on a noisy grey
canvas.
"""
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(_HERE, "assets")
WEIGHTS = os.path.join(_HERE, "models", "weights_mtcnn")

_faces = None


def bundled_faces():
    """The 20 bundled RGB face crops (18x 181x181, 2x 127x127), sorted by file name (reference data/*.png)."""
    global _faces
    if _faces is None:
        from PIL import Image
        d = os.path.join(ASSETS, "faces")
        names = sorted(f for f in os.listdir(d) if f.endswith(".png"))
        _faces = [(n, np.asarray(Image.open(os.path.join(d, n)).convert("RGB"))) for n in names]
    return _faces


def _resize_u8(img, k):
    import cv2
    return cv2.resize(img, (k, k), interpolation=cv2.INTER_AREA if k < img.shape[0] else cv2.INTER_LINEAR)


def frame_1080p(frame_seed, h=1080, w=1920, n_faces=12, face_px=181, pitch=201):
    """Config 3 frame: grey 96 + uniform noise [0,16) canvas, ``n_faces`` bundled faces pasted at ``face_px`` on a
    ``pitch`` grid from (10, 10); face choice by rng.randint(20); seed = global frame index."""
    rng = np.random.RandomState(frame_seed)
    canvas = (96 + rng.randint(0, 16, size=(h, w, 3))).astype(np.uint8)
    faces = bundled_faces()
    per_row = max(1, (w - 10) // pitch)
    for i in range(n_faces):
        r, c = divmod(i, per_row)
        y0, x0 = 10 + r * pitch, 10 + c * pitch
        if y0 + face_px > h or x0 + face_px > w:
            break
        f = faces[rng.randint(len(faces))][1]
        if f.shape[0] != face_px:
            f = _resize_u8(f, face_px)
        canvas[y0:y0 + face_px, x0:x0 + face_px] = f
    return canvas


def frame_crowded(frame_seed, h=2160, w=3840, n_faces=50, sizes=(30, 60, 120)):
    """Config 4 frame: same canvas, ``n_faces`` faces resized to k x k (k cycling through ``sizes``), pitch k + 20."""
    rng = np.random.RandomState(frame_seed)
    canvas = (96 + rng.randint(0, 16, size=(h, w, 3))).astype(np.uint8)
    faces = bundled_faces()
    x, y, row_h = 10, 10, 0
    for i in range(n_faces):
        k = sizes[i % len(sizes)]
        if x + k + 20 > w:
            x = 10
            y += row_h + 20
            row_h = 0
        if y + k > h:
            break
        f = _resize_u8(faces[rng.randint(len(faces))][1], k)
        canvas[y:y + k, x:x + k] = f
        x += k + 20
        row_h = max(row_h, k)
    return canvas


def small_frame(frame_seed, h=360, w=480, n_faces=3, face_px=100):
    """Small multi-face frame for fast CPU/GPU parity tests."""
    return frame_1080p(frame_seed, h=h, w=w, n_faces=n_faces, face_px=face_px, pitch=face_px + 30)


def frames(kind, n, first_seed=0):
    fn = {"1080p": frame_1080p, "4k": frame_crowded, "small": small_frame}[kind]
    return np.stack([fn(first_seed + i) for i in range(n)])


def crops_160(n, seed=1):
    """Config 2 input: (n,3,160,160) fp32 standardised crops = bundled faces resized to 160, random flips/shifts."""
    import cv2
    rng = np.random.RandomState(seed)
    faces = bundled_faces()
    out = np.empty((n, 3, 160, 160), dtype=np.float32)
    for i in range(n):
        f = faces[rng.randint(len(faces))][1]
        f = cv2.resize(f, (160, 160), interpolation=cv2.INTER_AREA)
        if rng.randint(2):
            f = f[:, ::-1]
        f = np.roll(f, (rng.randint(-8, 9), rng.randint(-8, 9)), axis=(0, 1))
        f = np.clip(f.astype(np.int32) + rng.randint(-10, 11), 0, 255)
        out[i] = ((f.astype(np.float32) - 127.5) / 128.0).transpose(2, 0, 1)
    return torch.from_numpy(out)


def mtcnn_state_dicts():
    """The bundled MTCNN weights (reference models/weights_mtcnn/*.pt, loaded at mtcnn.py:32-36, 78-82, 132-136)."""
    return {n: torch.load(os.path.join(WEIGHTS, n + ".pt"), map_location="cpu") for n in ("pnet", "rnet", "onet")}
