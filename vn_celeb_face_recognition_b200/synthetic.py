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


# --------------------------------------------------------------------------------------------------------------------
# seeded random-init weights (BASELINE.json configs use random-init encoder / MLP weights): DATA, shared by bench.py, the
# tests and the oracle (oracle/nets.py re-exports these; the BN calibration that needs a forward pass lives there)
# --------------------------------------------------------------------------------------------------------------------
def encoder_conv_specs():
    """(prefix, cin, cout, (kh,kw), has_bn) for every conv in construction order of inception_resnet_v1.py:219-254."""
    specs = []

    def b(p, cin, cout, k):
        k = (k, k) if isinstance(k, int) else k
        specs.append((p, cin, cout, k, True))

    b("conv2d_1a", 3, 32, 3); b("conv2d_2a", 32, 32, 3); b("conv2d_2b", 32, 64, 3)
    b("conv2d_3b", 64, 80, 1); b("conv2d_4a", 80, 192, 3); b("conv2d_4b", 192, 256, 3)
    for i in range(5):
        p = "repeat_1.%d" % i
        b(p + ".branch0", 256, 32, 1)
        b(p + ".branch1.0", 256, 32, 1); b(p + ".branch1.1", 32, 32, 3)
        b(p + ".branch2.0", 256, 32, 1); b(p + ".branch2.1", 32, 32, 3); b(p + ".branch2.2", 32, 32, 3)
        specs.append((p + ".conv2d", 96, 256, (1, 1), False))
    b("mixed_6a.branch0", 256, 384, 3)
    b("mixed_6a.branch1.0", 256, 192, 1); b("mixed_6a.branch1.1", 192, 192, 3); b("mixed_6a.branch1.2", 192, 256, 3)
    for i in range(10):
        p = "repeat_2.%d" % i
        b(p + ".branch0", 896, 128, 1)
        b(p + ".branch1.0", 896, 128, 1); b(p + ".branch1.1", 128, 128, (1, 7)); b(p + ".branch1.2", 128, 128, (7, 1))
        specs.append((p + ".conv2d", 256, 896, (1, 1), False))
    b("mixed_7a.branch0.0", 896, 256, 1); b("mixed_7a.branch0.1", 256, 384, 3)
    b("mixed_7a.branch1.0", 896, 256, 1); b("mixed_7a.branch1.1", 256, 256, 3)
    b("mixed_7a.branch2.0", 896, 256, 1); b("mixed_7a.branch2.1", 256, 256, 3); b("mixed_7a.branch2.2", 256, 256, 3)
    for i in list(range(5)) + [None]:
        p = "repeat_3.%d" % i if i is not None else "block8"
        b(p + ".branch0", 1792, 192, 1)
        b(p + ".branch1.0", 1792, 192, 1); b(p + ".branch1.1", 192, 192, (1, 3)); b(p + ".branch1.2", 192, 192, (3, 1))
        specs.append((p + ".conv2d", 384, 1792, (1, 1), False))
    return specs



def encoder_state_dict_uncalibrated(seed=0):
    """Seeded random-init InceptionResnetV1 ``state_dict`` with the reference's keys: kaiming-normal
    (fan_in, relu) conv weights, BN affine drawn near (1, 0), BN running statistics (0, 1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for p, cin, cout, (kh, kw), has_bn in encoder_conv_specs():
        fan_in = cin * kh * kw
        w = torch.randn(cout, cin, kh, kw, generator=g) * (2.0 / fan_in) ** 0.5
        if has_bn:
            sd[p + ".conv.weight"] = w
            sd[p + ".bn.weight"] = 1.0 + 0.1 * torch.randn(cout, generator=g)
            sd[p + ".bn.bias"] = 0.1 * torch.randn(cout, generator=g)
            sd[p + ".bn.running_mean"] = torch.zeros(cout)
            sd[p + ".bn.running_var"] = torch.ones(cout)
            sd[p + ".bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        else:
            sd[p + ".weight"] = w
            sd[p + ".bias"] = 0.1 * torch.randn(cout, generator=g)
    sd["last_linear.weight"] = torch.randn(512, 1792, generator=g) * (1.0 / 1792) ** 0.5
    sd["last_bn.weight"] = 1.0 + 0.1 * torch.randn(512, generator=g)
    sd["last_bn.bias"] = 0.1 * torch.randn(512, generator=g)
    sd["last_bn.running_mean"] = torch.zeros(512)
    sd["last_bn.running_var"] = torch.ones(512)
    sd["last_bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def encoder_state_dict_seed0():
    """The seed-0 weights with the BN running statistics of a seeded calibration batch (assets/encoder_bn_seed0.npz, written
    by oracle/make_golden.py: default PyTorch init + eval-mode BN is DEGENERATE -- every input maps to the same embedding,
    SURVEY.md section 4.5 -- so parity and label agreement would be meaningless on it).  Bit-identical on every host."""
    g = np.load(os.path.join(ASSETS, "encoder_bn_seed0.npz"), allow_pickle=False)
    sd = encoder_state_dict_uncalibrated(0)
    off, vals = 0, g["bn_vals"]
    for k in g["bn_keys"]:
        k = str(k)
        n = sd[k].numel()
        sd[k] = torch.from_numpy(vals[off:off + n].copy()).reshape(sd[k].shape)
        off += n
    assert off == len(vals)
    return sd


def mlp_state_dict(num_classes=1001, input_dim=512, seed=0):
    """MLPModel ``state_dict`` (models/mlp_model.py:6-8) with a wide seeded init so argmax depends on the input."""
    g = torch.Generator().manual_seed(1000 + seed)
    return {
        "dense_1.weight": torch.randn(2048, input_dim, generator=g) * (2.0 / input_dim) ** 0.5 * 4.0,
        "dense_1.bias": 0.1 * torch.randn(2048, generator=g),
        "dense_2.weight": torch.randn(num_classes, 2048, generator=g) * (1.0 / 2048) ** 0.5 * 4.0,
        "dense_2.bias": 0.1 * torch.randn(num_classes, generator=g),
    }
