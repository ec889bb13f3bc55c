import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def ragged(g, prefix):
    n = g[prefix + "_n"]
    return [g["%s_%d" % (prefix, i)] for i in range(len(n))]


def golden_encoder_state_dict():
    """Seed-0 conditioned encoder weights with the BN running statistics pinned from the golden file (so the weights
    are bit-identical on every host, independent of the CPU conv kernels used during calibration)."""
    import torch
    from oracle import nets
    g = load_golden("encoder_seed0")
    sd = nets.make_encoder_state_dict(seed=0, calibrate=False)
    off = 0
    vals = g["bn_vals"]
    for k in g["bn_keys"]:
        k = str(k)
        n = sd[k].numel()
        sd[k] = torch.from_numpy(vals[off:off + n].copy()).reshape(sd[k].shape)
        off += n
    assert off == len(vals)
    return sd


def iou_matrix(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter)


def assert_boxes_match(got, want, min_iou=0.99):
    """Identical face counts and a one-to-one matching at IoU >= min_iou (north-star tolerance), in the same order."""
    got = np.asarray(got, dtype=np.float32).reshape(-1, 4)
    want = np.asarray(want, dtype=np.float32).reshape(-1, 4)
    assert got.shape == want.shape, "face count %d != %d" % (len(got), len(want))
    if len(got):
        iou = iou_matrix(got, want)
        assert np.all(np.diag(iou) >= min_iou), "box IoU %s" % np.diag(iou)
