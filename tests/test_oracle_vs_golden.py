"""Pins the oracle (CPU restatement) against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only.

In the build container the oracle reproduces the reference bit for bit.  The tolerances below (boxes 2e-3 px, probs
1e-5) only absorb the CPU conv kernels' dependence on ISA / buffer alignment (oneDNN picks different accumulation
orders on different hosts); face counts, orders and integer results are compared exactly."""

BOX_ATOL, PROB_ATOL = 2e-3, 1e-5


def close(a, b, atol):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=0, atol=atol)

import os

import numpy as np
import pytest
import torch

from conftest import load_golden, ragged, golden_encoder_state_dict, assert_boxes_match
from oracle import detect, nets, synth, align, pipeline


@pytest.fixture(scope="module")
def sds():
    return synth.mtcnn_state_dicts()


@pytest.mark.parametrize("name,kind,n,seed", [("detect_small_min50", "small", 3, 0), ("detect_small_min20", "small", 2, 7)])
def test_detect_matches_reference_bit_exact(sds, name, kind, n, seed):
    g = load_golden(name)
    fr = synth.frames(kind, n, first_seed=seed)
    for sl, tag in ((True, "largest"), (False, "prob")):
        b, p, l = detect.mtcnn_detect(fr, sds, min_face_size=int(g["min_face_size"]), select_largest=sl)
        for i in range(n):
            close(b[i], ragged(g, "boxes_" + tag)[i], BOX_ATOL)
            close(p[i], ragged(g, "probs_" + tag)[i], PROB_ATOL)
            close(l[i], ragged(g, "points_" + tag)[i], BOX_ATOL)


def test_detect_bundled_one_face_each(sds):
    g = load_golden("detect_bundled")
    assert np.all(g["boxes_n"] == 1)          # exactly one face in each bundled PNG
    for i, (_, img) in enumerate(synth.bundled_faces()[:6]):
        b, p, l = detect.mtcnn_detect(img[None], sds, min_face_size=50)
        close(b[0], ragged(g, "boxes")[i], BOX_ATOL)
        close(p[0], ragged(g, "probs")[i], PROB_ATOL)
        close(l[0], ragged(g, "points")[i], BOX_ATOL)


def test_detect_per_image_nms_variant_equals_faithful(sds):
    """faithful=False (per-image NMS, the CUDA semantics) gives the same faces as the torchvision offset trick."""
    fr = synth.frames("small", 3)
    a = detect.mtcnn_detect(fr, sds, min_face_size=50, faithful=True)
    b = detect.mtcnn_detect(fr, sds, min_face_size=50, faithful=False)
    for i in range(3):
        assert_boxes_match(b[0][i], a[0][i], 0.999)
        np.testing.assert_allclose(b[1][i], a[1][i], atol=1e-6)


def test_area_resize_restatement_bit_exact():
    rng = np.random.RandomState(0)
    for (H, W, oh, ow) in [(270, 480, 65, 116), (37, 53, 24, 24), (13, 9, 24, 24), (181, 181, 48, 48), (97, 120, 160, 160)]:
        img = rng.randint(0, 256, size=(3, H, W)).astype(np.uint8)
        a = detect.area_resize(torch.from_numpy(img).float()[None], (oh, ow))[0].numpy()
        np.testing.assert_array_equal(a, detect.area_resize_np(img, oh, ow))


def test_nms_restatement_matches_torchvision():
    import torchvision
    rng = np.random.RandomState(1)
    for n in [1, 2, 17, 300]:
        b = rng.rand(n, 4).astype(np.float32) * 200
        b[:, 2:] = b[:, :2] + rng.rand(n, 2).astype(np.float32) * 80 + 1
        s = np.round(rng.rand(n).astype(np.float32), 2)          # forces score ties
        for thr in (0.5, 0.7):
            k = torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), thr).numpy()
            np.testing.assert_array_equal(k, detect.nms_iou(b, s, thr))
        idx = rng.randint(0, 3, size=n)
        k = torchvision.ops.batched_nms(torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(idx), 0.7).numpy()
        np.testing.assert_array_equal(k, detect.batched_nms(b, s, idx, 0.7))


def test_warp_affine_restatement_bit_exact_vs_cv2():
    import cv2
    rng = np.random.RandomState(2)
    for _ in range(8):
        H, W = rng.randint(40, 200), rng.randint(40, 200)
        src = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
        ang, sc = rng.uniform(-0.5, 0.5), rng.uniform(0.5, 2.5)
        M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-20, 20)],
                      [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-20, 20)]])
        np.testing.assert_array_equal(cv2.warpAffine(src, M, (160, 160), borderValue=0.0),
                                      align.warp_affine_u8(src, M, 160, 160))


def test_extract_face_matches_reference():
    g = load_golden("extract_small")
    fr = synth.frames("small", 2)
    for i in range(2):
        boxes = g["boxes_%d" % i]
        ft = torch.stack([detect.extract_face_tensor(fr[i], b) for b in boxes])
        fn = torch.stack([detect.extract_face_ndarray(fr[i], b) for b in boxes])
        np.testing.assert_array_equal(ft.numpy(), g["faces_tensor_%d" % i])
        np.testing.assert_array_equal(fn.numpy(), g["faces_ndarray_%d" % i])


def test_encoder_and_mlp_match_reference():
    g = load_golden("encoder_seed0")
    sd = golden_encoder_state_dict()
    mlp = nets.make_mlp_state_dict(1001, seed=0)
    x = synth.crops_160(8, seed=1)
    with torch.no_grad():
        e = nets.encoder_forward(sd, x)
        lp = nets.mlp_forward(mlp, e)
    np.testing.assert_allclose(e.numpy(), g["emb"], atol=2e-5)
    assert lp.argmax(1).tolist() == g["logp_argmax"].tolist()
    np.testing.assert_allclose(lp[:, :16].numpy(), g["logp_head"], atol=2e-3)
    # well-conditioned: embeddings differ between inputs and labels are diverse
    cos = (e @ e.T).numpy()
    assert cos[~np.eye(8, dtype=bool)].max() < 0.9
    assert len(set(g["logp_argmax"].tolist())) >= 5


def test_demo_video_path_matches_reference():
    g = load_golden("demo_video_small")
    sds = synth.mtcnn_state_dicts()
    fr = synth.frames("small", 2, first_seed=3)
    faces, boxes = pipeline.parallel_detect_and_align(list(fr), sds, align.CENTER_POINTS[(160, 160)], (160, 160))
    labels, _ = pipeline.recognize(faces, golden_encoder_state_dict(), nets.make_mlp_state_dict(1001, seed=0))
    for i in range(2):
        d = np.abs(np.stack(faces[i]).astype(int) - g["aligned_%d" % i].astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3      # u8 warp of boxes equal to 2e-3 px
        close(np.asarray(boxes[i], np.float32), g["boxes_%d" % i], BOX_ATOL)
        assert labels[i] == g["labels_%d" % i].tolist()


def test_cal_embedding_matches_reference(tmp_path):
    import torchvision.transforms as tf
    g = load_golden("find_embedding_bundled")
    tr = tf.Compose([tf.Resize(160), pipeline.transforms_default])
    pipeline.cal_embedding(os.path.join(synth.ASSETS, "faces"), 64, golden_encoder_state_dict(), tr, str(tmp_path))
    files = sorted(os.listdir(tmp_path))
    assert files == [str(f) for f in g["files"]]
    embs = np.stack([np.load(os.path.join(tmp_path, f))["arr_0"] for f in files])
    assert embs.shape == (20, 512) and embs.dtype == np.float32
    np.testing.assert_allclose(embs, g["emb"], atol=2e-5)
