"""GPU parity of the individual detection kernels against the oracle (bit-exact where the arithmetic is integer /
order-defined; fp32 tolerance for the P-Net convolutions)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "needs a CUDA device"
    return torch.device("cuda:0")


def _plan(B, H, W, minsize, factor=0.709):
    from vn_celeb_face_recognition_b200 import _lib
    p = _lib.Pyramid()
    _lib.call("vnfr_pyramid_plan", B, H, W, minsize, factor, C.byref(p))
    return p


# noise:HxW cases exercise the strip kernel's geometry: several column tiles (W*3/16 > 384 chunks), more than 10 levels
# (two launches), many strips per frame, and the one-task-per-row fallback (rows not 16-byte aligned / upscaling level)
@pytest.mark.parametrize("kind,n,minsize", [("small", 3, 50), ("small", 2, 20), ("1080p", 1, 50), ("1080p", 2, 20),
                                            ("noise:300x2560", 2, 40), ("noise:257x4112", 1, 16), ("noise:700x96", 3, 13),
                                            ("noise:120x482", 2, 30), ("noise:96x128", 2, 12)])
def test_pyramid_resize_bit_exact(dev, kind, n, minsize):
    from oracle import detect, synth
    from vn_celeb_face_recognition_b200 import _lib
    if kind.startswith("noise:"):
        h, w = [int(v) for v in kind[6:].split("x")]
        fr = np.random.RandomState(h * w).randint(0, 256, size=(n, h, w, 3)).astype(np.uint8)
        fr[0, : h // 2] = 255                                   # saturated columns: the u16 lanes must not overflow
    else:
        fr = synth.frames(kind, n)
    B, H, W, _ = fr.shape
    p = _plan(B, H, W, minsize)
    scales = detect.scale_pyramid(H, W, minsize, 0.709)
    assert p.n_levels == len(scales)
    assert [p.scale_d[i] for i in range(p.n_levels)] == scales            # fp64 scale list identical to the reference
    levels = torch.full((p.level_off[p.n_levels],), float("nan"), device=dev)
    d_fr = torch.from_numpy(fr).to(dev)
    _lib.call("vnfr_pyramid_resize_norm", C.byref(p), _lib.ptr(d_fr), _lib.ptr(levels), _lib.stream_ptr())
    torch.cuda.synchronize()
    x = torch.from_numpy(fr).permute(0, 3, 1, 2).float()
    for l, s in enumerate(scales):
        lh, lw = detect.level_size(H, W, s)
        assert (p.lh[l], p.lw[l]) == (lh, lw)
        ref = detect.normalize(detect.area_resize(x, (lh, lw)))
        got = levels[p.level_off[l]:p.level_off[l + 1]].view(B, 3, lh, lw).cpu()
        assert torch.equal(got, ref), "level %d differs: max %g" % (l, (got - ref).abs().max())


def _pack_pnet(sd):
    order = ["conv1.weight", "conv1.bias", "prelu1.weight", "conv2.weight", "conv2.bias", "prelu2.weight", "conv3.weight",
             "conv3.bias", "prelu3.weight", "conv4_1.weight", "conv4_1.bias", "conv4_2.weight", "conv4_2.bias"]
    return torch.cat([sd[k].float().reshape(-1) for k in order]).contiguous()


@pytest.mark.parametrize("kind,n,minsize", [("small", 3, 50), ("small", 2, 20)])
def test_pnet_sweep_matches_oracle(dev, kind, n, minsize):
    from oracle import detect, nets, synth
    from vn_celeb_face_recognition_b200 import _lib
    sds = synth.mtcnn_state_dicts()
    fr = synth.frames(kind, n)
    B, H, W, _ = fr.shape
    p = _plan(B, H, W, minsize)
    L = p.n_levels
    w = _pack_pnet(sds["pnet"])
    wdev = _lib.pack_pnet_weights(w, dev)                        # caller-owned device copy: the library keeps no weights
    levels = torch.empty(p.level_off[L], device=dev)
    d_fr = torch.from_numpy(fr).to(dev)
    _lib.call("vnfr_pyramid_resize_norm", C.byref(p), _lib.ptr(d_fr), _lib.ptr(levels), _lib.stream_ptr())
    cap = 4096
    cnt = torch.zeros(B * L, dtype=torch.int32, device=dev)
    cell = torch.zeros(B * L, cap, dtype=torch.int32, device=dev)
    score = torch.zeros(B * L, cap, device=dev)
    reg = torch.zeros(B * L, cap, 4, device=dev)
    dprob = torch.full((p.map_off[L],), float("nan"), device=dev)
    dreg = torch.full((4 * p.map_off[L],), float("nan"), device=dev)
    _lib.call("vnfr_pnet_sweep_compact", C.byref(p), _lib.ptr(levels), _lib.ptr(wdev), 0.6, cap, _lib.ptr(cnt), _lib.ptr(cell), _lib.ptr(score),
              _lib.ptr(reg), _lib.ptr(dprob), _lib.ptr(dreg), _lib.stream_ptr())
    torch.cuda.synchronize()
    taps = {}
    detect.detect_face(fr, minsize, sds["pnet"], sds["rnet"], sds["onet"], [0.6, 0.7, 0.7], 0.709, taps=taps)
    cnt = cnt.cpu().numpy(); cell = cell.cpu().numpy(); score = score.cpu().numpy(); reg = reg.cpu().numpy()
    n_border = 0
    for l in range(L):
        oh, ow = p.oh[l], p.ow[l]
        rp, rr = taps["pnet_prob%d" % l], taps["pnet_reg%d" % l]
        assert tuple(rp.shape[2:]) == (oh, ow)
        got_p = dprob[p.map_off[l]:p.map_off[l + 1]].view(B, oh, ow).cpu()
        got_r = dreg[4 * p.map_off[l]:4 * p.map_off[l + 1]].view(B, 4, oh, ow).cpu()
        assert (got_p - rp[:, 1]).abs().max().item() < 2e-5, "prob map level %d" % l
        assert (got_r - rr).abs().max().item() < 2e-4, "reg map level %d" % l
        boxes, inds = taps["cand%d" % l]
        for b in range(B):
            seg = b * L + l
            k = cnt[seg]
            assert k <= cap
            got = {(int(c) >> 16, int(c) & 0xFFFF): (s, r) for c, s, r in zip(cell[seg, :k], score[seg, :k], reg[seg, :k])}
            assert len(got) == k, "duplicate cells"
            # every compacted candidate carries exactly the dense map's values at its cell
            for (y, x), (s, r) in got.items():
                assert s == got_p[b, y, x].item() and np.array_equal(r, got_r[b, :, y, x].numpy())
                assert s >= np.float32(0.6)
            ref_mask = (rp[b, 1] >= 0.6)
            ref_cells = {(int(y), int(x)) for y, x in ref_mask.nonzero().tolist()}
            diff = ref_cells.symmetric_difference(got.keys())
            # cells whose score is within fp32 conv noise of the threshold may flip; everything else must agree
            for (y, x) in diff:
                assert abs(rp[b, 1, y, x].item() - 0.6) < 2e-5
            n_border += len(diff)
    assert n_border <= 2


def _rand_boxes(rng, n, ties=True):
    b = rng.rand(n, 4).astype(np.float32) * 300
    b[:, 2:] = b[:, :2] + rng.rand(n, 2).astype(np.float32) * 90 + 1
    s = rng.rand(n).astype(np.float32)
    if ties:
        s = np.round(s, 2)
    return b, s


@pytest.mark.parametrize("mode", [0, 1])
def test_nms_segments_bit_exact(dev, mode):
    from oracle import detect
    from vn_celeb_face_recognition_b200 import _lib
    rng = np.random.RandomState(3)
    sizes = [0, 1, 2, 63, 64, 65, 130, 700, 1500]
    cap = 2048
    nseg = len(sizes)
    boxes = np.zeros((nseg, cap, 4), np.float32); scores = np.zeros((nseg, cap), np.float32)
    for i, n in enumerate(sizes):
        boxes[i, :n], scores[i, :n] = _rand_boxes(rng, n)
        if n >= 64:      # a dense cluster so that suppression chains across 64-box chunks
            boxes[i, :n // 2, :2] = boxes[i, 0, :2] + rng.rand(n // 2, 2) * 30
            boxes[i, :n // 2, 2:] = boxes[i, :n // 2, :2] + 60 + rng.rand(n // 2, 2) * 10
    d_b = torch.from_numpy(boxes).to(dev); d_s = torch.from_numpy(scores).to(dev)
    d_c = torch.tensor(sizes, dtype=torch.int32, device=dev)
    kc = torch.zeros(nseg, dtype=torch.int32, device=dev)
    keep = torch.full((nseg, cap), -1, dtype=torch.int32, device=dev)
    thr = 0.5 if mode == 0 else 0.7
    _lib.call("vnfr_nms_segments", nseg, cap, _lib.ptr(d_c), _lib.ptr(d_b), _lib.ptr(d_s), thr, mode, _lib.ptr(kc), _lib.ptr(keep),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    kc = kc.cpu().numpy(); keep = keep.cpu().numpy()
    for i, n in enumerate(sizes):
        if mode == 0:
            ref = detect.nms_iou(boxes[i, :n], scores[i, :n], thr)
        else:
            ref = detect.nms_min(boxes[i, :n], scores[i, :n], thr, "Min", tie="stable")
        assert kc[i] == len(ref), "segment %d: kept %d vs %d" % (i, kc[i], len(ref))
        np.testing.assert_array_equal(keep[i, :kc[i]], ref)
    if mode == 0:
        import torchvision
        ref_tv = torchvision.ops.nms(torch.from_numpy(boxes[7, :700]), torch.from_numpy(scores[7, :700]), thr).numpy()
        np.testing.assert_array_equal(keep[7, :kc[7]], ref_tv)
