"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/vnfr_b200.h declares, the
host-side entry points behave (pyramid plan = reference's scale list, argument validation), the drop-in classes keep
the reference's state_dict surface, and the product refuses to run without CUDA instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vn_celeb_face_recognition_b200 import _lib, build
    build.build()           # no-op when up to date; compiles with nvcc otherwise
    return _lib


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "vnfr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vnfr_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    l = lib.lib()
    for sym in sorted(declared):
        assert hasattr(l, sym), "libvnfr_b200.so does not export %s" % sym
    assert set(lib.exported_symbols()) <= declared | {"vnfr_last_error"}
    assert l.vnfr_version() >= 100


@pytest.mark.parametrize("H,W,minsize,n_levels,cells", [(1080, 1920, 50, 9, 54797), (1080, 1920, 20, 12, 362055),
                                                        (2160, 3840, 20, 14, 1474378), (181, 181, 50, 4, 455)])
def test_pyramid_plan_matches_reference_scale_list(lib, H, W, minsize, n_levels, cells):
    """vnfr_pyramid_plan is host code: identical fp64 scale list / level sizes to detect_face.py:48-60, :71, and the
    P-Net map geometry of SURVEY.md Appendix C."""
    from oracle import detect
    p = lib.Pyramid()
    lib.call("vnfr_pyramid_plan", 3, H, W, minsize, 0.709, C.byref(p))
    scales = detect.scale_pyramid(H, W, minsize, 0.709)
    assert p.n_levels == n_levels == len(scales)
    assert [p.scale_d[i] for i in range(n_levels)] == scales
    assert [(p.lh[i], p.lw[i]) for i in range(n_levels)] == [detect.level_size(H, W, s) for s in scales]
    assert sum(p.oh[i] * p.ow[i] for i in range(n_levels)) == cells
    assert p.map_off[n_levels] == 3 * cells and p.level_off[n_levels] == 3 * 3 * sum(p.lh[i] * p.lw[i] for i in range(n_levels))


def test_argument_validation_reports_errors(lib):
    p = lib.Pyramid()
    with pytest.raises(lib.VnfrError, match="bad arguments"):
        lib.call("vnfr_pyramid_plan", 1, 0, 100, 20, 0.709, C.byref(p))
    with pytest.raises(lib.VnfrError, match="cap"):
        lib.call("vnfr_nms_segments", 1, 100000, None, None, None, 0.5, 0, None, None, None)
    op = lib.ConvOp()
    op.cin = 3                                           # not a multiple of 8
    with pytest.raises(lib.VnfrError, match="multiples of 8"):
        lib.call("vnfr_conv_prepare", C.byref(op))


def test_dropin_state_dict_surface():
    from oracle import nets, synth
    from vn_celeb_face_recognition_b200.models import MTCNN, InceptionResnetV1, MLPModel
    enc = InceptionResnetV1(pretrained=None)
    sd = nets.make_encoder_state_dict(seed=0, calibrate=False)
    assert set(enc.state_dict().keys()) == set(sd.keys()) and len(sd) == 714
    enc.load_state_dict(sd)
    enc_c = InceptionResnetV1(pretrained=None, classify=True, num_classes=10)
    assert enc_c.state_dict()["logits.weight"].shape == (10, 512)
    mlp = MLPModel(512, 1001)
    mlp.load_state_dict(nets.make_mlp_state_dict(1001))
    det = MTCNN(image_size=160, keep_all=True, min_face_size=50)             # cfg/detection/mtcnn.json minus device
    ref = synth.mtcnn_state_dicts()
    for name in ("pnet", "rnet", "onet"):
        got = getattr(det, name).state_dict()
        assert set(got.keys()) == set(ref[name].keys())
        for k in got:
            assert torch.equal(got[k], ref[name][k]), "bundled %s weights not auto-loaded" % name
    with pytest.raises(Exception, match="num_classes"):
        InceptionResnetV1(pretrained=None, classify=True)
    with pytest.raises(Exception, match="network"):
        InceptionResnetV1(pretrained="vggface2")


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA -- never compute on the CPU."""
    from vn_celeb_face_recognition_b200 import _lib
    from vn_celeb_face_recognition_b200.models import MTCNN, InceptionResnetV1, MLPModel
    with pytest.raises(_lib.VnfrError):
        InceptionResnetV1(pretrained=None).eval()(torch.zeros(1, 3, 160, 160))
    with pytest.raises(_lib.VnfrError):
        MLPModel(512, 10).eval()(torch.zeros(1, 512))
    with pytest.raises(_lib.VnfrError):
        MTCNN(min_face_size=50).detect(np.zeros((64, 64, 3), np.uint8))
    src = ""
    pkg = os.path.join(ROOT, "vn_celeb_face_recognition_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src += open(os.path.join(dp, f)).read()
    assert "import oracle" not in src and "from oracle" not in src, "the product must never import the oracle"


def test_pack_layouts_are_permutations():
    """The R/O-Net FC repack maps the reference's (W,H,C) flatten (mtcnn.py:93-94) onto the kernels' (C,H,W) order."""
    from vn_celeb_face_recognition_b200.models import mtcnn as M
    torch.manual_seed(0)
    w = torch.randn(128, 576, dtype=torch.float64)
    packed = M._fc_whc_to_chw(w, 64, 3, 3).reshape(576, 128)
    x = torch.randn(2, 64, 3, 3, dtype=torch.float64)             # conv output (N,C,H,W)
    ref = torch.nn.functional.linear(x.permute(0, 3, 2, 1).contiguous().view(2, -1), w)
    got = x.reshape(2, -1) @ packed.double()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-4)
    assert M._pack_pnet({k: v for k, v in __import__("oracle.synth", fromlist=["x"]).mtcnn_state_dicts()["pnet"].items()}).numel() == 6632


def test_stem_space_to_depth_repack_is_the_same_convolution():
    """encoder_plan.pack_stem_s2d: the stride-2 3x3 stem convolution (inception_resnet_v1.py:219) equals a stride-1 2x2
    convolution over the 16-channel space-to-depth image with the re-indexed weights (host-side packing logic, checked
    on the CPU for even and odd input sizes)."""
    from vn_celeb_face_recognition_b200 import encoder_plan as ep
    torch.manual_seed(0)
    sd = {"c.conv.weight": torch.randn(32, 3, 3, 3), "c.bn.weight": torch.rand(32) + 0.5, "c.bn.bias": torch.randn(32),
          "c.bn.running_mean": torch.randn(32), "c.bn.running_var": torch.rand(32) + 0.5}
    pc = ep.pack_stem_s2d(sd, "c", torch.device("cpu"), dtype=torch.float32)
    assert (pc.kh, pc.kw, pc.cin, pc.cout) == (2, 2, 16, 32) and tuple(pc.w.shape) == (32, 64)
    w2 = pc.w.view(32, 2, 2, 16).permute(0, 3, 1, 2)                       # [co][ch][ty][tx]
    w, s, b = ep.fold_bn(sd, "c")
    for h, wd in ((20, 24), (19, 23)):
        x = torch.randn(2, 3, h, wd)
        ref = torch.nn.functional.conv2d(x, w * s.view(-1, 1, 1, 1), b, stride=2)
        h2, w2d = (h + 1) // 2, (wd + 1) // 2
        xp = torch.zeros(2, 4, 2 * h2, 2 * w2d)
        xp[:, :3, :h, :wd] = x
        s2d = xp.view(2, 4, h2, 2, w2d, 2).permute(0, 3, 5, 1, 2, 4).reshape(2, 16, h2, w2d)   # ch = (sy*2+sx)*4 + c
        got = torch.nn.functional.conv2d(s2d, w2, pc.bias[:32])
        assert got.shape == ref.shape
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


def test_select_boxes_all_methods_match_reference_golden():
    """MTCNN.select_boxes (mtcnn.py:363-456; host numpy) for all four methods and the un-batched form, against the
    unmodified reference on its own detections (tests/golden/select_boxes_small.npz)."""
    from PIL import Image
    from conftest import load_golden
    from oracle import synth
    from vn_celeb_face_recognition_b200.models import MTCNN
    g = load_golden("select_boxes_small")
    imgs = [Image.fromarray(f) for f in synth.frames("small", 3)]
    boxes = [g["boxes_%d" % i] for i in range(3)]
    probs = [g["probs_%d" % i] for i in range(3)]
    points = [g["points_%d" % i] for i in range(3)]
    m = MTCNN(image_size=160, keep_all=False)
    picks = set()
    for method in ["probability", "largest", "center_weighted_size", "largest_over_threshold"]:
        sb, sp, spt = m.select_boxes(boxes, probs, points, imgs, method=method, threshold=float(g["threshold"]))
        for i in range(3):
            if bool(g["%s_none_%d" % (method, i)]):
                assert sb[i] is None and sp[i][0] is None and spt[i] is None
                continue
            np.testing.assert_array_equal(np.asarray(sb[i], np.float32), g["%s_box_%d" % (method, i)])
            np.testing.assert_array_equal(np.asarray(sp[i], np.float32), g["%s_prob_%d" % (method, i)])
            np.testing.assert_array_equal(np.asarray(spt[i], np.float32), g["%s_point_%d" % (method, i)])
            picks.add((method, i, tuple(np.asarray(sb[i]).ravel().tolist())))
    assert len({p[2] for p in picks}) > 3, "the methods must not all pick the same box"
    # ndarray frames (the reference's center_weighted_size needs PIL's .width; ours also accepts arrays)
    sb2, _, _ = m.select_boxes(boxes, probs, points, np.stack([np.asarray(im) for im in imgs]), method="center_weighted_size")
    for i in range(3):
        np.testing.assert_array_equal(np.asarray(sb2[i], np.float32), g["center_weighted_size_box_%d" % i])
    sb, sp, spt = m.select_boxes(boxes[0], probs[0], points[0], imgs[0], method="probability")
    np.testing.assert_array_equal(np.asarray(sb, np.float32), g["single_box"])
    assert np.float32(sp) == g["single_prob"] and np.asarray(spt).shape == (1, 5, 2)


def test_tracker_rows_byte_identical_to_reference_lines():
    """video.tracker_rows vs the text the reference's own demo_video.py lines 154-181 produce (tests/golden/tracker_rows.npz,
    oracle/make_golden_tracker.py exec's them): Time, "names", Frame_idx, "boxes / [w,h,w,h]"."""
    from conftest import load_golden
    from vn_celeb_face_recognition_b200 import video
    g = load_golden("tracker_rows")
    if str(g["numpy_version"]).split(".")[0] != np.__version__.split(".")[0]:
        pytest.skip("the list repr of numpy scalars differs between numpy major versions")
    counts, boxes, names = g["counts"].tolist(), g["boxes"], [str(n) for n in g["names"]]
    per_boxes, per_names, o = [], [], 0
    for c in counts:
        per_boxes.append([boxes[o + i] for i in range(c)])
        per_names.append(names[o:o + c])
        o += c
    info = [[float(t), i + 1] for i, t in enumerate(g["times"])]
    assert video.tracker_rows(info, per_names, per_boxes, tuple(g["frame_shape"].tolist())) == str(g["text"])


def test_frame_batcher_follows_demo_video_queue():
    """demo_video.py:78-101: batches of n_frames, a short last batch at the end of the video, time = count / fps, count
    from 1; frames land in (pinned when CUDA is present) batch buffers unchanged."""
    from vn_celeb_face_recognition_b200 import video

    class Cap:
        def __init__(self, n):
            self.frames = [np.full((4, 6, 3), i, np.uint8) for i in range(n)]
            self.i = 0

        def read(self):
            if self.i >= len(self.frames):
                return False, None
            self.i += 1
            return True, self.frames[self.i - 1]

        def get(self, prop):
            assert prop == 5
            return 25.0

    out = [(f.clone(), info) for f, info in video.FrameBatcher(Cap(5), 2)]
    assert [f.shape[0] for f, _ in out] == [2, 2, 1]
    assert [i for _, info in out for i in info] == [[(k + 1) / 25.0, k + 1] for k in range(5)]
    assert [int(f[j, 0, 0, 0]) for f, _ in out for j in range(f.shape[0])] == [0, 1, 2, 3, 4]
    assert list(video.FrameBatcher(Cap(0), 4)) == []
    assert [f.shape[0] for f, _ in video.FrameBatcher(Cap(4), 4)] == [4]


def test_heads_back_split_weights_layout():
    """models.mtcnn.HeadsBackWeights (VnfrHeadsBack, csrc/heads_chain.cu): hi + lo planes reproduce the fp32 weights to 2^-22,
    the 2x2 convolution's K order is (tap = ky*2 + kx, 64 channels, zero padded), the dense layer is torch's matrix as is
    (the reference's (W,H,C) flatten, mtcnn.py:93-94 / :150-151), the heads are stacked in output order."""
    from vn_celeb_face_recognition_b200.models import mtcnn as M
    torch.manual_seed(0)
    for onet, net in ((False, M.RNet(pretrained=False)), (True, M.ONet(pretrained=False))):
        sd = {k: torch.randn_like(v) for k, v in net.state_dict().items()}
        hb = M.HeadsBackWeights(sd, onet, "cpu")
        conv, fc = ("conv4", "dense5") if onet else ("conv3", "dense4")
        heads = ["dense6_1", "dense6_2", "dense6_3"] if onet else ["dense5_1", "dense5_2"]
        L1, L2, L3 = hb.layers
        co, ci = sd[conv + ".weight"].shape[:2]
        assert (L1.N, L1.K, L1.N_pad) == (co, 256, 128) and tuple(L1.w.shape) == (256, 256)
        w1 = (L1.w[:co].float() + L1.w[128:128 + co].float()).reshape(co, 2, 2, 64)
        ref = sd[conv + ".weight"].permute(0, 2, 3, 1)                         # co, ky, kx, c
        assert (w1[..., :ci] - ref).abs().max() <= 2.0 ** -21 * ref.abs().max() and (w1[..., ci:] == 0).all()
        assert (L1.w[co:128] == 0).all() and (L1.w[128 + co:] == 0).all()      # N padding rows
        w2 = L2.w[:L2.N].float() + L2.w[L2.N_pad:L2.N_pad + L2.N].float()
        assert tuple(w2.shape) == tuple(sd[fc + ".weight"].shape) == (L2.N, 9 * co)
        assert (w2 - sd[fc + ".weight"]).abs().max() <= 2.0 ** -21 * sd[fc + ".weight"].abs().max()
        w3 = L3.w[:L3.N].float() + L3.w[128:128 + L3.N].float()
        assert torch.allclose(w3, torch.cat([sd[h + ".weight"] for h in heads]), atol=1e-5) and L3.N == (16 if onet else 6)
        assert torch.equal(L3.bias[:L3.N], torch.cat([sd[h + ".bias"] for h in heads]))
        assert torch.equal(hb.alpha[0][:co], sd["prelu" + conv[-1] + ".weight"]) and (hb.alpha[0][co:] == 0).all()
        assert torch.equal(hb.alpha[1][:L2.N], sd["prelu" + fc[-1] + ".weight"])
        s = hb.struct(torch.zeros(8))
        assert s.w[0] == L1.w.data_ptr() and s.bias[2] == L3.bias.data_ptr() and s.alpha[1] == hb.alpha[1].data_ptr()


def test_trainer_accuracy_metric():
    """losses/metrics.py:3-7"""
    from vn_celeb_face_recognition_b200.trainer import accuracy
    out = torch.tensor([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7], [0.6, 0.4]]).log()
    assert accuracy(out, torch.tensor([1, 0, 0, 0])) == 0.75


def test_grouped_block35_packing_is_the_two_convolutions():
    """encoder_plan.pack_basic_grouped: Block35's branch1.1 / branch2.1 (3x3, 32 -> 32 each, inception_resnet_v1.py:44-51) as one
    block-diagonal 64 -> 64 convolution: unpacking the K-major weights ([cout][ky][kx][cin]) and running torch's conv2d gives the
    two branch outputs (eval-mode BN folded)."""
    from vn_celeb_face_recognition_b200 import encoder_plan as ep
    g = torch.Generator().manual_seed(5)
    sd = {}
    for p in ("a", "b"):
        sd[p + ".conv.weight"] = torch.randn(32, 32, 3, 3, generator=g) * 0.1
        sd[p + ".bn.weight"] = torch.rand(32, generator=g) + 0.5
        sd[p + ".bn.bias"] = torch.randn(32, generator=g) * 0.1
        sd[p + ".bn.running_mean"] = torch.randn(32, generator=g) * 0.1
        sd[p + ".bn.running_var"] = torch.rand(32, generator=g) + 0.5
    pc = ep.pack_basic_grouped(sd, ["a", "b"], "cpu", dtype=torch.float32)
    assert (pc.cout, pc.cin, pc.kh, pc.kw) == (64, 64, 3, 3) and tuple(pc.w.shape) == (64, 576)
    w = pc.w.reshape(64, 3, 3, 64).permute(0, 3, 1, 2)                       # -> [cout][cin][ky][kx]
    x = torch.randn(2, 64, 9, 9, generator=g)
    got = torch.nn.functional.conv2d(x, w, pc.bias[:64], padding=1)
    for i, p in enumerate(("a", "b")):
        wf, s, b = ep.fold_bn(sd, p)
        ref = torch.nn.functional.conv2d(x[:, 32 * i:32 * i + 32], wf * s.view(-1, 1, 1, 1), b, padding=1)
        assert torch.allclose(got[:, 32 * i:32 * i + 32], ref, atol=1e-5)
    assert (w[:32, 32:] == 0).all() and (w[32:, :32] == 0).all()             # off-diagonal blocks


def test_split_linear_planes_and_split_k_choice():
    """tail.SplitLinear: rows [0, N_pad) = fp16(w), rows [N_pad, 2 N_pad) = fp16(w - hi), K padded to 64 / N to 128 with zeros;
    tail.pick_split_k covers the SMs with (m tiles x n tiles x splits) without leaving a split fewer than two K blocks."""
    from vn_celeb_face_recognition_b200 import tail
    g = torch.Generator().manual_seed(6)
    w, b = torch.randn(1001, 500, generator=g), torch.randn(1001, generator=g)
    L = tail.SplitLinear(w, b, "cpu")
    assert (L.N, L.K, L.K_real, L.N_pad) == (1001, 512, 500, 1024) and tuple(L.w.shape) == (2048, 512)
    rec = L.w[:1001, :500].float() + L.w[1024:2025, :500].float()
    assert (rec - w).abs().max() <= 2.0 ** -21 * w.abs().max()
    assert (L.w[:, 500:] == 0).all() and (L.w[1001:1024] == 0).all() and (L.w[2025:] == 0).all()
    assert torch.equal(L.bias[:1001], b) and (L.bias[1001:] == 0).all()
    for m_tiles, n_tiles, kb in ((6, 4, 28), (6, 16, 8), (1, 8, 32), (32, 8, 8), (1, 1, 3)):
        s = tail.pick_split_k(m_tiles, n_tiles, kb)
        assert 1 <= s <= 8 and (s == 1 or kb // s >= 2) and m_tiles * n_tiles * s <= max(148, m_tiles * n_tiles)
