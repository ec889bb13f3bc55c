"""GPU parity: tcgen05 implicit-GEMM convolution, encoder companions, full InceptionResnetV1 + MLP vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_encoder_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "needs a CUDA device"
    return torch.device("cuda:0")


def _conv_case(dev, n, h, w, cin, cout, k, stride, pad, cin_real=None, residual=False, relu=True, split=None, f32=False,
               block_n=None, seed=0, dt=torch.bfloat16, sv=None, expect_mode=None):
    from vn_celeb_face_recognition_b200 import encoder_plan as ep
    g = torch.Generator(device="cpu").manual_seed(seed)
    kh, kw = (k, k) if isinstance(k, int) else k
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    cin_real = cin_real or cin
    x = torch.randn(n, h, w, cin, generator=g)
    x[..., cin_real:] = 0
    wt = torch.randn(cout, cin_real, kh, kw, generator=g) / (cin_real * kh * kw) ** 0.5
    bias = torch.randn(cout, generator=g)
    xb = x.to(dev).to(dt)
    pc = ep.pack_conv(wt, None, bias, dev, cin_pad=max(cin, sv or 0), block_n=block_n, dtype=dt)
    oh, ow = (h + 2 * ph - kh) // stride + 1, (w + 2 * pw - kw) // stride + 1
    ref = torch.nn.functional.conv2d(xb.float().permute(0, 3, 1, 2)[:, :cin_real], wt.to(dev).to(dt).float(),
                                     bias.to(dev), stride=stride, padding=(ph, pw))
    res = None
    if residual:
        res = torch.randn(n, oh, ow, cout, generator=g).to(dev).to(dt)
        ref = ref + res.float().permute(0, 3, 1, 2)
    if relu:
        ref = ref.relu()
    ref = ref.permute(0, 2, 3, 1).contiguous()
    ol = ep.OpList()
    if f32:
        out = torch.full((n * oh * ow, cout), float("nan"), dtype=torch.float32, device=dev)
        ol.conv(pc, ep.View(xb), None, stride=stride, pad=(ph, pw), relu=relu, out_f32=out)
        ol.run()
        torch.cuda.synchronize()
        got = out.view(n, oh, ow, cout)
        tol = 2e-3
    elif split:
        wide = torch.full((n, oh, ow, split + 24), float("nan"), dtype=dt, device=dev)   # slice of a wider buffer
        o1 = torch.full((n, oh, ow, cout - split), float("nan"), dtype=dt, device=dev)
        ol.conv(pc, ep.View(xb), ep.View(wide, 8, split), stride=stride, pad=(ph, pw), relu=relu, dst1=ep.View(o1),
                n_split=split, residual=None if res is None else ep.View(res))
        ol.run()
        torch.cuda.synchronize()
        assert torch.isnan(wide[..., :8].float()).all() and torch.isnan(wide[..., 8 + split:].float()).all(), "wrote outside slice"
        got = torch.cat([wide[..., 8:8 + split], o1], dim=-1).float()
        tol = 2e-2
    else:
        out = torch.full((n, oh, ow, cout), float("nan"), dtype=dt, device=dev)
        ol.conv(pc, ep.View(xb), ep.View(out), stride=stride, pad=(ph, pw), relu=relu,
                residual=None if res is None else ep.View(res), sv=sv)
        ol.run()
        torch.cuda.synchronize()
        got = out.float()
        tol = 2e-2
    if expect_mode is not None:
        assert ol.ops[-1].conv.a_mode == expect_mode, "kernel variant %d ran, expected %d" % (ol.ops[-1].conv.a_mode, expect_mode)
    assert torch.isfinite(got).all(), "non-finite / unwritten outputs"
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * max(scale, 1.0), "max abs err %g (ref max %g)" % (err, scale)


@pytest.mark.parametrize("case", [
    dict(n=2, h=9, w=9, cin=64, cout=64, k=1, stride=1, pad=0),                       # plain GEMM, K = 64
    dict(n=3, h=8, w=8, cin=896, cout=256, k=1, stride=1, pad=0, block_n=128, split=128),   # Block17 fused 1x1, 2 N tiles
    dict(n=2, h=17, w=17, cin=256, cout=96, k=1, stride=1, pad=0, split=32),          # Block35 fused 1x1, in-tile split
    dict(n=2, h=17, w=17, cin=32, cout=32, k=3, stride=1, pad=1),                     # Block35 3x3, two taps per K block
    dict(n=2, h=15, w=15, cin=8, cin_real=3, cout=32, k=3, stride=2, pad=0),          # stem conv2d_1a (padded channels)
    dict(n=1, h=12, w=12, cin=80, cout=192, k=3, stride=1, pad=0),                    # conv2d_4a: K = 720 (tail block)
    dict(n=2, h=8, w=8, cin=128, cout=128, k=(1, 7), stride=1, pad=(0, 3)),
    dict(n=2, h=8, w=8, cin=128, cout=128, k=(7, 1), stride=1, pad=(3, 0)),
    dict(n=2, h=17, w=17, cin=256, cout=384, k=3, stride=2, pad=0),                   # Mixed_6a branch0, 2 N tiles of 192
    dict(n=5, h=3, w=3, cin=384, cout=1792, k=1, stride=1, pad=0, residual=True),     # Block8 projection + residual + ReLU
    dict(n=5, h=3, w=3, cin=384, cout=1792, k=1, stride=1, pad=0, residual=True, relu=False),
    dict(n=300, h=1, w=1, cin=1792, cout=512, k=1, stride=1, pad=0, relu=False, f32=True),   # last_linear, fp32 out
    dict(n=130, h=1, w=1, cin=2048, cout=1008, k=1, stride=1, pad=0, relu=False, f32=True),  # MLP dense_2 (block_n 144)
    dict(n=4, h=40, w=40, cin=64, cout=80, k=1, stride=1, pad=0),                     # many M tiles, N = 80
])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_igemm_conv_matches_torch(dev, case, dt):
    _conv_case(dev, dt=dt, **case)


@pytest.mark.parametrize("case", [
    # shifted-view kernel (csrc/sv_conv.cu, a_mode 3): resident weights, 64-byte swizzle rows (32 channels per plane)
    dict(n=3, h=40, w=40, cin=32, cout=32, k=3, stride=1, pad=0, sv=32),              # conv2d_2a-like: several bands per image
    dict(n=2, h=37, w=37, cin=32, cout=64, k=3, stride=1, pad=1, sv=32),              # conv2d_2b-like: TMA zero-fill = padding
    dict(n=9, h=17, w=17, cin=32, cout=32, k=3, stride=1, pad=1, sv=32),              # Block35 3x3: several images per band
    dict(n=2, h=17, w=17, cin=32, cout=32, k=3, stride=1, pad=1, sv=32, residual=True),
    # 128-byte swizzle rows; 32 real channels zero-extended to a 64-channel plane by the TMA unit
    dict(n=3, h=40, w=40, cin=32, cout=32, k=3, stride=1, pad=0, sv=64),
    dict(n=9, h=17, w=17, cin=32, cout=32, k=3, stride=1, pad=1, sv=64),
    dict(n=2, h=21, w=21, cin=64, cout=64, k=3, stride=1, pad=1, sv=64),
    dict(n=3, h=40, w=40, cin=16, cout=32, k=2, stride=1, pad=0, sv=16),              # s2d stem: 32-byte swizzle rows
    # streamed weights, two accumulator tiles per weight stage
    dict(n=7, h=8, w=8, cin=128, cout=128, k=(1, 7), stride=1, pad=(0, 3), sv=64),
    dict(n=7, h=8, w=8, cin=128, cout=128, k=(7, 1), stride=1, pad=(3, 0), sv=64),
    dict(n=150, h=8, w=8, cin=128, cout=128, k=(1, 7), stride=1, pad=(0, 3), sv=64),    # more bands than SMs
])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_shifted_view_conv_matches_torch(dev, case, dt):
    _conv_case(dev, dt=dt, expect_mode=3, **case)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_pool_norm_softmax_kernels(dev, dt):
    from vn_celeb_face_recognition_b200 import _lib
    code = 1 if dt == torch.float16 else 0
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 17, 17, 264, generator=g).to(dev).to(dt)
    out = torch.zeros(3, 8, 8, 64 + 256, dtype=dt, device=dev)
    _lib.call("vnfr_maxpool3s2_nhwc", _lib.ptr(x), 3, 17, 17, 256, 264, _lib.ptr(out[..., 64:]), 320, code, _lib.stream_ptr())
    ref = torch.nn.functional.max_pool2d(x[..., :256].float().permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert torch.equal(out[..., 64:].float(), ref) and (out[..., :64] == 0).all()

    y = torch.randn(5, 3, 3, 1792, generator=g).to(dev).to(dt)
    pooled = torch.empty(5, 1792, dtype=dt, device=dev)
    _lib.call("vnfr_avgpool_nhwc", _lib.ptr(y), 5, 9, 1792, 1792, _lib.ptr(pooled), code, _lib.stream_ptr())
    torch.testing.assert_close(pooled.float(), y.float().mean(dim=(1, 2)), atol=1e-2, rtol=1e-2)

    z = torch.randn(7, 3, 20, 24, generator=g).to(dev)
    nhwc = torch.empty(7, 20, 24, 8, dtype=dt, device=dev)
    _lib.call("vnfr_nchw3_to_nhwc8", _lib.ptr(z), 7, 20, 24, _lib.ptr(nhwc), code, _lib.stream_ptr())
    assert torch.equal(nhwc[..., :3], z.permute(0, 2, 3, 1).to(dt)) and (nhwc[..., 3:] == 0).all()
    for hh, ww in ((20, 24), (19, 23)):                    # even and odd sizes (odd: zero pad row / column)
        z2 = torch.randn(7, 3, hh, ww, generator=g).to(dev)
        s2d = torch.full((7, (hh + 1) // 2, (ww + 1) // 2, 16), float("nan"), dtype=dt, device=dev)
        _lib.call("vnfr_nchw3_to_s2d16", _lib.ptr(z2), 7, hh, ww, _lib.ptr(s2d), code, _lib.stream_ptr())
        zp = torch.zeros(7, 4, 2 * ((hh + 1) // 2), 2 * ((ww + 1) // 2), device=dev)
        zp[:, :3, :hh, :ww] = z2
        exp = zp.view(7, 4, (hh + 1) // 2, 2, (ww + 1) // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(7, (hh + 1) // 2, (ww + 1) // 2, 16)
        assert torch.equal(s2d, exp.to(dt))

    e = torch.randn(9, 512, generator=g).to(dev)
    emb = torch.empty_like(e)
    emb16 = torch.empty(9, 512, dtype=dt, device=dev)
    _lib.call("vnfr_l2norm_rows", _lib.ptr(e), 9, 512, 512, _lib.ptr(emb), _lib.ptr(emb16), code, _lib.stream_ptr())
    torch.testing.assert_close(emb, torch.nn.functional.normalize(e, p=2, dim=1), atol=1e-6, rtol=1e-5)
    assert torch.equal(emb16, emb.to(dt))

    lg = torch.randn(11, 1008, generator=g).to(dev)
    logp = torch.empty(11, 1001, device=dev)
    lab = torch.empty(11, dtype=torch.int64, device=dev)
    pr = torch.empty(11, device=dev)
    _lib.call("vnfr_logsoftmax_argmax", _lib.ptr(lg), 11, 1001, 1008, _lib.ptr(logp), _lib.ptr(lab), _lib.ptr(pr), _lib.stream_ptr())
    ref = torch.log_softmax(lg[:, :1001], dim=1)
    torch.testing.assert_close(logp, ref, atol=1e-5, rtol=1e-5)
    assert torch.equal(lab, ref.argmax(1))
    torch.testing.assert_close(pr, ref.max(1)[0].exp(), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("dt,min_cos,margin_thr", [(torch.float16, 0.999, 1e-3), (torch.bfloat16, 0.995, 0.15)])
def test_encoder_and_mlp_match_oracle_and_golden(dev, dt, min_cos, margin_thr):
    """North-star tolerances: embedding cosine >= 0.999 and identical labels -- met by the default fp16 storage type
    (fp32 accumulation) with the tail (bottleneck, L2 norm, MLP) at fp32-level accuracy.  The optional bf16 storage type is
    covered at the accuracy it actually delivers (measured 0.9977-0.9997 over 130 layers), which is why it is not the
    default.  A label may differ from the fp32 oracle's only where the oracle's own top-2 log-prob margin is below
    ``margin_thr`` (fp16: 1e-3); every such face is printed."""
    from oracle import nets, synth
    from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
    sd = golden_encoder_state_dict()
    mlp_sd = nets.make_mlp_state_dict(1001, seed=0)
    enc = InceptionResnetV1(pretrained=None, device=dev).eval()
    enc.half_dtype = dt
    enc.load_state_dict(sd)
    mlp = MLPModel(512, 1001).to(dev).eval()
    mlp.load_state_dict(mlp_sd)
    x = synth.crops_160(24, seed=1)
    with torch.no_grad():
        taps = {}
        e_ref = nets.encoder_forward(sd, x, taps=taps)
        lp_ref = nets.mlp_forward(mlp_sd, e_ref)
        e = enc(x.to(dev))
        lp = mlp(e)
    torch.cuda.synchronize()
    # per-stage taps first: a failure names the first broken stage
    plan = enc._plans[(enc._bucket(24), 160, 160)]             # plans are bucketed: the first 24 rows are ours
    for name, key in [("conv2d_1a", "conv2d_1a"), ("conv2d_2b", "conv2d_2b"), ("conv2d_4b_repeat_1", "repeat_1.4"),
                      ("repeat_2", "repeat_2.9"), ("block8", "block8")]:
        got = plan.taps[name][:24].float().permute(0, 3, 1, 2).cpu()
        ref = taps[key]
        rel = (got - ref).norm() / ref.norm()
        print("stage %-12s relative error %.4f" % (key, rel))
        assert rel < 0.08, "stage %s relative error %.4f" % (key, rel)      # debugging aid; the real bar is the cosine below
    e = e.cpu()
    cos = torch.nn.functional.cosine_similarity(e, e_ref, dim=1)
    print("embedding cosine min %.6f mean %.6f" % (cos.min(), cos.mean()))
    assert cos.min().item() >= min_cos, "embedding cosine %s" % cos
    assert torch.allclose(e.norm(dim=1), torch.ones(24), atol=1e-5)
    g = load_golden("encoder_seed0")
    cos_g = torch.nn.functional.cosine_similarity(e[:8], torch.from_numpy(g["emb"]), dim=1)
    assert cos_g.min().item() >= min_cos
    top2 = lp_ref.topk(2, dim=1)[0]
    margin = top2[:, 0] - top2[:, 1]
    lab, lab_ref = lp.argmax(1).cpu(), lp_ref.argmax(1)
    sure = margin > margin_thr
    for k in (lab != lab_ref).nonzero().flatten().tolist():
        print("face %d: label %d, oracle %d, oracle margin %.3e" % (k, lab[k], lab_ref[k], margin[k]))
    assert sure.sum() >= 20 or dt != torch.float16
    assert torch.equal(lab[sure], lab_ref[sure]), "labels differ where the oracle margin is > %g" % margin_thr
    print("label agreement %d/24 (margin>%g: %d), min cosine %.5f" % ((lab == lab_ref).sum(), margin_thr, sure.sum(), cos.min()))
    # MLP alone on identical inputs: fp32-level accuracy (split-precision contractions), identical labels
    with torch.no_grad():
        lp_same = mlp(e_ref.to(dev)).cpu()
    assert (lp_same - lp_ref).abs().max().item() < 2e-5
    assert torch.equal(lp_same.argmax(1), lab_ref)
    # 112x112 crops (demo_video default target size) also run
    with torch.no_grad():
        x112 = torch.nn.functional.interpolate(x[:2], size=(112, 112), mode="bilinear", align_corners=False)
        e112 = enc(x112.to(dev)).cpu()
    cos112 = torch.nn.functional.cosine_similarity(e112, torch.from_numpy(g["emb112"]), dim=1)
    assert cos112.min().item() >= min_cos


def test_gallery_topk_matches_torch(dev):
    """Cosine top-5 against a gallery shard (config 5): the fused score-GEMM + top-k kernel (scores never leave TMEM; 12 gallery
    tiles split over CTAs and merged, then a single-split run over 40 query tiles... see the loop below) vs
    torch.topk(E @ G^T) in fp32.  Operands are 16-bit, so near-ties may swap: every returned row must be (re-scored in
    fp32) within 2e-3 of the true k-th best, values must match the fp32 scores of the returned rows, and well-separated
    matches (planted duplicates) must be found exactly with global indices."""
    from vn_celeb_face_recognition_b200 import gallery
    g = torch.Generator(device="cpu").manual_seed(0)
    G = torch.nn.functional.normalize(torch.randn(3000, 512, generator=g), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(70, 512, generator=g), dim=1)
    Q[:10] = torch.nn.functional.normalize(G[100:110] + 0.05 * torch.randn(10, 512, generator=g), dim=1)   # planted matches
    shard = gallery.GalleryShard(G.to(dev), index_offset=5000)
    vals, idx = shard.topk(Q.to(dev), k=5)                         # 1 query tile -> 12 gallery splits, merged
    v1, i1 = shard.topk(Q.to(dev), k=5, sms=1)                     # the same without splitting: one CTA walks all 12 tiles
    assert torch.equal(idx, i1) and torch.equal(vals, v1)
    torch.cuda.synchronize()
    ref = Q @ G.t()
    rv, ri = torch.topk(ref, 5, dim=1)
    idx, vals = idx.cpu(), vals.cpu()
    assert idx.shape == (70, 5) and (idx >= 5000).all() and (idx < 8000).all()
    assert (idx[:10, 0] == torch.arange(100, 110) + 5000).all()
    rescored = torch.gather(ref, 1, idx - 5000)
    assert (rescored >= rv[:, 4:5] - 2e-3).all()
    assert (vals - rescored).abs().max() < 2e-3
    assert (vals[:, :-1] >= vals[:, 1:]).all()
    for r in range(70):
        assert len(set(idx[r].tolist())) == 5
    # fewer gallery rows than k: missing entries are -1 / -inf
    small = gallery.GalleryShard(G[:3].to(dev))
    v2, i2 = small.topk(Q[:4].to(dev), k=5)
    assert (i2[:, 3:] == -1).all() and torch.isinf(v2[:, 3:]).all() and (i2[:, :3] >= 0).all()
    # many query tiles (persistent CTAs walk several items), ragged last tile, exact ties -> lower index
    g2 = torch.Generator(device="cpu").manual_seed(1)
    Gb = torch.nn.functional.normalize(torch.randn(700, 512, generator=g2), dim=1)
    Gb[650] = Gb[13]                                               # duplicate rows: identical scores
    Qb = torch.nn.functional.normalize(torch.randn(20000, 512, generator=g2), dim=1)
    Qb[5] = Gb[13]
    big = gallery.GalleryShard(Gb.to(dev))
    vb, ib = big.topk(Qb.to(dev), k=5)
    torch.cuda.synchronize()
    refb = Qb.half().float() @ Gb.half().float().t()
    rvb, rib = torch.topk(refb, 5, dim=1)
    assert ib[5, 0].item() == 13 and ib[5, 1].item() == 650
    resc = torch.gather(refb, 1, ib.cpu())
    assert (vb.cpu() - resc).abs().max() < 1e-3 and (resc >= rvb[:, 4:5] - 1e-3).all()
    assert (ib.cpu() == rib).float().mean() > 0.995


def test_gallery_topk_two_cta_kernel_matches(dev):
    """The tcgen05.mma.cta_group::2 variant of the gallery search (VNFR_GALLERY_2CTA=1: CTA pairs, 256 queries per item, each CTA
    loads half of every gallery tile) returns exactly what the one-CTA kernel returns (same products, same accumulation order)."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, ".")
        from vn_celeb_face_recognition_b200 import gallery
        g = torch.Generator().manual_seed(4)
        G = torch.nn.functional.normalize(torch.randn(5000, 512, generator=g), dim=1).cuda()
        Q = torch.nn.functional.normalize(torch.randn(1000, 512, generator=g), dim=1).cuda()
        v, i = gallery.GalleryShard(G).topk(Q, k=5)
        v2, i2 = gallery.GalleryShard(G).topk(Q[:300], k=5, sms=1000)       # several splits x two query pairs (ragged)
        torch.save((v.cpu(), i.cpu(), v2.cpu(), i2.cpu()), sys.argv[1])
    """)
    import os, tempfile
    outs = []
    for two in (False, True):
        env = dict(os.environ)
        env.pop("VNFR_GALLERY_2CTA", None)
        if two:
            env["VNFR_GALLERY_2CTA"] = "1"
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            outs.append(torch.load(f.name))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode", [1, 2])
def test_shifted_view_split_precision_conv(dev, mode):
    """Split-precision modes of the shifted-view kernel: fp32 activations and weights as 3 bf16 parts each (6 tensor-core
    products per K slice) or 2 fp16 parts each (3 products), fp32 accumulation -> fp32-level accuracy (the mode the
    detector heads need: their thresholded decisions must match the fp32 reference).  Also covers PReLU, fp32 output and
    the device-side image count."""
    from vn_celeb_face_recognition_b200 import encoder_plan as ep
    g = torch.Generator(device="cpu").manual_seed(3)
    n, h, w, cin, cout = 5, 23, 23, 32, 64
    x = torch.randn(n, h, w, cin, generator=g)
    x[0, :4] *= 1e-3                                                                   # small activations: fp16 subnormal residuals
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g)
    alpha = torch.rand(cout, generator=g)
    if mode == 1:
        xs = torch.cat(ep.split3_bf16(x), dim=-1).to(dev).contiguous()                 # (n, h, w, 96): hi | mid | lo
        pc = ep.pack_conv_split3(wt, bias, dev, ck=32)
    else:
        xs = torch.cat(ep.split2_fp16(x), dim=-1).to(dev).contiguous()                 # (n, h, w, 64): hi | lo
        pc = ep.pack_conv_split2(wt, bias, dev, ck=32)
    out = torch.full((n * 21 * 21, cout), float("nan"), dtype=torch.float32, device=dev)
    live = torch.tensor([4], dtype=torch.int32, device=dev)                            # only 4 of the 5 images are valid
    ol = ep.OpList()
    ol.conv(pc, ep.View(xs), None, relu=False, out_f32=out, sv=32, alpha=alpha.to(dev), n_img_dev=live)
    assert ol.ops[-1].conv.a_mode == 3
    ol.run()
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), wt.double(), bias.double())
    ref = torch.where(ref > 0, ref, ref * alpha.double().view(1, -1, 1, 1)).permute(0, 2, 3, 1).reshape(n * 21 * 21, cout)
    got = out.cpu().double()
    assert torch.isnan(got[4 * 441:]).all(), "images beyond the device-side count must not be written"
    err = (got[:4 * 441] - ref[:4 * 441]).abs().max().item()
    # an fp32 FMA convolution of the same operands for scale: the split modes must be in the same error class
    ref32 = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), wt, bias)
    ref32 = torch.where(ref32 > 0, ref32, ref32 * alpha.view(1, -1, 1, 1)).permute(0, 2, 3, 1).reshape(n * 21 * 21, cout)
    err32 = (ref32[:4 * 441].double() - ref[:4 * 441]).abs().max().item()
    print("split mode %d: max err %.3e (fp32 conv2d: %.3e)" % (mode, err, err32))
    assert err < 3e-6 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("n", [1, 5, 301])
def test_fused_block17_matches_unfused_and_torch(dev, n):
    """csrc/block17_fused.cu (1x1 -> 1x7 -> 7x1 -> projection + residual + ReLU in one persistent kernel, in place) against
    the four-launch form of the same packed weights and against torch fp32 (inception_resnet_v1.py:70-95).  n = 1 / 5: odd
    image counts (half-empty last tile); n = 301: 151 tiles on 148 CTAs (persistent loop, barrier phases across tiles)."""
    from oracle import nets
    from vn_celeb_face_recognition_b200 import encoder_plan as ep
    sd = {k: v.to(dev) for k, v in golden_encoder_state_dict().items() if k.startswith("repeat_2.3.")}
    p = "repeat_2.3"
    dt = torch.float16
    P_in = ep.pack_basic(sd, [p + ".branch0", p + ".branch1.0"], dev, block_n=256, dtype=dt)
    P_a = ep.pack_basic(sd, [p + ".branch1.1"], dev, dtype=dt)
    P_b = ep.pack_basic(sd, [p + ".branch1.2"], dev, dtype=dt)
    P_out = ep.pack_projection(sd, p + ".conv2d", 0.10, dev, dtype=dt)
    g = torch.Generator().manual_seed(n)
    x0 = torch.relu(torch.randn(n, 8, 8, 896, generator=g)).to(dev).to(dt)
    # fused, in place
    xf = x0.clone()
    ol = ep.OpList()
    ol.block17(P_in, P_a, P_b, P_out, xf)
    ol.run()
    # the four-launch form
    xu = x0.clone()
    cat, ta, tb = (torch.empty(n, 8, 8, c, dtype=dt, device=dev) for c in (256, 128, 128))
    ol2 = ep.OpList()
    ol2.conv(P_in, ep.View(xu), ep.View(cat, 0, 128), dst1=ep.View(ta), n_split=128)
    ol2.conv(P_a, ep.View(ta), ep.View(tb), pad=(0, 3))
    ol2.conv(P_b, ep.View(tb), ep.View(cat, 128, 128), pad=(3, 0))
    ol2.conv(P_out, ep.View(cat), ep.View(xu), residual=ep.View(xu), relu=True)
    ol2.run()
    torch.cuda.synchronize()
    ref = nets._block17({k: v.float().cpu() for k, v in sd.items()}, p, x0.float().cpu().permute(0, 3, 1, 2), 0.10).permute(0, 2, 3, 1)
    got, unf = xf.float().cpu(), xu.float().cpu()
    scale = ref.abs().max().item()
    e_f, e_u, e_fu = (got - ref).abs().max().item(), (unf - ref).abs().max().item(), (got - unf).abs().max().item()
    print("n=%d: fused vs torch %.4f, unfused vs torch %.4f, fused vs unfused %.4f (scale %.2f)" % (n, e_f, e_u, e_fu, scale))
    assert e_f < 0.02 * scale and e_fu < 0.02 * scale
    rel = (got - ref).norm() / ref.norm()
    assert rel < 3e-3, rel
