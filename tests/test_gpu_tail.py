"""The fused tail kernel (csrc/tail_fused.cu: pool -> bottleneck -> L2-normalise -> MLP -> log-softmax -> argmax in one
cooperative launch, split-precision tensor-core contractions) against fp64 / fp32 torch on the same operands.  The bar is
fp32-level accuracy: the predicted label must be the fp32 reference's (inception_resnet_v1.py:294-302, mlp_model.py:10-15,
demo_image.py:113-137)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "needs a CUDA device"
    return torch.device("cuda:0")


def _ref_chain(x8, layers, dtype=torch.float64):
    """x8 (n, hw, C) -> per-layer row-op outputs in ``dtype`` on the CPU."""
    v = x8.to(dtype).mean(dim=1)
    outs = []
    for w, b, op in layers:
        v = v @ w.to(dtype).t() + (0 if b is None else b.to(dtype))
        if op == "relu":
            v = torch.relu(v)
        elif op == "l2norm":
            v = torch.nn.functional.normalize(v, p=2, dim=1)
        elif op == "logsoftmax":
            v = torch.log_softmax(v, dim=1)
        outs.append(v)
    return outs


@pytest.mark.parametrize("n,dt", [(1, torch.float16), (6, torch.float16), (130, torch.bfloat16), (768, torch.float16)])
def test_tail_full_chain_matches_fp64(dev, n, dt):
    from vn_celeb_face_recognition_b200 import tail
    g = torch.Generator().manual_seed(n)
    x8 = torch.relu(torch.randn(n, 9, 1792, generator=g)).to(dt)                      # block8 output (post-ReLU scale)
    w1 = torch.randn(512, 1792, generator=g) / 1792 ** 0.5
    b1 = 0.1 * torch.randn(512, generator=g)
    w2 = (torch.rand(2048, 512, generator=g) * 2 - 1) / 512 ** 0.5                     # nn.Linear default init
    b2 = (torch.rand(2048, generator=g) * 2 - 1) / 512 ** 0.5
    w3 = (torch.rand(1001, 2048, generator=g) * 2 - 1) / 2048 ** 0.5
    b3 = (torch.rand(1001, generator=g) * 2 - 1) / 2048 ** 0.5
    layers = [(tail.SplitLinear(w1, b1, dev), "l2norm"), (tail.SplitLinear(w2, b2, dev), "relu"),
              (tail.SplitLinear(w3, b3, dev), "logsoftmax")]
    plan = tail.TailPlan(layers, n, dev, in_mode=0)
    cap = n + 5
    payload = torch.full((cap + 1, 514), float("nan"), device=dev)
    logp = torch.full((n, 1001), float("nan"), device=dev)
    label = torch.full((n,), -1, dtype=torch.int64, device=dev)
    prob = torch.full((n,), float("nan"), device=dev)
    emb16 = torch.zeros(n, 512, dtype=dt, device=dev)
    for rep in range(2):                                                              # second run: barrier word / pipeline state reset
        plan.run(n, x=x8.to(dev), out_vecs=[payload[:n, :512], None, logp], emb_half=emb16, label=label, prob=prob,
                 payload=payload[:n], n_classes=1001, count_cell=payload[-1, :1], count_value=n)
        torch.cuda.synchronize()
    ref = _ref_chain(x8.float(), [(w1, b1, "l2norm"), (w2, b2, "relu"), (w3, b3, "logsoftmax")])
    ref32 = _ref_chain(x8.float(), [(w1, b1, "l2norm"), (w2, b2, "relu"), (w3, b3, "logsoftmax")], torch.float32)
    emb = payload[:n, :512].cpu().double()
    err_e, err_e32 = (emb - ref[0]).abs().max().item(), (ref32[0].double() - ref[0]).abs().max().item()
    err_l, err_l32 = (logp.cpu().double() - ref[2]).abs().max().item(), (ref32[2].double() - ref[2]).abs().max().item()
    print("n=%d: emb err %.2e (torch fp32 %.2e), logp err %.2e (torch fp32 %.2e)" % (n, err_e, err_e32, err_l, err_l32))
    assert err_e < 2e-6 and err_l < 2e-5, (err_e, err_l)
    # labels: identical wherever the fp64 margin exceeds the fp32-level error bar
    top2 = ref[2].topk(2, dim=1)[0]
    sure = (top2[:, 0] - top2[:, 1]) > 1e-4
    lab = label.cpu()
    assert torch.equal(lab[sure], ref[2].argmax(1)[sure])
    assert sure.float().mean() > 0.9
    torch.testing.assert_close(prob.cpu().double(), ref[2].max(1)[0].exp(), atol=1e-6, rtol=1e-4)
    # send-buffer columns and count cell
    assert torch.equal(payload[:n, 512].cpu().long(), lab) and torch.equal(payload[:n, 513], prob)
    assert payload[-1, 0].item() == float(n) and torch.isnan(payload[n:cap]).all()
    assert torch.equal(emb16.float().cpu(), payload[:n, :512].to(dt).float().cpu())


def test_tail_threshold_and_per_class_threshold(dev):
    """identify_person's threshold (demo_image.py:113-137): scalar and per-class dict -> label or n_classes."""
    from vn_celeb_face_recognition_b200 import tail
    g = torch.Generator().manual_seed(5)
    n, C = 40, 37
    x = torch.randn(n, 100, generator=g)                                              # K = 100: padded to 128 inside
    w2, b2 = torch.randn(2048, 100, generator=g) * 0.1, torch.randn(2048, generator=g) * 0.1
    w3, b3 = torch.randn(C, 2048, generator=g) * 0.2, torch.randn(C, generator=g)
    layers = [(tail.SplitLinear(w2, b2, dev), "relu"), (tail.SplitLinear(w3, b3, dev), "logsoftmax")]
    plan = tail.TailPlan(layers, n, dev, in_mode=1)
    lp_ref = torch.log_softmax(torch.relu(x.double() @ w2.double().t() + b2.double()) @ w3.double().t() + b3.double(), 1)
    pred, pmax = lp_ref.argmax(1), lp_ref.max(1)[0].exp()
    label = torch.empty(n, dtype=torch.int64, device=dev)
    prob = torch.empty(n, device=dev)
    logp = torch.empty(n, C, device=dev)
    xd = x.to(dev)
    plan.run(n, x_f32=xd, out_vecs=[None, logp], label=label, prob=prob, n_classes=C)
    assert torch.equal(label.cpu(), pred)
    assert (logp.cpu().double() - lp_ref).abs().max() < 3e-6 * lp_ref.abs().max()      # logits of magnitude ~40 here
    thr = float(pmax.median())
    plan.run(n, x_f32=xd, label=label, prob=prob, thr=thr, n_classes=C)
    exp = torch.where(pmax >= thr, pred, torch.full_like(pred, C))
    border = (pmax - thr).abs() < 1e-6
    assert torch.equal(label.cpu()[~border], exp[~border]) and (label == C).any() and (label < C).any()
    per_class = torch.rand(C, generator=g)
    plan.run(n, x_f32=xd, label=label, prob=prob, thr_class=per_class.to(dev), n_classes=C)
    exp = torch.where(pmax >= per_class[pred].double(), pred, torch.full_like(pred, C))
    border = (pmax - per_class[pred].double()).abs() < 1e-6
    assert torch.equal(label.cpu()[~border], exp[~border])


def test_mlp_and_classify_encoder_heads(dev):
    """MLPModel.forward on fp32 embeddings and InceptionResnetV1(classify=True) (inception_resnet_v1.py:298-300) through
    the fused tail vs the oracle."""
    from oracle import nets
    from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
    mlp_sd = nets.make_mlp_state_dict(1001, seed=0)
    mlp = MLPModel(512, 1001).to(dev).eval()
    mlp.load_state_dict(mlp_sd)
    g = torch.Generator().manual_seed(1)
    e = torch.nn.functional.normalize(torch.randn(300, 512, generator=g), dim=1)
    with torch.no_grad():
        lp = mlp(e.to(dev)).cpu()
        lp_ref = nets.mlp_forward(mlp_sd, e)
    assert lp.shape == (300, 1001)
    assert (lp - lp_ref).abs().max().item() < 2e-5
    assert torch.equal(lp.argmax(1), lp_ref.argmax(1))
    # classify=True: logits Linear(512, C) + log_softmax on the UN-normalised bottleneck output
    torch.manual_seed(3)
    sd = nets.make_encoder_state_dict(seed=0)
    sd["logits.weight"] = torch.randn(10, 512) * 0.05
    sd["logits.bias"] = torch.randn(10) * 0.1
    enc = InceptionResnetV1(pretrained=None, classify=True, num_classes=10, device=dev).eval()
    enc.load_state_dict(sd)
    from oracle import synth
    x = synth.crops_160(4, seed=2)
    with torch.no_grad():
        out = enc(x.to(dev)).cpu()
        ref = nets.encoder_forward(sd, x, classify=True)
    assert out.shape == (4, 10)
    torch.testing.assert_close(out.exp().sum(1), torch.ones(4), atol=1e-5, rtol=0)
    assert (out - ref).abs().max().item() < 0.02          # fp16 convolutions upstream; the head itself is fp32-accurate


def test_frozen_encoder_trainer_step(dev):
    """SURVEY §8 f4 (trainer/online_aug_trainer.py:21-35): encoder frozen in eval mode on the CUDA path, the MLP trains on the
    detached embeddings; after optimizer.step() the eval-mode forward (fused tail kernel) uses the UPDATED weights."""
    from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
    from vn_celeb_face_recognition_b200.trainer import FrozenEncoderTrainer
    from vn_celeb_face_recognition_b200 import synthetic
    torch.manual_seed(0)
    enc = InceptionResnetV1(device=dev).eval()
    enc.load_state_dict(synthetic.encoder_state_dict_seed0())
    mlp = MLPModel(512, 16).to(dev)
    tr = FrozenEncoderTrainer(mlp, enc, lr=1e-2, weight_decay=0.0)
    g = torch.Generator().manual_seed(1)
    data = torch.randn(32, 3, 160, 160, generator=g).clamp_(-1, 1)
    target = torch.arange(32) % 16
    enc_before = {k: v.clone() for k, v in enc.state_dict().items()}
    emb = tr.embed(data)
    assert emb.shape == (32, 512) and not emb.requires_grad
    assert torch.allclose(emb.norm(dim=1), torch.ones(32, device=dev), atol=1e-5)
    before = tr.validate_epoch([(data, target)])
    losses = [tr.train_step(data, target)[0] for _ in range(30)]
    after = tr.validate_epoch([(data, target)])
    assert after["val_neg_log_llhood"] < before["val_neg_log_llhood"] - 0.2, (before, after, losses[:3], losses[-3:])
    # the eval-mode pass went through the fused tail kernel with the updated weights: same as torch on the same parameters
    mlp.eval()
    with torch.no_grad():
        ref = torch.log_softmax(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(emb, mlp.dense_1.weight, mlp.dense_1.bias)),
                                                         mlp.dense_2.weight, mlp.dense_2.bias), dim=1)
    got = mlp(emb)
    assert (got - ref).abs().max().item() < 2e-5
    for k, v in enc.state_dict().items():
        assert torch.equal(v, enc_before[k]), k            # frozen
    ep = tr.train_epoch([(data, target), (data[:8], target[:8])])
    assert set(ep) == {"neg_log_llhood", "accuracy"} and 0.0 <= ep["accuracy"] <= 1.0
