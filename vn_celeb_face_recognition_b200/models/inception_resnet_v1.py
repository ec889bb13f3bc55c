"""Drop-in for the reference's ``models.InceptionResnetV1`` (models/inception_resnet_v1.py:184-303).

Same constructor, same 714-key ``state_dict``, same ``forward(x)`` contract -- x: (B,3,H,W) fp32
standardised crops on the GPU, returns (B,512) fp32 L2-normalised embeddings (or log-probs when ``classify``) -- but
the forward pass is the sm_100a path: NHWC bf16 activations, tcgen05 implicit-GEMM convolutions with fused
BN/ReLU/residual/concat epilogues (csrc/igemm_conv.cu), driven by ``encoder_plan.EncoderPlan``.  The nn.Module tree
below only HOLDS parameters (for load_state_dict / state_dict / .to()); it is never used to compute.
There is no CPU fallback: calling forward without CUDA raises.
"""
import torch
from torch import nn

from .. import _lib, encoder_plan


def _basic(cin, cout, k):
    """Parameter holder with the reference's BasicConv2d names: .conv (no bias), .bn (eps 1e-3)  (:12-33)."""
    m = nn.Module()
    m.conv = nn.Conv2d(cin, cout, kernel_size=k, bias=False)
    m.bn = nn.BatchNorm2d(cout, eps=0.001, momentum=0.1, affine=True)
    return m


def _seq(*mods):
    return nn.Sequential(*mods)


def _block(kind):
    """Parameter holders of Block35 / Block17 / Block8 (:36-126)."""
    m = nn.Module()
    if kind == 35:
        m.branch0 = _basic(256, 32, 1)
        m.branch1 = _seq(_basic(256, 32, 1), _basic(32, 32, 3))
        m.branch2 = _seq(_basic(256, 32, 1), _basic(32, 32, 3), _basic(32, 32, 3))
        m.conv2d = nn.Conv2d(96, 256, kernel_size=1)
    elif kind == 17:
        m.branch0 = _basic(896, 128, 1)
        m.branch1 = _seq(_basic(896, 128, 1), _basic(128, 128, (1, 7)), _basic(128, 128, (7, 1)))
        m.conv2d = nn.Conv2d(256, 896, kernel_size=1)
    else:
        m.branch0 = _basic(1792, 192, 1)
        m.branch1 = _seq(_basic(1792, 192, 1), _basic(192, 192, (1, 3)), _basic(192, 192, (3, 1)))
        m.conv2d = nn.Conv2d(384, 1792, kernel_size=1)
    return m


class InceptionResnetV1(nn.Module):
    """See module docstring.  Keyword arguments as in the reference (inception_resnet_v1.py:202)."""

    #: crops per internal chunk: bounds activation memory (about 2.7 MB per crop).  Large chunks are faster: the small
    #: late layers (8x8 / 3x3 maps) only fill the 148 SMs at several hundred crops (measured 768 crops: 4.8 ms in one
    #: chunk vs 10 ms in six chunks of 128)
    chunk = 1024
    #: batch-size granularity of the cached plans of the fused pipeline path (embed_s2d)
    plan_bucket = 64

    def __init__(self, pretrained=None, classify=False, num_classes=None, dropout_prob=0.6, device=None):
        super().__init__()
        self.pretrained = pretrained
        self.classify = classify
        self.num_classes = num_classes
        if pretrained is not None:
            # inception_resnet_v1.py:306-331 downloads weights over HTTP; there is no network here
            raise Exception("pretrained=%r needs a network download; construct with pretrained=None and load a "
                            "state_dict" % (pretrained,))
        if self.classify and self.num_classes is None:
            raise Exception('If "pretrained" is not specified and "classify" is True, "num_classes" must be specified')

        self.conv2d_1a = _basic(3, 32, 3)
        self.conv2d_2a = _basic(32, 32, 3)
        self.conv2d_2b = _basic(32, 64, 3)
        self.conv2d_3b = _basic(64, 80, 1)
        self.conv2d_4a = _basic(80, 192, 3)
        self.conv2d_4b = _basic(192, 256, 3)
        self.repeat_1 = _seq(*[_block(35) for _ in range(5)])
        self.mixed_6a = nn.Module()
        self.mixed_6a.branch0 = _basic(256, 384, 3)
        self.mixed_6a.branch1 = _seq(_basic(256, 192, 1), _basic(192, 192, 3), _basic(192, 256, 3))
        self.repeat_2 = _seq(*[_block(17) for _ in range(10)])
        self.mixed_7a = nn.Module()
        self.mixed_7a.branch0 = _seq(_basic(896, 256, 1), _basic(256, 384, 3))
        self.mixed_7a.branch1 = _seq(_basic(896, 256, 1), _basic(256, 256, 3))
        self.mixed_7a.branch2 = _seq(_basic(896, 256, 1), _basic(256, 256, 3), _basic(256, 256, 3))
        self.repeat_3 = _seq(*[_block(8) for _ in range(5)])
        self.block8 = _block(8)
        self.last_linear = nn.Linear(1792, 512, bias=False)
        self.last_bn = nn.BatchNorm1d(512, eps=0.001, momentum=0.1, affine=True)
        if self.classify and self.num_classes is not None:
            self.logits = nn.Linear(512, self.num_classes)

        self._packed = None
        self._plans = {}
        self.device = torch.device("cpu")
        if device is not None:
            self.device = torch.device(device) if isinstance(device, str) else device
            self.to(device)

    # ---- weight lifecycle: any change of parameters invalidates the packed bf16 copies
    def _invalidate(self):
        self._packed = None
        self._plans = {}

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    def train(self, mode=True):
        if mode:
            # BN batch statistics / dropout are the training side-car (out of scope)
            pass
        return super().train(mode)

    #: 16-bit compute type (torch.float16 default / torch.bfloat16); None = encoder_plan.HALF
    half_dtype = None

    def _ensure(self, dev):
        if self._packed is None:
            self._packed = encoder_plan.EncoderWeights(self.state_dict(), dev, self.half_dtype)
            self._plans = {}
        return self._packed

    #: cached EncoderPlans (activation buffers + captured graph, ~2.7 MB per crop), least recently used first
    max_plans = 3

    def _plan(self, n, h, w, dev):
        key = (n, h, w)
        self._ensure(dev)
        plan = self._plans.pop(key, None)
        if plan is None:
            while len(self._plans) >= self.max_plans:          # evict the least recently used plan (unbounded cache)
                self._plans.pop(next(iter(self._plans)))
            plan = encoder_plan.EncoderPlan(self._packed, n, h, w, dev)
        self._plans[key] = plan                                # most recently used last
        return plan

    def _bucket(self, m):
        return min(self.chunk, -(-m // self.plan_bucket) * self.plan_bucket)

    def _tail_plan(self, plan, classifier):
        """The fused tail of ``plan``: last_linear+last_bn -> L2-normalise [-> classifier's dense_1/ReLU/dense_2/log-softmax]
        (or -> logits -> log_softmax when ``self.classify``), cached per (plan, classifier weights)."""
        from .. import tail
        dev = plan.x8.device
        if self.classify:
            key, layers = "logits", [(self._packed.last, "identity"), (self._packed.logits, "logsoftmax")]
        elif classifier is not None:
            cl = classifier.split_layers(dev)
            # keyed on the packed dense_1 object, which the cached plan keeps alive (a re-packed classifier is a new object)
            key, layers = ("mlp", id(cl[0][0])), [(self._packed.last, "l2norm")] + cl
        else:
            key, layers = "emb", [(self._packed.last, "l2norm")]
        tp = plan.tails.get(key)
        if tp is None:
            if len(plan.tails) > 3:
                plan.tails.clear()
            tp = plan.tails[key] = tail.TailPlan(layers, plan.n, dev, in_mode=0)
            tp.layers_ref = layers
        return tp

    def embed_s2d(self, x_s2d, size, classifier=None, threshold=0.0, thr_class=None, payload=None, want_half=False, mark=None):
        """Device-resident fast path used by the fused pipeline: x 16-bit space-to-depth crops (n, ceil(S/2), ceil(S/2),
        16) of S x S faces (vnfr_face_crops half_layout 1) -> dict(emb fp32 (n,512) L2-normalised[, emb16][, label int64,
        prob fp32 when ``classifier`` is given]).  One graph replay of the convolutions + ONE fused tail kernel per chunk
        of ``self.chunk`` crops.  ``payload``: (cap + 1, 514) fp32 all-gather send buffer -- embeddings / labels / probs
        are then written straight into its rows [0, n) and n into its count cell payload[cap, 0]."""
        n = x_s2d.shape[0]
        h = w = int(size)
        dev = x_s2d.device
        out = {}
        if payload is not None:
            assert payload.shape[1] == 514 and payload.shape[0] > n
            emb = payload[:n, :512]
        else:
            emb = torch.empty(n, 512, dtype=torch.float32, device=dev)
        out["emb"] = emb
        emb16 = torch.empty(n, 512, dtype=x_s2d.dtype, device=dev) if want_half else None
        out["emb16"] = emb16
        if classifier is not None:
            out["label"] = torch.empty(n, dtype=torch.int64, device=dev)
            out["prob"] = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            for s in range(0, n, self.chunk):
                m = min(self.chunk, n - s)
                # plans (activation buffers + captured graph) are cached per batch size: face counts vary from batch to batch,
                # so round up to a bucket of 64 crops (rows are independent; the padding rows are computed and ignored)
                plan = self._plan(self._bucket(m), h, w, dev)
                plan.x0[:m].copy_(x_s2d[s:s + m])
                plan.run()
                if mark is not None:
                    mark("encoder")                       # stage marker (bench.py): the convolutions end here
                tp = self._tail_plan(plan, classifier)
                last = s + m >= n
                tp.run(m, x=plan.x8.view(plan.n, -1, plan.x8.shape[-1]),
                       out_vecs=[emb[s:s + m]] + [None] * (len(tp.layers) - 1),
                       emb_half=None if emb16 is None else emb16[s:s + m],
                       label=out["label"][s:s + m] if classifier is not None else None,
                       prob=out["prob"][s:s + m] if classifier is not None else None,
                       payload=None if payload is None else payload[s:s + m],
                       thr=threshold, thr_class=thr_class, n_classes=classifier.num_classes if classifier is not None else 0,
                       count_cell=payload[-1, :1] if (payload is not None and last) else None, count_value=n)
        return out

    def forward(self, x):
        """inception_resnet_v1.py:272-303 (eval semantics)."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise _lib.VnfrError("InceptionResnetV1.forward needs a CUDA tensor: this package has no CPU path")
        if self.training:
            raise _lib.VnfrError("training-mode forward (batch-stat BN, dropout) is out of scope; call .eval()")
        dev = x.device
        x = x.contiguous().float()
        n, c, h, w = x.shape
        assert c == 3, "expected (B,3,H,W)"
        out = torch.empty(n, self.num_classes if self.classify else 512, dtype=torch.float32, device=dev)
        with torch.no_grad(), torch.cuda.device(dev):
            for s in range(0, n, self.chunk):
                m = min(self.chunk, n - s)
                plan = self._plan(self._bucket(m), h, w, dev)
                _lib.call("vnfr_nchw3_to_s2d16", _lib.ptr(x[s:s + m]), m, h, w, _lib.ptr(plan.x0),
                          encoder_plan.dtype_code(plan.dtype), _lib.stream_ptr())
                plan.run()
                tp = self._tail_plan(plan, None)
                x8 = plan.x8.view(plan.n, -1, plan.x8.shape[-1])
                if self.classify:
                    # logits + log_softmax (inception_resnet_v1.py:298-300)
                    tp.run(m, x=x8, out_vecs=[None, out[s:s + m]], n_classes=self.num_classes)
                else:
                    tp.run(m, x=x8, out_vecs=[out[s:s + m]])
        return out
