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

    def _plan(self, n, h, w, dev):
        key = (n, h, w)
        if key not in self._plans:
            self._plans[key] = encoder_plan.EncoderPlan(self._ensure(dev), n, h, w, dev)
        return self._plans[key]

    def embed_s2d(self, x_s2d, size):
        """Device-resident fast path used by the fused pipeline: x 16-bit space-to-depth crops (n, ceil(S/2), ceil(S/2),
        16) of S x S faces (vnfr_face_crops half_layout 1) -> (emb fp32 (n,512), emb 16-bit (n,512)), both
        L2-normalised.  Chunked by ``self.chunk``."""
        n = x_s2d.shape[0]
        h = w = int(size)
        dev = x_s2d.device
        emb = torch.empty(n, 512, dtype=torch.float32, device=dev)
        emb16 = torch.empty(n, 512, dtype=x_s2d.dtype, device=dev)
        for s in range(0, n, self.chunk):
            m = min(self.chunk, n - s)
            # plans (activation buffers + captured graph) are cached per batch size: face counts vary from batch to batch,
            # so round up to a bucket of 64 crops (rows are independent; the padding rows are computed and ignored)
            plan = self._plan(min(self.chunk, -(-m // self.plan_bucket) * self.plan_bucket), h, w, dev)
            plan.x0[:m].copy_(x_s2d[s:s + m])
            plan.run()
            _lib.call("vnfr_l2norm_rows", _lib.ptr(plan.emb_raw), m, 512, 512, _lib.ptr(emb[s:s + m]),
                      _lib.ptr(emb16[s:s + m]), encoder_plan.dtype_code(plan.dtype), _lib.stream_ptr())
        return emb, emb16

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
        with torch.no_grad():
            for s in range(0, n, self.chunk):
                m = min(self.chunk, n - s)
                plan = self._plan(m, h, w, dev)
                _lib.call("vnfr_nchw3_to_s2d16", _lib.ptr(x[s:s + m]), m, h, w, _lib.ptr(plan.x0),
                          encoder_plan.dtype_code(plan.dtype), _lib.stream_ptr())
                plan.run()
                if self.classify:
                    self._classify(plan, m, out[s:s + m])
                else:
                    _lib.call("vnfr_l2norm_rows", _lib.ptr(plan.emb_raw), m, 512, 512, _lib.ptr(out[s:s + m]), None, 0,
                              _lib.stream_ptr())
        return out

    def _classify(self, plan, m, out):
        # logits + log_softmax (inception_resnet_v1.py:298-300)
        pc = self._packed.P["logits"]
        if not hasattr(plan, "cls"):
            x16 = torch.empty(m, 1, 1, 512, dtype=plan.dtype, device=out.device)
            logits = torch.empty(m, pc.cout, dtype=torch.float32, device=out.device)
            ol = encoder_plan.OpList()
            ol.conv(pc, encoder_plan.View(x16), None, relu=False, out_f32=logits)
            plan.cls = (x16, logits, ol)
        x16, logits, ol = plan.cls
        x16.view(m, 512).copy_(plan.emb_raw)
        ol.run()
        _lib.call("vnfr_logsoftmax_argmax", _lib.ptr(logits), m, self.num_classes, logits.shape[1], _lib.ptr(out), None,
                  None, _lib.stream_ptr())
