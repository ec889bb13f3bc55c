"""CPU (torch fp32) restatement of the networks on the hot path, as pure functions of a ``state_dict``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each function cites the reference lines it follows.  The
``state_dict`` key names are the reference's so the same dict loads into the reference modules,
into this oracle, and into the CUDA drop-in classes.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------------
# MTCNN P/R/O nets -- models/mtcnn.py:9-157
# --------------------------------------------------------------------------------------------------------------------
def pnet_forward(sd, x):
    """models/mtcnn.py:38-49.  x (B,3,h,w) -> reg (B,4,h',w'), prob (B,2,h',w')."""
    x = F.prelu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"]), sd["prelu1.weight"])
    x = F.max_pool2d(x, 2, 2, ceil_mode=True)
    x = F.prelu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"]), sd["prelu2.weight"])
    x = F.prelu(F.conv2d(x, sd["conv3.weight"], sd["conv3.bias"]), sd["prelu3.weight"])
    a = F.softmax(F.conv2d(x, sd["conv4_1.weight"], sd["conv4_1.bias"]), dim=1)
    b = F.conv2d(x, sd["conv4_2.weight"], sd["conv4_2.bias"])
    return b, a


def rnet_forward(sd, x):
    """models/mtcnn.py:84-99.  x (N,3,24,24) -> reg (N,4), prob (N,2).  Flatten order is (W,H,C) (line 93)."""
    x = F.prelu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"]), sd["prelu1.weight"])
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)
    x = F.prelu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"]), sd["prelu2.weight"])
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)
    x = F.prelu(F.conv2d(x, sd["conv3.weight"], sd["conv3.bias"]), sd["prelu3.weight"])
    x = x.permute(0, 3, 2, 1).contiguous()
    x = F.prelu(F.linear(x.view(x.shape[0], -1), sd["dense4.weight"], sd["dense4.bias"]), sd["prelu4.weight"])
    a = F.softmax(F.linear(x, sd["dense5_1.weight"], sd["dense5_1.bias"]), dim=1)
    b = F.linear(x, sd["dense5_2.weight"], sd["dense5_2.bias"])
    return b, a


def onet_forward(sd, x):
    """models/mtcnn.py:138-157.  x (N,3,48,48) -> reg (N,4), landmarks (N,10), prob (N,2)."""
    x = F.prelu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"]), sd["prelu1.weight"])
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)
    x = F.prelu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"]), sd["prelu2.weight"])
    x = F.max_pool2d(x, 3, 2, ceil_mode=True)
    x = F.prelu(F.conv2d(x, sd["conv3.weight"], sd["conv3.bias"]), sd["prelu3.weight"])
    x = F.max_pool2d(x, 2, 2, ceil_mode=True)
    x = F.prelu(F.conv2d(x, sd["conv4.weight"], sd["conv4.bias"]), sd["prelu4.weight"])
    x = x.permute(0, 3, 2, 1).contiguous()
    x = F.prelu(F.linear(x.view(x.shape[0], -1), sd["dense5.weight"], sd["dense5.bias"]), sd["prelu5.weight"])
    a = F.softmax(F.linear(x, sd["dense6_1.weight"], sd["dense6_1.bias"]), dim=1)
    b = F.linear(x, sd["dense6_2.weight"], sd["dense6_2.bias"])
    c = F.linear(x, sd["dense6_3.weight"], sd["dense6_3.bias"])
    return b, c, a


# --------------------------------------------------------------------------------------------------------------------
# InceptionResnetV1 -- models/inception_resnet_v1.py:12-303
# --------------------------------------------------------------------------------------------------------------------
BN_EPS = 1e-3  # inception_resnet_v1.py:22 and :256


def _bconv(sd, p, x, stride=1, padding=0):
    """BasicConv2d, inception_resnet_v1.py:12-33: conv(no bias) -> BN(eps 1e-3, running stats) -> ReLU."""
    x = F.conv2d(x, sd[p + ".conv.weight"], None, stride=stride, padding=padding)
    x = F.batch_norm(x, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                     sd[p + ".bn.bias"], False, 0.0, BN_EPS)
    return F.relu(x)


def _block35(sd, p, x, scale):
    """inception_resnet_v1.py:36-67."""
    x0 = _bconv(sd, p + ".branch0", x)
    x1 = _bconv(sd, p + ".branch1.1", _bconv(sd, p + ".branch1.0", x), padding=1)
    x2 = _bconv(sd, p + ".branch2.0", x)
    x2 = _bconv(sd, p + ".branch2.1", x2, padding=1)
    x2 = _bconv(sd, p + ".branch2.2", x2, padding=1)
    out = F.conv2d(torch.cat((x0, x1, x2), 1), sd[p + ".conv2d.weight"], sd[p + ".conv2d.bias"])
    return F.relu(out * scale + x)


def _block17(sd, p, x, scale):
    """inception_resnet_v1.py:70-95."""
    x0 = _bconv(sd, p + ".branch0", x)
    x1 = _bconv(sd, p + ".branch1.0", x)
    x1 = _bconv(sd, p + ".branch1.1", x1, padding=(0, 3))
    x1 = _bconv(sd, p + ".branch1.2", x1, padding=(3, 0))
    out = F.conv2d(torch.cat((x0, x1), 1), sd[p + ".conv2d.weight"], sd[p + ".conv2d.bias"])
    return F.relu(out * scale + x)


def _block8(sd, p, x, scale, no_relu=False):
    """inception_resnet_v1.py:98-126."""
    x0 = _bconv(sd, p + ".branch0", x)
    x1 = _bconv(sd, p + ".branch1.0", x)
    x1 = _bconv(sd, p + ".branch1.1", x1, padding=(0, 1))
    x1 = _bconv(sd, p + ".branch1.2", x1, padding=(1, 0))
    out = F.conv2d(torch.cat((x0, x1), 1), sd[p + ".conv2d.weight"], sd[p + ".conv2d.bias"])
    out = out * scale + x
    return out if no_relu else F.relu(out)


def _mixed_6a(sd, x):
    """inception_resnet_v1.py:129-149."""
    x0 = _bconv(sd, "mixed_6a.branch0", x, stride=2)
    x1 = _bconv(sd, "mixed_6a.branch1.0", x)
    x1 = _bconv(sd, "mixed_6a.branch1.1", x1, padding=1)
    x1 = _bconv(sd, "mixed_6a.branch1.2", x1, stride=2)
    x2 = F.max_pool2d(x, 3, 2)
    return torch.cat((x0, x1, x2), 1)


def _mixed_7a(sd, x):
    """inception_resnet_v1.py:152-181."""
    x0 = _bconv(sd, "mixed_7a.branch0.1", _bconv(sd, "mixed_7a.branch0.0", x), stride=2)
    x1 = _bconv(sd, "mixed_7a.branch1.1", _bconv(sd, "mixed_7a.branch1.0", x), stride=2)
    x2 = _bconv(sd, "mixed_7a.branch2.0", x)
    x2 = _bconv(sd, "mixed_7a.branch2.1", x2, padding=1)
    x2 = _bconv(sd, "mixed_7a.branch2.2", x2, stride=2)
    x3 = F.max_pool2d(x, 3, 2)
    return torch.cat((x0, x1, x2, x3), 1)


def encoder_trunk(sd, x, taps=None):
    """inception_resnet_v1.py:281-293: everything up to (and including) block8.  ``taps`` (dict) collects
    intermediates by name for per-layer parity tests."""
    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    x = tap("conv2d_1a", _bconv(sd, "conv2d_1a", x, stride=2))
    x = tap("conv2d_2a", _bconv(sd, "conv2d_2a", x))
    x = tap("conv2d_2b", _bconv(sd, "conv2d_2b", x, padding=1))
    x = tap("maxpool_3a", F.max_pool2d(x, 3, 2))
    x = tap("conv2d_3b", _bconv(sd, "conv2d_3b", x))
    x = tap("conv2d_4a", _bconv(sd, "conv2d_4a", x))
    x = tap("conv2d_4b", _bconv(sd, "conv2d_4b", x, stride=2))
    for i in range(5):
        x = tap("repeat_1.%d" % i, _block35(sd, "repeat_1.%d" % i, x, 0.17))
    x = tap("mixed_6a", _mixed_6a(sd, x))
    for i in range(10):
        x = tap("repeat_2.%d" % i, _block17(sd, "repeat_2.%d" % i, x, 0.10))
    x = tap("mixed_7a", _mixed_7a(sd, x))
    for i in range(5):
        x = tap("repeat_3.%d" % i, _block8(sd, "repeat_3.%d" % i, x, 0.20))
    x = tap("block8", _block8(sd, "block8", x, 1.0, no_relu=True))
    return x


def encoder_forward(sd, x, classify=False, taps=None):
    """InceptionResnetV1.forward in eval mode, inception_resnet_v1.py:272-303."""
    x = encoder_trunk(sd, x, taps)
    x = F.adaptive_avg_pool2d(x, 1)
    x = F.linear(x.view(x.shape[0], -1), sd["last_linear.weight"])
    x = F.batch_norm(x, sd["last_bn.running_mean"], sd["last_bn.running_var"], sd["last_bn.weight"],
                     sd["last_bn.bias"], False, 0.0, BN_EPS)
    if classify:
        x = F.log_softmax(F.linear(x, sd["logits.weight"], sd["logits.bias"]), dim=1)
    else:
        x = F.normalize(x, p=2, dim=1)
    return x


def mlp_forward(sd, x):
    """MLPModel.forward in eval mode, models/mlp_model.py:10-15."""
    x = F.relu(F.linear(x, sd["dense_1.weight"], sd["dense_1.bias"]))
    x = F.linear(x, sd["dense_2.weight"], sd["dense_2.bias"])
    return F.log_softmax(x, dim=1)


# --------------------------------------------------------------------------------------------------------------------
# Encoder / MLP architecture table + seeded, well-conditioned random init
# --------------------------------------------------------------------------------------------------------------------
from vn_celeb_face_recognition_b200.synthetic import encoder_conv_specs  # noqa: E402,F401  (seeded weight DATA lives in the package)


def make_encoder_state_dict(seed=0, calibrate=True, calib_batch=None):
    """Seeded random-init InceptionResnetV1 ``state_dict`` with the reference's keys.

    Default PyTorch init + eval-mode BN makes the network degenerate (all inputs -> the same embedding, SURVEY.md
    section 4.5), so parity would pass for a kernel that ignores its input.  Here: kaiming-normal(fan_in, relu) conv
    weights, BN affine drawn near (1, 0) (vn_celeb_face_recognition_b200.synthetic.encoder_state_dict_uncalibrated), and
    -- when ``calibrate`` -- BN running statistics set layer by layer to the batch statistics of a seeded calibration
    batch pushed through the oracle itself.
    """
    from vn_celeb_face_recognition_b200 import synthetic
    sd = synthetic.encoder_state_dict_uncalibrated(seed)
    if calibrate:
        if calib_batch is None:
            from . import synth
            calib_batch = synth.crops_160(32, seed=99)      # real face statistics (bundled crops, seeded jitter)
        _calibrate_bn(sd, calib_batch)
    return sd


def _calibrate_bn(sd, x):
    """Sets every BN's running stats to the statistics its input has on ``x`` (sequentially, so later layers see
    calibrated earlier layers)."""
    import torch.nn.functional as F_
    real_bn = F_.batch_norm

    def calib_bn(inp, rm, rv, w, b, training, momentum, eps):
        dims = [0] + list(range(2, inp.dim()))
        with torch.no_grad():
            rm.copy_(inp.mean(dim=dims))
            rv.copy_(inp.var(dim=dims, unbiased=False) + 1e-5)
        return real_bn(inp, rm, rv, w, b, False, 0.0, eps)

    F_.batch_norm = calib_bn
    try:
        with torch.no_grad():
            encoder_forward(sd, x)
    finally:
        F_.batch_norm = real_bn


def make_mlp_state_dict(num_classes=1001, input_dim=512, seed=0):
    """MLPModel ``state_dict`` (models/mlp_model.py:6-8) with a wide seeded init so argmax depends on the input."""
    from vn_celeb_face_recognition_b200 import synthetic
    return synthetic.mlp_state_dict(num_classes, input_dim, seed)
