"""Host-side graph of the InceptionResnetV1 encoder and the MLP classifier as a flat list of C-ABI ops.

Weight packing (done once per ``load_state_dict`` / first forward, on the GPU with torch tensor ops -- plumbing):
  * BatchNorm (eval, eps 1e-3) is folded into the conv: W' = W * gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)
    (inception_resnet_v1.py:12-33); the residual scale of Block35/17/8 is folded into the projection conv
    (``out*scale + x``, :64-66, :92-94, :121-125).
  * weights go to 16-bit [cout_pad][k_pad] with k = (ky*KW + kx)*cin + c (NHWC implicit-GEMM order), zero padded.
  * sibling 1x1 branch convs that read the same input are concatenated along cout and run as ONE GEMM whose epilogue
    scatters column ranges to different destinations (the concat buffer / the branch scratch).
Activations are NHWC 16-bit (fp16 / bf16); ``torch.cat`` never runs: every conv writes straight into its channel slice.
"""
import ctypes as C
import os

import torch

from . import _lib

BN_EPS = 1e-3

# 16-bit storage/compute type of the encoder and classifier (fp32 accumulation in TMEM either way; identical tensor-core
# rate).  fp16 is the default because its 11-bit significand keeps the 130-layer encoder at cosine >= 0.9999 vs the fp32
# reference, where bf16 (8 bits) measures 0.9977-0.9997 -- below the 0.999 parity bar.  VNFR_HALF_DTYPE=bf16 selects bf16.
HALF = torch.bfloat16 if os.environ.get("VNFR_HALF_DTYPE", "fp16").lower() in ("bf16", "bfloat16") else torch.float16


# channels-per-plane of the shifted-view kernel by input channel count (0 / missing = generic gather kernel).
# 32-channel inputs use 64-byte swizzle rows; everything else 128-byte rows.
SV_DEFAULT = {16: 16, 32: 32, 64: 64, 128: 64, 192: 64, 256: 64}
if os.environ.get("VNFR_NO_SV"):
    SV_DEFAULT = {}
for _c in os.environ.get("VNFR_SV_OFF_CIN", "").split(","):      # experiments: generic gather kernel for these input widths
    if _c.strip():
        SV_DEFAULT.pop(int(_c), None)


#: minimum fraction of real (non-padding) pixels in a shifted-view band
SV_MIN_USEFUL = float(os.environ.get("VNFR_SV_MIN_USEFUL", "0.7"))

USE_GRAPHS = not os.environ.get("VNFR_NO_GRAPH")
#: N tile of mixed_7a's two 3x3 / stride-2 convolutions with 256 outputs (8x8 -> 3x3 maps: 54 row tiles)
M7A_BLOCK_N = int(os.environ.get("VNFR_M7A_BLOCK_N", "128")) or None      # measured: 256 -> 21.5 us, 128 -> 19.0 us, 64 -> 30.9 us per launch
#: N tile of Block8's 1x3 / 3x1 convolutions (192 channels on 54 row tiles: one N tile leaves 94 SMs idle)
B8_BLOCK_N = int(os.environ.get("VNFR_B8_BLOCK_N", "96")) or None      # measured: 192 -> 126 us, 96 -> 116 us, 64 -> 152 us (12 launches)
#: Block17 as one fused kernel per block (csrc/block17_fused.cu) when its map is 8x8 (160x160 crops); VNFR_NO_FUSED_B17=1
#: keeps the four-launch form (A/B measurements)
FUSED_B17 = not os.environ.get("VNFR_NO_FUSED_B17")
#: crops per pass of the Block35 section (VNFR_B35_CHUNK).  Default: one pass over the whole batch -- measured for 768 crops
#: (encoder stage): one pass 4.16 ms, chunks of 384 / 256 / 192 / 128 crops 4.22 / 4.32 / 4.48 / 4.67 ms: keeping a chunk's trunk
#: L2-resident across the five blocks does not pay for the extra launches (the section is launch- / latency-bound, not DRAM-bound)
B35_CHUNK = int(os.environ.get("VNFR_B35_CHUNK", str(1 << 30)))
#: Block35's two parallel 3x3 convolutions as one block-diagonal launch (VNFR_NO_GROUPED_B35=1: two launches)
GROUPED_B35 = not os.environ.get("VNFR_NO_GROUPED_B35")


def dtype_code(dt):
    return 1 if dt == torch.float16 else 0


def _ceil(a, b):
    return (a + b - 1) // b * b


class PackedConv:
    """Device-resident packed weights of one (possibly branch-fused) convolution."""

    def __init__(self, w, bias, kh, kw, cin, cout, block_n):
        self.w, self.bias, self.kh, self.kw, self.cin, self.cout, self.block_n = w, bias, kh, kw, cin, cout, block_n
        self.cout_pad, self.k_pad = w.shape


def pick_block_n(cout):
    """N tile: one tcgen05.mma costs the same for any N <= 256 (the shared-memory read of the 128-row A operand is the
    floor), so tiles are as wide as possible; multiples of 64 keep the TMA-store panels inside their tile."""
    if cout <= 256:
        return cout
    for bn in (256, 192):
        if cout % bn == 0:
            return bn
    return 256            # zero-padded last tile (cout_pad), e.g. 896 -> 4 x 256


def pack_conv(w, scale, bias, device, cin_pad=None, block_n=None, dtype=None):
    """w (cout,cin,kh,kw) fp32, scale (cout,) or None, bias (cout,) fp32 -> PackedConv on ``device``."""
    w = w.detach().to(device=device, dtype=torch.float32)
    cout, cin, kh, kw = w.shape
    if scale is not None:
        w = w * scale.to(device).view(-1, 1, 1, 1)
    cin_p = cin_pad or _ceil(cin, 8)
    w = w.permute(0, 2, 3, 1)                                   # cout, kh, kw, cin
    if cin_p != cin:
        w = torch.nn.functional.pad(w, (0, cin_p - cin))
    K = kh * kw * cin_p
    bn = block_n or pick_block_n(_ceil(cout, 16))
    cout16 = _ceil(cout, 16)
    cout_pad = _ceil(cout16, bn)
    k_pad = _ceil(K, 64)
    dtype = dtype or HALF
    wp = torch.zeros(cout_pad, k_pad, dtype=dtype, device=device)
    wp[:cout, :K] = w.reshape(cout, K).clamp(-65504.0, 65504.0).to(dtype)
    bp = torch.zeros(cout_pad, dtype=torch.float32, device=device)
    bp[:cout] = bias.detach().to(device=device, dtype=torch.float32)
    return PackedConv(wp, bp, kh, kw, cin_p, cout16, bn)


def split3_bf16(x):
    """fp32 tensor -> (hi, mid, lo) bf16 parts with hi + mid + lo == x to ~2^-24 relative (each part is the bf16 rounding
    of what the previous parts left over)."""
    x = x.float()
    hi = x.to(torch.bfloat16)
    r1 = x - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def split2_fp16(x):
    """fp32 tensor -> (hi, lo) fp16 parts with hi + lo == x to ~2^-22 relative (fp16 keeps 11 significant bits per part;
    small residuals fall into fp16's subnormal range, whose absolute step 6e-8 is far below the rounding of the large
    terms of a dot product)."""
    x = x.float()
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return hi, lo


def pack_conv_split2(w, bias, device, ck):
    """Two-part fp16 variant of pack_conv_split3 (VnfrConvOp.split3 = 2): fp16 [cout][taps*3*ck], K step (tap, j) holding
    weight part (0,1,0)[j] -- the partner of activation part (0,0,1)[j]; the dropped lo*lo product is O(2^-22)."""
    w = w.detach().to(device=device, dtype=torch.float32)
    cout, cin, kh, kw = w.shape
    assert cin <= ck and cout % 16 == 0
    wp = torch.zeros(cout, kh * kw, ck, dtype=torch.float32, device=device)
    wp[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    parts = split2_fp16(wp)
    packed = torch.stack([parts[j] for j in (0, 1, 0)], dim=2).reshape(cout, kh * kw * 3 * ck).contiguous()
    k_pad = _ceil(packed.shape[1], 64)
    if k_pad != packed.shape[1]:
        packed = torch.nn.functional.pad(packed, (0, k_pad - packed.shape[1]))
    pc = PackedConv(packed.contiguous(), bias.detach().to(device=device, dtype=torch.float32).contiguous(), kh, kw, 2 * ck, cout, cout)
    pc.split3 = 2
    return pc


def pack_conv_split3(w, bias, device, ck):
    """Split-precision weights for the shifted-view kernel (VnfrConvOp.split3): w (cout, cin, kh, kw) fp32 with
    cin <= ck -> bf16 [cout][taps*6*ck], K step (tap, j) holding weight part (0,1,0,2,1,0)[j] -- the partner of activation
    part (0,0,1,0,1,2)[j] -- so that the six bf16 products sum to the fp32 product up to O(2^-24)."""
    w = w.detach().to(device=device, dtype=torch.float32)
    cout, cin, kh, kw = w.shape
    assert cin <= ck and cout % 16 == 0
    wp = torch.zeros(cout, kh * kw, ck, dtype=torch.float32, device=device)
    wp[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    parts = split3_bf16(wp)
    order = (0, 1, 0, 2, 1, 0)
    packed = torch.stack([parts[j] for j in order], dim=2)                 # cout, taps, 6, ck
    packed = packed.reshape(cout, kh * kw * 6 * ck).contiguous()
    k_pad = _ceil(packed.shape[1], 64)
    if k_pad != packed.shape[1]:
        packed = torch.nn.functional.pad(packed, (0, k_pad - packed.shape[1]))
    pc = PackedConv(packed.contiguous(), bias.detach().to(device=device, dtype=torch.float32).contiguous(), kh, kw, 3 * ck, cout, cout)
    pc.split3 = 1
    return pc


def fold_bn(sd, p):
    """BasicConv2d ``p`` -> (weight, scale, bias) with eval-mode BN folded."""
    g, b = sd[p + ".bn.weight"].float(), sd[p + ".bn.bias"].float()
    m, v = sd[p + ".bn.running_mean"].float(), sd[p + ".bn.running_var"].float()
    s = g / torch.sqrt(v + BN_EPS)
    return sd[p + ".conv.weight"].float(), s, b - m * s


def pack_basic(sd, prefixes, device, cin_pad=None, block_n=None, dtype=None):
    """One or more sibling BasicConv2d (same input, same kernel) fused along cout."""
    ws, bs = [], []
    for p in prefixes:
        w, s, b = fold_bn(sd, p)
        ws.append(w * s.view(-1, 1, 1, 1))
        bs.append(b)
    return pack_conv(torch.cat(ws, 0), None, torch.cat(bs, 0), device, cin_pad, block_n, dtype)


def pack_stem_s2d(sd, prefix, device, dtype=None):
    """conv2d_1a (3x3, stride 2, 3 -> 32; inception_resnet_v1.py:219) re-expressed on the space-to-depth input
    [n][H/2][W/2][16] (channel = ((y&1)*2 + (x&1))*4 + c): a stride-1 2x2 convolution over 16 channels whose tap (ty,tx)
    / sub-pixel (sy,sx) weight is the original tap (ky,kx) = (2ty+sy, 2tx+sx) (zero where ky or kx would be 3)."""
    w, s, b = fold_bn(sd, prefix)
    w = (w * s.view(-1, 1, 1, 1)).float()                       # (32, 3, 3, 3) [co][ci][ky][kx]
    co = w.shape[0]
    w2 = torch.zeros(co, 16, 2, 2, dtype=torch.float32, device=w.device)
    for ty in range(2):
        for tx in range(2):
            for sy in range(2):
                for sx in range(2):
                    ky, kx = 2 * ty + sy, 2 * tx + sx
                    if ky < 3 and kx < 3:
                        w2[:, (sy * 2 + sx) * 4:(sy * 2 + sx) * 4 + 3, ty, tx] = w[:, :, ky, kx]
    return pack_conv(w2, None, b, device, cin_pad=16, dtype=dtype)


def pack_basic_grouped(sd, prefixes, device, dtype=None):
    """Sibling BasicConv2d with the SAME kernel size but DIFFERENT inputs (each reads its own channel slice of one tensor, in
    order) as ONE block-diagonal convolution: cout = sum of couts, cin = sum of cins, zero weights off the diagonal.  Block35's
    branch1.1 and branch2.1 (3x3, 32 -> 32 each, inception_resnet_v1.py:44-51) become one 64 -> 64 launch: the tensor pipe
    does twice the (tiny) work, but a launch of this size is latency-bound, so one launch costs about what each of the two did."""
    ws, bs = [], []
    for p in prefixes:
        w, s, b = fold_bn(sd, p)
        ws.append(w * s.view(-1, 1, 1, 1))
        bs.append(b)
    cout, cin = sum(w.shape[0] for w in ws), sum(w.shape[1] for w in ws)
    kh, kw = ws[0].shape[2:]
    big = torch.zeros(cout, cin, kh, kw, dtype=torch.float32, device=ws[0].device)
    o = i = 0
    for w in ws:
        big[o:o + w.shape[0], i:i + w.shape[1]] = w
        o += w.shape[0]
        i += w.shape[1]
    return pack_conv(big, None, torch.cat(bs, 0), device, None, None, dtype)


def pack_projection(sd, p, scale, device, block_n=None, dtype=None):
    """Block projection conv2d (with bias), residual scale folded in."""
    # N tiles of 128: the K x 128 weight slab of a projection (K = 96 / 256 / 384) then stays resident in shared memory
    # (igemm_conv.cu use_resident_weights) and two C staging buffers fit
    if block_n is None:
        block_n = int(os.environ.get("VNFR_PROJ_BLOCK_N", "128"))
    return pack_conv(sd[p + ".weight"].float() * scale, None, sd[p + ".bias"].float() * scale, device, None, block_n, dtype)


class View:
    """A channel slice of an NHWC 16-bit (fp16 / bf16) activation buffer."""

    def __init__(self, t, c0=0, c=None):
        self.t, self.c0 = t, c0
        self.c = (t.shape[-1] - c0) if c is None else c

    @property
    def pitch(self):
        return self.t.shape[-1]

    @property
    def ptr(self):
        return self.t.data_ptr() + 2 * self.c0

    @property
    def n(self):
        return self.t.shape[0]

    @property
    def h(self):
        return self.t.shape[1]

    @property
    def w(self):
        return self.t.shape[2]


class OpList:
    """Builds and owns a ctypes array of VnfrOp plus references to every tensor it points at."""

    def __init__(self):
        self.ops = []
        self.keep = []
        self._arr = None
        self._graph, self._runs = None, 0

    def conv(self, pc, src, dst0, stride=1, pad=(0, 0), relu=True, dst1=None, n_split=None, residual=None, out_f32=None,
             sv=None, alpha=None, n_img_dev=None):
        """``sv`` = 32 / 64 requests the shifted-view kernel (csrc/sv_conv.cu) with that many channels per plane; the
        packed weights must then be laid out with cin padded to a multiple of ``sv`` (pc.cin)."""
        split3 = int(getattr(pc, "split3", 0))
        if sv is None:
            sv = SV_DEFAULT.get(src.c) if (stride == 1 and pc.kh * pc.kw > 1 and pc.cout <= 256 and out_f32 is None
                                            and pc.block_n == pc.cout == pc.cout_pad) else 0
            if sv and pc.cin != _ceil(src.c, sv):
                sv = 0
            # the shifted-view kernel issues MMA rows for the zero padding too: on small maps with wide padding (1x7 / 7x1
            # on 8x8: 8 of 14 band pixels are real) the generic gather kernel is faster (measured 31 / 38 us -> 25 / 25 us)
            if sv and (src.h * src.w) < SV_MIN_USEFUL * (src.h + 2 * pad[0]) * (src.w + 2 * pad[1]):
                sv = 0
        assert (src.c == pc.cin) or (sv and pc.cin == _ceil(src.c, sv)), (src.c, pc.cin, sv)
        op = _lib.Op()
        op.kind = 0
        c = op.conv
        c.inp, c.weights, c.bias = src.ptr, pc.w.data_ptr(), pc.bias.data_ptr()
        c.n_img, c.in_h, c.in_w, c.cin, c.in_pitch = src.n, src.h, src.w, src.c, src.pitch
        c.kh, c.kw, c.stride, c.pad_h, c.pad_w = pc.kh, pc.kw, stride, pad[0], pad[1]
        c.out_h = (src.h + 2 * pad[0] - pc.kh) // stride + 1
        c.out_w = (src.w + 2 * pad[1] - pc.kw) // stride + 1
        c.cout, c.cout_pad, c.k_pad, c.block_n = pc.cout, pc.cout_pad, pc.k_pad, pc.block_n
        c.relu = 1 if relu else 0
        c.dtype = dtype_code(pc.w.dtype)
        assert src.t.dtype == pc.w.dtype, "activation / weight dtype mismatch"
        if out_f32 is not None:
            c.out_f32, c.out_f32_pitch = out_f32.data_ptr(), out_f32.shape[-1]
            c.n_split = pc.cout
            self.keep.append(out_f32)
        else:
            c.out0, c.out0_pitch = dst0.ptr, dst0.pitch
            assert (dst0.n, dst0.h, dst0.w) == (src.n, c.out_h, c.out_w), "destination geometry mismatch"
            if dst1 is not None:
                c.out1, c.out1_pitch, c.n_split = dst1.ptr, dst1.pitch, n_split
                assert dst0.c == n_split and dst1.c == pc.cout - n_split
            else:
                c.n_split = pc.cout
                assert dst0.c == pc.cout, (dst0.c, pc.cout)
        if residual is not None:
            c.residual, c.res_pitch = residual.ptr, residual.pitch
        c.reserved[0] = int(sv or 0)
        c.split3 = split3
        if alpha is not None:
            c.prelu_alpha = alpha.data_ptr()
            self.keep.append(alpha)
        if n_img_dev is not None:
            c.n_img_dev = n_img_dev.data_ptr()
            self.keep.append(n_img_dev)
        _lib.call("vnfr_conv_prepare", C.byref(c))
        if split3 and c.a_mode != 3:
            raise _lib.VnfrError("split-precision convolution needs the shifted-view kernel, but the geometry does not qualify")
        if pc.cin != src.c and c.a_mode != 3:
            raise _lib.VnfrError("weights were packed for the shifted-view kernel but the geometry does not qualify")
        self.ops.append(op)
        self.keep += [pc, src, dst0, dst1, residual]
        self._arr = None

    def block17(self, p_in, p_a, p_b, p_out, x):
        """One fused Block17 (csrc/block17_fused.cu), in place on ``x`` (n, 8, 8, 896): ``p_in`` = branch0 | branch1.0 packed
        as one N = 256 GEMM, ``p_a`` / ``p_b`` the 1x7 / 7x1 convolutions, ``p_out`` the projection (residual scale folded)."""
        assert tuple(x.shape[1:]) == (8, 8, 896) and x.is_contiguous()
        assert tuple(p_in.w.shape) == (256, 896) and tuple(p_a.w.shape) == (128, 896) and tuple(p_b.w.shape) == (128, 896)
        assert tuple(p_out.w.shape) == (896, 256) and (p_a.kh, p_a.kw, p_b.kh, p_b.kw) == (1, 7, 7, 1)
        b = _lib.Block17Op()
        b.x, b.n_img, b.dtype = x.data_ptr(), x.shape[0], dtype_code(x.dtype)
        b.w1, b.w2, b.w3, b.w4 = p_in.w.data_ptr(), p_a.w.data_ptr(), p_b.w.data_ptr(), p_out.w.data_ptr()
        b.b1, b.b2, b.b3, b.b4 = p_in.bias.data_ptr(), p_a.bias.data_ptr(), p_b.bias.data_ptr(), p_out.bias.data_ptr()
        _lib.call("vnfr_block17_prepare", C.byref(b))
        op = _lib.Op()
        op.kind = 3
        op.ext = C.addressof(b)
        self.ops.append(op)
        self.keep += [b, p_in, p_a, p_b, p_out, x]
        self._arr = None

    def maxpool(self, src, dst):
        op = _lib.Op()
        op.kind = 1
        c = op.conv
        c.inp, c.out0 = src.ptr, dst.ptr
        c.n_img, c.in_h, c.in_w, c.cin, c.in_pitch, c.out0_pitch = src.n, src.h, src.w, src.c, src.pitch, dst.pitch
        assert dst.c == src.c and dst.h == (src.h - 3) // 2 + 1
        c.dtype = dtype_code(src.t.dtype)
        self.ops.append(op)
        self.keep += [src, dst]
        self._arr = None

    def avgpool(self, src, dst2d):
        op = _lib.Op()
        op.kind = 2
        c = op.conv
        c.inp, c.out0 = src.ptr, dst2d.data_ptr()
        c.n_img, c.in_h, c.in_w, c.cin, c.in_pitch = src.n, src.h, src.w, src.c, src.pitch
        c.dtype = dtype_code(src.t.dtype)
        self.ops.append(op)
        self.keep += [src, dst2d]
        self._arr = None

    def run(self):
        if not self.ops:
            return
        if self._arr is None:
            self._arr = (_lib.Op * len(self.ops))(*self.ops)
            self._graph, self._runs = None, 0
        if self._graph is not None:
            self._graph.replay()                        # one launch of the whole op list (no per-kernel launch gaps)
            _lib.call("vnfr_count_launches", len(self.ops))
            return
        _lib.call("vnfr_run_ops", self._arr, len(self.ops), _lib.stream_ptr())
        self._runs += 1
        if USE_GRAPHS and self._runs == 2 and len(self.ops) >= 8 and not torch.cuda.is_current_stream_capturing():
            # the list is static (fixed pointers and shapes): capture it into a CUDA graph after two eager runs
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    _lib.call("vnfr_run_ops", self._arr, len(self.ops), _lib.stream_ptr())
                self._graph = g
            except Exception:
                self._graph = None


class EncoderWeights:
    """All packed convolutions of InceptionResnetV1 (inception_resnet_v1.py:219-257), fused per the module doc."""

    def __init__(self, sd, device, dtype=None):
        d = device
        self.dtype = dtype or HALF
        import functools
        pack_basic = functools.partial(globals()["pack_basic"], dtype=self.dtype)
        pack_projection = functools.partial(globals()["pack_projection"], dtype=self.dtype)
        pack_conv = functools.partial(globals()["pack_conv"], dtype=self.dtype)
        P = {}
        P["conv2d_1a"] = pack_stem_s2d(sd, "conv2d_1a", d, dtype=self.dtype)
        for n in ["conv2d_2a", "conv2d_2b", "conv2d_3b", "conv2d_4a", "conv2d_4b"]:
            P[n] = pack_basic(sd, [n], d)
        for i in range(5):
            p = "repeat_1.%d" % i
            P[p + ".in"] = pack_basic(sd, [p + ".branch0", p + ".branch1.0", p + ".branch2.0"], d)      # N = 96
            P[p + ".b1"] = pack_basic(sd, [p + ".branch1.1"], d)
            P[p + ".b2a"] = pack_basic(sd, [p + ".branch2.1"], d)
            P[p + ".b12"] = pack_basic_grouped(sd, [p + ".branch1.1", p + ".branch2.1"], d, dtype=self.dtype)
            P[p + ".b2b"] = pack_basic(sd, [p + ".branch2.2"], d)
            P[p + ".out"] = pack_projection(sd, p + ".conv2d", 0.17, d)
        P["m6a.b0"] = pack_basic(sd, ["mixed_6a.branch0"], d)
        P["m6a.b1a"] = pack_basic(sd, ["mixed_6a.branch1.0"], d)
        P["m6a.b1b"] = pack_basic(sd, ["mixed_6a.branch1.1"], d)
        P["m6a.b1c"] = pack_basic(sd, ["mixed_6a.branch1.2"], d)
        for i in range(10):
            p = "repeat_2.%d" % i
            P[p + ".in"] = pack_basic(sd, [p + ".branch0", p + ".branch1.0"], d, block_n=256)            # N = 256
            P[p + ".b1a"] = pack_basic(sd, [p + ".branch1.1"], d)
            P[p + ".b1b"] = pack_basic(sd, [p + ".branch1.2"], d)
            P[p + ".out"] = pack_projection(sd, p + ".conv2d", 0.10, d)
        P["m7a.in"] = pack_basic(sd, ["mixed_7a.branch0.0", "mixed_7a.branch1.0", "mixed_7a.branch2.0"], d)   # N = 768
        P["m7a.b0"] = pack_basic(sd, ["mixed_7a.branch0.1"], d)
        P["m7a.b1"] = pack_basic(sd, ["mixed_7a.branch1.1"], d, block_n=M7A_BLOCK_N)
        P["m7a.b2a"] = pack_basic(sd, ["mixed_7a.branch2.1"], d)
        P["m7a.b2b"] = pack_basic(sd, ["mixed_7a.branch2.2"], d, block_n=M7A_BLOCK_N)
        for i in list(range(5)) + [None]:
            p = "repeat_3.%d" % i if i is not None else "block8"
            P[p + ".in"] = pack_basic(sd, [p + ".branch0", p + ".branch1.0"], d, block_n=192)            # N = 384
            P[p + ".b1a"] = pack_basic(sd, [p + ".branch1.1"], d, block_n=B8_BLOCK_N)
            P[p + ".b1b"] = pack_basic(sd, [p + ".branch1.2"], d, block_n=B8_BLOCK_N)
            P[p + ".out"] = pack_projection(sd, p + ".conv2d", 0.20 if i is not None else 1.0, d)
        # last_linear (no bias) + last_bn folded (inception_resnet_v1.py:296-297)
        g, b = sd["last_bn.weight"].float(), sd["last_bn.bias"].float()
        m, v = sd["last_bn.running_mean"].float(), sd["last_bn.running_var"].float()
        s = g / torch.sqrt(v + BN_EPS)
        # the tail runs in split precision inside the fused tail kernel (tail.py / csrc/tail_fused.cu)
        from . import tail
        self.last = tail.SplitLinear(sd["last_linear.weight"].float().to(d) * s.to(d).view(-1, 1), b - m * s, d)
        self.logits = None
        if "logits.weight" in sd:
            self.logits = tail.SplitLinear(sd["logits.weight"].float(), sd["logits.bias"].float(), d)
        self.P = P


def _out_hw(h, k, s, p=0):
    return (h + 2 * p - k) // s + 1


class EncoderPlan:
    """Op list + activation buffers of one forward for a fixed (batch, H, W).  Input: ``self.x0``, the 16-bit
    space-to-depth crop tensor (n, ceil(H/2), ceil(W/2), 16) (see pack_stem_s2d); output: ``self.x8``, the block8 activations
    (n, h7, w7, 1792) that the fused tail kernel pools (avgpool_1a -> last_linear -> last_bn -> ..., tail.py)."""

    def __init__(self, weights, n, h, w, device):
        P = weights.P
        bf = dict(dtype=weights.dtype, device=device)
        self.dtype = weights.dtype
        buf = lambda hh, ww, c: torch.empty(n, hh, ww, c, **bf)
        ol = OpList()
        self.ol = ol
        self.n = n
        self.x0 = torch.zeros(n, (h + 1) // 2, (w + 1) // 2, 16, **bf)
        h1, w1 = _out_hw(h, 3, 2), _out_hw(w, 3, 2)
        assert h1 == (h + 1) // 2 - 1 and w1 == (w + 1) // 2 - 1
        c1a = buf(h1, w1, 32)
        h2, w2 = h1 - 2, w1 - 2
        c2a, c2b = buf(h2, w2, 32), buf(h2, w2, 64)
        h3, w3 = _out_hw(h2, 3, 2), _out_hw(w2, 3, 2)
        mp, c3b = buf(h3, w3, 64), buf(h3, w3, 80)
        h4, w4 = h3 - 2, w3 - 2
        c4a = buf(h4, w4, 192)
        h5, w5 = _out_hw(h4, 3, 2), _out_hw(w4, 3, 2)
        x35 = buf(h5, w5, 256)
        ol.conv(P["conv2d_1a"], View(self.x0), View(c1a), sv=16)
        ol.conv(P["conv2d_2a"], View(c1a), View(c2a))
        ol.conv(P["conv2d_2b"], View(c2a), View(c2b), pad=(1, 1))
        ol.maxpool(View(c2b), View(mp))
        ol.conv(P["conv2d_3b"], View(mp), View(c3b))
        ol.conv(P["conv2d_4a"], View(c3b), View(c4a))
        ol.conv(P["conv2d_4b"], View(c4a), View(x35), stride=2)
        # ---- 5 x Block35 (scale 0.17)
        cat_f, t1_f, t2_f = buf(h5, w5, 96), buf(h5, w5, 64), buf(h5, w5, 32)
        # Optionally (VNFR_B35_CHUNK) the five blocks run chunk by chunk over the batch so that a chunk's trunk + branch buffers
        # (0.26 MB per crop) stay in L2 across its 20 launches; measured slower than one pass (see B35_CHUNK), so off by default.
        n_b35 = max(1, min(n, -(-n // B35_CHUNK)))
        bounds35 = [(k * n // n_b35, (k + 1) * n // n_b35) for k in range(n_b35)]
        for a0, a1 in bounds35:
            if a1 <= a0:
                continue
            xs, cat, t1, t2 = x35[a0:a1], cat_f[a0:a1], t1_f[a0:a1], t2_f[a0:a1]
            for i in range(5):
                p = "repeat_1.%d" % i
                ol.conv(P[p + ".in"], View(xs), View(cat, 0, 32), dst1=View(t1), n_split=32)
                if GROUPED_B35:
                    # branch1.1 and branch2.1 as one block-diagonal 64 -> 64 convolution writing both destinations
                    ol.conv(P[p + ".b12"], View(t1), View(cat, 32, 32), dst1=View(t2), n_split=32, pad=(1, 1))
                else:
                    ol.conv(P[p + ".b1"], View(t1, 0, 32), View(cat, 32, 32), pad=(1, 1))
                    ol.conv(P[p + ".b2a"], View(t1, 32, 32), View(t2), pad=(1, 1))
                ol.conv(P[p + ".b2b"], View(t2), View(cat, 64, 32), pad=(1, 1))
                ol.conv(P[p + ".out"], View(cat), View(xs), residual=View(xs), relu=True)
        # ---- Mixed_6a
        h6, w6 = _out_hw(h5, 3, 2), _out_hw(w5, 3, 2)
        x17 = buf(h6, w6, 896)
        t6a, t6b = buf(h5, w5, 192), buf(h5, w5, 192)
        ol.conv(P["m6a.b0"], View(x35), View(x17, 0, 384), stride=2)
        ol.conv(P["m6a.b1a"], View(x35), View(t6a))
        ol.conv(P["m6a.b1b"], View(t6a), View(t6b), pad=(1, 1))
        ol.conv(P["m6a.b1c"], View(t6b), View(x17, 384, 256), stride=2)
        ol.maxpool(View(x35), View(x17, 640, 256))
        # ---- 10 x Block17 (scale 0.10)
        cat17, t17a, t17b = buf(h6, w6, 256), buf(h6, w6, 128), buf(h6, w6, 128)
        fused17 = FUSED_B17 and (h6, w6) == (8, 8)
        for i in range(10):
            p = "repeat_2.%d" % i
            if fused17:
                # the whole block in one persistent kernel: branch activations never leave shared memory / TMEM
                ol.block17(P[p + ".in"], P[p + ".b1a"], P[p + ".b1b"], P[p + ".out"], x17)
                continue
            ol.conv(P[p + ".in"], View(x17), View(cat17, 0, 128), dst1=View(t17a), n_split=128)
            ol.conv(P[p + ".b1a"], View(t17a), View(t17b), pad=(0, 3))
            ol.conv(P[p + ".b1b"], View(t17b), View(cat17, 128, 128), pad=(3, 0))
            ol.conv(P[p + ".out"], View(cat17), View(x17), residual=View(x17), relu=True)
        # ---- Mixed_7a
        h7, w7 = _out_hw(h6, 3, 2), _out_hw(w6, 3, 2)
        x8 = buf(h7, w7, 1792)
        t7, t7b = buf(h6, w6, 768), buf(h6, w6, 256)
        ol.conv(P["m7a.in"], View(x17), View(t7))
        ol.conv(P["m7a.b0"], View(t7, 0, 256), View(x8, 0, 384), stride=2)
        ol.conv(P["m7a.b1"], View(t7, 256, 256), View(x8, 384, 256), stride=2)
        ol.conv(P["m7a.b2a"], View(t7, 512, 256), View(t7b), pad=(1, 1))
        ol.conv(P["m7a.b2b"], View(t7b), View(x8, 640, 256), stride=2)
        ol.maxpool(View(x17), View(x8, 896, 896))
        # ---- 5 x Block8 (scale 0.20) + block8 (scale 1, no ReLU)
        cat8, t8a, t8b = buf(h7, w7, 384), buf(h7, w7, 192), buf(h7, w7, 192)
        for i in list(range(5)) + [None]:
            p = "repeat_3.%d" % i if i is not None else "block8"
            ol.conv(P[p + ".in"], View(x8), View(cat8, 0, 192), dst1=View(t8a), n_split=192)
            ol.conv(P[p + ".b1a"], View(t8a), View(t8b), pad=(0, 1))
            ol.conv(P[p + ".b1b"], View(t8b), View(cat8, 192, 192), pad=(1, 0))
            ol.conv(P[p + ".out"], View(cat8), View(x8), residual=View(x8), relu=(i is not None))
        # ---- avgpool -> last_linear + last_bn -> ...: the fused tail kernel reads x8 (tail.TailPlan, in_mode 0)
        self.x8 = x8
        self.weights = weights
        self.tails = {}
        self.taps = {"conv2d_1a": c1a, "conv2d_2b": c2b, "conv2d_4b_repeat_1": x35, "repeat_2": x17, "block8": x8}

    def run(self):
        self.ol.run()
