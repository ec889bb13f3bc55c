"""Drop-in for the reference's ``models.MTCNN`` (models/mtcnn.py:160-518) on hand-written sm_100a kernels.

Same constructor keywords (``MTCNN(**cfg/detection/mtcnn.json)`` works), same ``.pnet/.rnet/.onet`` ``state_dict`` keys
(the bundled weights are auto-loaded from ``weights_mtcnn/``, mtcnn.py:32-36, :78-82, :132-136), same methods and
return conventions: ``detect`` / ``inference`` return host numpy ``(boxes, probs[, points])``, ``forward`` returns
``(faces, boxes[, probs])``, ``select_boxes`` and ``extract`` keep their semantics.

What runs underneath (include/vnfr_b200.h) -- no stage leaves the GPU, nothing synchronises until the final read-back:
  pyramid_resize_norm -> pnet_sweep_compact -> stage1_boxes -> rnet_forward -> stage2_boxes -> onet_forward ->
  stage3_faces [-> face_crops]
The nn.Module tree only HOLDS parameters; there is no CPU fallback (a missing CUDA device / library raises).
"""
import ctypes as C
import os

import numpy as np
import torch
from torch import nn

from .. import _lib

_WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights_mtcnn")


# ----------------------------------------------------------------------------------------------------------------
# parameter holders with the reference's names (mtcnn.py:9-157)
# ----------------------------------------------------------------------------------------------------------------
class PNet(nn.Module):
    def __init__(self, pretrained=True):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 10, kernel_size=3); self.prelu1 = nn.PReLU(10)
        self.conv2 = nn.Conv2d(10, 16, kernel_size=3); self.prelu2 = nn.PReLU(16)
        self.conv3 = nn.Conv2d(16, 32, kernel_size=3); self.prelu3 = nn.PReLU(32)
        self.conv4_1 = nn.Conv2d(32, 2, kernel_size=1)
        self.conv4_2 = nn.Conv2d(32, 4, kernel_size=1)
        self.training = False
        if pretrained:
            self.load_state_dict(torch.load(os.path.join(_WEIGHTS_DIR, "pnet.pt"), map_location="cpu"))


class RNet(nn.Module):
    def __init__(self, pretrained=True):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 28, kernel_size=3); self.prelu1 = nn.PReLU(28)
        self.conv2 = nn.Conv2d(28, 48, kernel_size=3); self.prelu2 = nn.PReLU(48)
        self.conv3 = nn.Conv2d(48, 64, kernel_size=2); self.prelu3 = nn.PReLU(64)
        self.dense4 = nn.Linear(576, 128); self.prelu4 = nn.PReLU(128)
        self.dense5_1 = nn.Linear(128, 2)
        self.dense5_2 = nn.Linear(128, 4)
        self.training = False
        if pretrained:
            self.load_state_dict(torch.load(os.path.join(_WEIGHTS_DIR, "rnet.pt"), map_location="cpu"))


class ONet(nn.Module):
    def __init__(self, pretrained=True):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, kernel_size=3); self.prelu1 = nn.PReLU(32)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3); self.prelu2 = nn.PReLU(64)
        self.conv3 = nn.Conv2d(64, 64, kernel_size=3); self.prelu3 = nn.PReLU(64)
        self.conv4 = nn.Conv2d(64, 128, kernel_size=2); self.prelu4 = nn.PReLU(128)
        self.dense5 = nn.Linear(1152, 256); self.prelu5 = nn.PReLU(256)
        self.dense6_1 = nn.Linear(256, 2)
        self.dense6_2 = nn.Linear(256, 4)
        self.dense6_3 = nn.Linear(256, 10)
        self.training = False
        if pretrained:
            self.load_state_dict(torch.load(os.path.join(_WEIGHTS_DIR, "onet.pt"), map_location="cpu"))


# ----------------------------------------------------------------------------------------------------------------
# weight packing (layouts consumed by csrc/detect_pnet.cu and csrc/detect_heads.cu)
# ----------------------------------------------------------------------------------------------------------------
def _pack_pnet(sd):
    """torch layouts [co][ci][ky][kx], flattened in the order csrc/detect_pnet.cu expects (6 632 floats)."""
    order = ["conv1.weight", "conv1.bias", "prelu1.weight", "conv2.weight", "conv2.bias", "prelu2.weight", "conv3.weight",
             "conv3.bias", "prelu3.weight", "conv4_1.weight", "conv4_1.bias", "conv4_2.weight", "conv4_2.bias"]
    return torch.cat([sd[k].detach().float().cpu().reshape(-1) for k in order]).contiguous()


def _kc(w):
    """conv weight (co,ci,kh,kw) -> [K = ci*kh*kw][co]"""
    return w.detach().float().cpu().reshape(w.shape[0], -1).t().contiguous().reshape(-1)


def _fc_whc_to_chw(w, c, h, wd):
    """dense weight (out, W*H*C) indexing the reference's (W,H,C) flatten (mtcnn.py:93-94, :150-151) -> [K = c*h*w][out]
    with K in (C,H,W) order, the layout the conv output has in shared memory."""
    out = w.shape[0]
    w4 = w.detach().float().cpu().reshape(out, wd, h, c).permute(0, 3, 2, 1)      # out, c, h, w
    return w4.reshape(out, c * h * wd).t().contiguous().reshape(-1)


def _pack_rnet(sd):
    f = lambda k: sd[k].detach().float().cpu().reshape(-1)
    heads_w = torch.zeros(128, 8)
    heads_w[:, 0:2] = sd["dense5_1.weight"].detach().float().cpu().t()
    heads_w[:, 2:6] = sd["dense5_2.weight"].detach().float().cpu().t()
    heads_b = torch.zeros(8)
    heads_b[0:2] = f("dense5_1.bias"); heads_b[2:6] = f("dense5_2.bias")
    parts = [_kc(sd["conv1.weight"]), f("conv1.bias"), f("prelu1.weight"),
             _kc(sd["conv2.weight"]), f("conv2.bias"), f("prelu2.weight"),
             _kc(sd["conv3.weight"]), f("conv3.bias"), f("prelu3.weight"),
             _fc_whc_to_chw(sd["dense4.weight"], 64, 3, 3), f("dense4.bias"), f("prelu4.weight"),
             heads_w.reshape(-1), heads_b]
    return torch.cat(parts).contiguous()


def _pack_onet(sd):
    f = lambda k: sd[k].detach().float().cpu().reshape(-1)
    heads_w = torch.zeros(256, 16)
    heads_w[:, 0:2] = sd["dense6_1.weight"].detach().float().cpu().t()
    heads_w[:, 2:6] = sd["dense6_2.weight"].detach().float().cpu().t()
    heads_w[:, 6:16] = sd["dense6_3.weight"].detach().float().cpu().t()
    heads_b = torch.cat([f("dense6_1.bias"), f("dense6_2.bias"), f("dense6_3.bias")])
    parts = [_kc(sd["conv1.weight"]), f("conv1.bias"), f("prelu1.weight"),
             _kc(sd["conv2.weight"]), f("conv2.bias"), f("prelu2.weight"),
             _kc(sd["conv3.weight"]), f("conv3.bias"), f("prelu3.weight"),
             _kc(sd["conv4.weight"]), f("conv4.bias"), f("prelu4.weight"),
             _fc_whc_to_chw(sd["dense5.weight"], 128, 3, 3), f("dense5.bias"), f("prelu5.weight"),
             heads_w.reshape(-1), heads_b]
    return torch.cat(parts).contiguous()


class HeadsBackWeights:
    """Split-precision weights (VnfrHeadsBack, include/vnfr_b200.h) of the layers after the last tensor-core convolution:
    R-Net conv3 / dense4 / dense5_* (mtcnn.py:84-99), O-Net conv4 / dense5 / dense6_* (mtcnn.py:138-157)."""

    def __init__(self, sd, onet, dev):
        from ..tail import SplitLinear
        f = lambda k: sd[k].detach().float().cpu()
        conv, fc = ("conv4", "dense5") if onet else ("conv3", "dense4")
        heads = ["dense6_1", "dense6_2", "dense6_3"] if onet else ["dense5_1", "dense5_2"]
        cw = f(conv + ".weight")                                   # (co, ci, 2, 2)
        co, ci = cw.shape[:2]
        w1 = torch.zeros(co, 4, 64)
        w1[:, :, :ci] = cw.permute(0, 2, 3, 1).reshape(co, 4, ci)   # column (ky*2 + kx)*64 + c
        self.layers = [SplitLinear(w1.reshape(co, 256), f(conv + ".bias"), dev),
                       SplitLinear(f(fc + ".weight"), f(fc + ".bias"), dev),
                       SplitLinear(torch.cat([f(h + ".weight") for h in heads]), torch.cat([f(h + ".bias") for h in heads]), dev)]
        self.alpha = []
        for name, lin in (("prelu" + conv[-1], self.layers[0]), ("prelu" + fc[-1], self.layers[1])):
            a = torch.zeros(lin.N_pad, dtype=torch.float32, device=dev)
            a[:lin.N] = f(name + ".weight").to(dev)
            self.alpha.append(a)
        self.onet = bool(onet)

    def struct(self, planes):
        hb = _lib.HeadsBack()
        for l in range(3):
            hb.w[l] = self.layers[l].w.data_ptr()
            hb.bias[l] = self.layers[l].bias.data_ptr()
        for l in range(2):
            hb.alpha[l] = self.alpha[l].data_ptr()
        hb.planes = planes.data_ptr()
        return hb


def heads_back_planes(onet, crop_cap, dev):
    """Workspace of the tensor-core back half for ``crop_cap`` crops (1024-byte aligned: torch allocations are 512-byte
    aligned, so over-allocate and slice)."""
    n = int(_lib.lib().vnfr_heads_back_workspace_bytes(1 if onet else 0, int(crop_cap)))
    buf = torch.empty(n + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + n]


class DetectWorkspace:
    """Device buffers of one detection pass for a fixed (B, H, W, min_face_size, factor, caps)."""

    def __init__(self, B, H, W, minsize, factor, caps, dev, crop_ws=(2048, 256), crop_floor=(2048, 512)):
        self.pyr = _lib.Pyramid()
        _lib.call("vnfr_pyramid_plan", B, H, W, int(minsize), float(factor), C.byref(self.pyr))
        p = self.pyr
        L = p.n_levels
        self.B, self.H, self.W, self.L = B, H, W, L
        cap1, cap2, cap3, capf = caps
        self.caps = caps
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        nseg = max(B * L, 1)
        self.levels = torch.empty(max(p.level_off[L], 1), **f32)
        # every counter lives in ONE int32 tensor so a single memset resets a pass
        self.counters = torch.zeros(nseg + nseg + 3 * B + 1, **i32)
        o = 0
        self.cand_count = self.counters[o:o + nseg]; o += nseg
        self.keep1_count = self.counters[o:o + nseg]; o += nseg
        self.s2_count = self.counters[o:o + B]; o += B
        self.s3_count = self.counters[o:o + B]; o += B
        self.out_count = self.counters[o:o + B]; o += B
        self.status = self.counters[o:o + 1]
        self.cand_cell = torch.empty(nseg * cap1, **i32)
        self.cand_score = torch.empty(nseg * cap1, **f32)
        self.cand_reg = torch.empty(nseg * cap1 * 4, **f32)
        self.keep1 = torch.empty(nseg * cap1, **i32)
        self.s2_box = torch.empty(B * cap2 * 4, **f32)
        self.s2_pad = torch.empty(B * cap2 * 4, **i32)
        self.s2_prob = torch.empty(B * cap2, **f32)
        self.s2_reg = torch.empty(B * cap2 * 4, **f32)
        self.s3_box = torch.empty(B * cap3 * 4, **f32)
        self.s3_pad = torch.empty(B * cap3 * 4, **i32)
        self.s3_prob = torch.empty(B * cap3, **f32)
        self.s3_reg = torch.empty(B * cap3 * 4, **f32)
        self.s3_lmk = torch.empty(B * cap3 * 10, **f32)
        self.offs = torch.zeros(B + 1, **i32)
        # crop workspaces of the R-Net / O-Net stages (fp32 [n][3][S][S]); sized for CROP_WS_PER_FRAME candidates per frame
        # on average (overflow raises through the status word, like the caps)
        self.rcrop_cap = max(1, min(B * cap2, max(crop_floor[0], B * crop_ws[0])))
        self.ocrop_cap = max(1, min(B * cap3, max(crop_floor[1], B * crop_ws[1])))
        self.rcrops = torch.empty(self.rcrop_cap * 3 * 24 * 24, **f32)
        self.ocrops = torch.empty(self.ocrop_cap * 3 * 48 * 48, **f32)
        self.rp1 = self.rc2 = self.rback = None
        if MTCNN.rnet_tensor_cores:
            # R-Net conv2 on the tensor cores: pooled conv1 map as 2 fp16 parts, conv2 output in fp32 (vnfr_rnet_forward_tc)
            self.rp1 = torch.empty(self.rcrop_cap * 11 * 11 * 64, dtype=torch.float16, device=dev)
            self.rc2 = torch.empty(self.rcrop_cap * 9 * 9 * 48, **f32)
            self.rback = heads_back_planes(False, self.rcrop_cap, dev) if MTCNN.heads_back_tensor_cores else None
        if MTCNN.onet_tensor_cores:
            # O-Net conv2 on the tensor cores: pooled conv1 map as 2 fp16 / 3 bf16 parts, conv2 output in fp32 (vnfr_onet_forward_tc)
            if MTCNN.onet_split_mode == 2:
                self.op1 = torch.empty(self.ocrop_cap * 23 * 23 * 64, dtype=torch.float16, device=dev)
            else:
                self.op1 = torch.empty(self.ocrop_cap * 23 * 23 * 96, dtype=torch.bfloat16, device=dev)
            self.oc2 = torch.empty(self.ocrop_cap * 21 * 21 * 64, **f32)
            self.op3 = self.oc3 = self.oback = None
            if MTCNN.onet_split_mode == 2 and MTCNN.onet_conv3_tensor_cores:
                # conv3 on the tensor cores too: pooled conv2 map as 2 fp16 parts, conv3 output in fp32
                self.op3 = torch.empty(self.ocrop_cap * 10 * 10 * 128, dtype=torch.float16, device=dev)
                self.oc3 = torch.empty(self.ocrop_cap * 8 * 8 * 64, **f32)
                self.oback = heads_back_planes(True, self.ocrop_cap, dev) if MTCNN.heads_back_tensor_cores else None
        self.out_box = torch.zeros(B, capf, 5, **f32)
        self.out_pts = torch.zeros(B, capf, 10, **f32)


class CropWorkspaceOverflow(_lib.VnfrError):
    """More R-/O-Net candidates than the crop workspaces hold: callers grow the workspace and repeat the pass."""


class ResultWorkspace:
    """Detections of a whole batch assembled from sub-batch passes (the host-frame path overlaps the H2D copy of
    sub-batch i+1 with the cascade of sub-batch i): just the tensors the face-crop stage and the read-back need."""

    def __init__(self, B, H, W, caps, dev):
        self.B, self.H, self.W, self.caps = B, H, W, caps
        capf = caps[3]
        self.counters = torch.zeros(B + 1, dtype=torch.int32, device=dev)      # out_count (B) + status
        self.out_count = self.counters[:B]
        self.status = self.counters[B:]
        self.out_box = torch.zeros(B, capf, 5, dtype=torch.float32, device=dev)
        self.out_pts = torch.zeros(B, capf, 10, dtype=torch.float32, device=dev)
        self.offs = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        self.frames = None


class MTCNN(nn.Module):
    """MTCNN face detection module -- see the module docstring; keyword arguments as in mtcnn.py:200-204."""

    #: per-stage capacities (candidates per (image, level) / boxes per image into R-Net / into O-Net / faces per image).
    #: Exceeding one raises VnfrError (results would otherwise be silently truncated; the reference has no such limit).
    #: The first three are bounded by CAPS_CEILING: every NMS runs inside ONE CTA with sort keys, boxes and keep list in
    #: shared memory (36 B per entry: 4 096 entries = 144 KB of the SM's 227 KB).  Measured headroom: the crowded 4K /
    #: min_face_size 20 config peaks at 979 R-Net boxes per frame and ~1.3 k P-Net candidates per (image, level).  A frame
    #: beyond the ceiling (tens of thousands of face-like patches) must be tiled by the caller.
    caps = (4096, 4096, 2048, 256)
    CAPS_CEILING = (4096, 4096, 4096, None)
    #: average R-Net / O-Net candidates per frame the crop workspaces are sized for (6.9 KB / 27.6 KB per crop)
    crop_ws_per_frame = (2048, 256)
    #: minimum size (crops) of the two workspaces whatever the batch
    crop_ws_floor = (2048, 512)
    #: run O-Net's conv2 on the tensor cores in split precision (fp32-level accuracy); VNFR_ONET_FMA=1 keeps it on the FMA pipe
    onet_tensor_cores = not os.environ.get("VNFR_ONET_FMA")
    #: R-Net's conv2 on the tensor cores (two fp16 parts, three products); VNFR_RNET_FMA=1 keeps it on the FMA pipe
    rnet_tensor_cores = not os.environ.get("VNFR_RNET_FMA")
    #: split of the fp32 operands of that convolution: 2 = two fp16 parts, three products (default); 1 = three bf16 parts, six
    #: products (VNFR_ONET_SPLIT=1)
    onet_split_mode = int(os.environ.get("VNFR_ONET_SPLIT", "2"))
    #: O-Net conv3 on the tensor cores as well (needs onet_split_mode 2); VNFR_ONET_CONV3_FMA=1 keeps it on the FMA pipe
    onet_conv3_tensor_cores = not os.environ.get("VNFR_ONET_CONV3_FMA")
    #: the layers after the last tensor-core convolution (2x2 conv, dense layer, heads) as split-precision GEMMs over all crops
    #: (csrc/heads_chain.cu); VNFR_HEADS_BACK_FMA=1 keeps the per-crop FMA kernels
    heads_back_tensor_cores = not os.environ.get("VNFR_HEADS_BACK_FMA")

    def __init__(self, image_size=160, margin=0, min_face_size=20, thresholds=[0.6, 0.7, 0.7], factor=0.709,
                 post_process=True, select_largest=True, selection_method=None, keep_all=False, device=None):
        super().__init__()
        self.image_size = image_size
        self.margin = margin
        self.min_face_size = min_face_size
        self.thresholds = thresholds
        self.factor = factor
        self.post_process = post_process
        self.select_largest = select_largest
        self.keep_all = keep_all
        self.selection_method = selection_method

        self.pnet = PNet()
        self.rnet = RNet()
        self.onet = ONet()

        self._packed = None
        self._ws = {}
        self.device = torch.device("cpu")
        if device is not None:
            self.device = torch.device(device) if isinstance(device, str) else device
            self.to(device)
        if not self.selection_method:
            self.selection_method = "largest" if self.select_largest else "probability"

    # ---- weights ------------------------------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._packed = None
        return r

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._packed = None
        return r

    def _cuda_device(self):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise _lib.VnfrError("MTCNN(device=%r): this package only runs on CUDA devices (no CPU path)" % (self.device,))
        if not torch.cuda.is_available():
            raise _lib.VnfrError("CUDA is not available: vn_celeb_face_recognition_b200 has no CPU fallback")
        return dev

    def _weights(self, dev):
        if self._packed is None or self._packed["dev"] != dev:
            pw = _pack_pnet(self.pnet.state_dict())
            rw = _pack_rnet(self.rnet.state_dict())
            ow = _pack_onet(self.onet.state_dict())
            assert rw.numel() == _lib.lib().vnfr_rnet_weight_floats() and ow.numel() == _lib.lib().vnfr_onet_weight_floats()
            from .. import encoder_plan
            osd = self.onet.state_dict()
            pack = encoder_plan.pack_conv_split2 if MTCNN.onet_split_mode == 2 else encoder_plan.pack_conv_split3
            w2s = pack(osd["conv2.weight"], osd["conv2.bias"], dev, 32)
            w3s = encoder_plan.pack_conv_split2(osd["conv3.weight"], osd["conv3.bias"], dev, 64)
            rsd = self.rnet.state_dict()
            rw2s = encoder_plan.pack_conv_split2(rsd["conv2.weight"], rsd["conv2.bias"], dev, 32)
            self._packed = {"dev": dev, "pnet": _lib.pack_pnet_weights(pw, dev), "rnet": rw.to(dev), "onet": ow.to(dev),
                            "onet_w2s": w2s.w, "onet_w3s": w3s.w, "rnet_w2s": rw2s.w,
                            "rnet_back": HeadsBackWeights(rsd, False, dev), "onet_back": HeadsBackWeights(osd, True, dev)}
        return self._packed

    # ---- the device pipeline ------------------------------------------------------------------------------------
    def detect_device(self, frames_u8, select_largest=None, mark=None, slot=0):
        with torch.cuda.device(frames_u8.device):      # the C library launches on the CURRENT device
            return self._detect_device(frames_u8, select_largest, mark, slot)

    def _detect_device(self, frames_u8, select_largest=None, mark=None, slot=0):
        """frames_u8: CUDA uint8 (B,H,W,3) RGB.  Runs the whole three-stage cascade on the current stream and returns
        the DetectWorkspace holding out_count (B,), out_box (B,capf,5), out_pts (B,capf,10), status -- all on device,
        nothing synchronised."""
        assert frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[3] == 3
        frames_u8 = frames_u8.contiguous()
        dev = frames_u8.device
        B, H, W, _ = frames_u8.shape
        wts = self._weights(dev)
        key = (B, H, W, self.min_face_size, self.factor, tuple(self.caps), dev, slot)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) > 12:
                self._ws.clear()
            ws = self._ws[key] = DetectWorkspace(B, H, W, self.min_face_size, self.factor, tuple(self.caps), dev,
                                                 tuple(self.crop_ws_per_frame), tuple(self.crop_ws_floor))
        cap1, cap2, cap3, capf = ws.caps
        st = _lib.stream_ptr()
        P = _lib.ptr
        t0, t1, t2 = [float(t) for t in self.thresholds]
        sl = self.select_largest if select_largest is None else select_largest
        mark = mark or (lambda name: None)        # optional stage markers (bench.py records CUDA events here)
        ws.counters.zero_()
        mark("start")
        if ws.L > 0:
            _lib.call("vnfr_pyramid_resize_norm", C.byref(ws.pyr), P(frames_u8), P(ws.levels), st)
            mark("pyramid")
            _lib.call("vnfr_pnet_sweep_compact", C.byref(ws.pyr), P(ws.levels), P(wts["pnet"]), t0, cap1, P(ws.cand_count), P(ws.cand_cell),
                      P(ws.cand_score), P(ws.cand_reg), None, None, st)
            mark("pnet")
        _lib.call("vnfr_stage1_boxes", C.byref(ws.pyr), cap1, P(ws.cand_count), P(ws.cand_cell), P(ws.cand_score),
                  P(ws.cand_reg), P(ws.keep1_count), P(ws.keep1), cap2, P(ws.s2_count), P(ws.s2_box), P(ws.s2_pad),
                  P(ws.status), st)
        mark("stage1_nms")
        if ws.rp1 is not None:
            _lib.call("vnfr_rnet_forward_tc", P(frames_u8), B, H, W, cap2, P(ws.s2_count), P(ws.s2_pad), P(wts["rnet"]),
                      P(wts["rnet_w2s"]), P(ws.s2_prob), P(ws.s2_reg), P(ws.offs), P(ws.rcrops), P(ws.rp1), P(ws.rc2),
                      ws.rcrop_cap, P(ws.status), C.byref(wts["rnet_back"].struct(ws.rback)) if ws.rback is not None else None, st)
        else:
            _lib.call("vnfr_rnet_forward", P(frames_u8), B, H, W, cap2, P(ws.s2_count), P(ws.s2_pad), P(wts["rnet"]),
                      P(ws.s2_prob), P(ws.s2_reg), P(ws.offs), P(ws.rcrops), ws.rcrop_cap, P(ws.status), st)
        mark("rnet")
        _lib.call("vnfr_stage2_boxes", B, H, W, cap2, P(ws.s2_count), P(ws.s2_box), P(ws.s2_prob), P(ws.s2_reg), t1, cap3,
                  P(ws.s3_count), P(ws.s3_box), P(ws.s3_pad), P(ws.status), st)
        mark("stage2_nms")
        if hasattr(ws, "op1"):
            _lib.call("vnfr_onet_forward_tc", P(frames_u8), B, H, W, cap3, P(ws.s3_count), P(ws.s3_pad), P(wts["onet"]),
                      P(wts["onet_w2s"]), MTCNN.onet_split_mode, P(ws.s3_prob), P(ws.s3_reg), P(ws.s3_lmk), P(ws.offs), P(ws.ocrops), P(ws.op1),
                      P(ws.oc2), P(wts["onet_w3s"] if ws.op3 is not None else None), P(ws.op3), P(ws.oc3), ws.ocrop_cap,
                      P(ws.status), C.byref(wts["onet_back"].struct(ws.oback)) if ws.oback is not None else None, st)
        else:
            _lib.call("vnfr_onet_forward", P(frames_u8), B, H, W, cap3, P(ws.s3_count), P(ws.s3_pad), P(wts["onet"]),
                      P(ws.s3_prob), P(ws.s3_reg), P(ws.s3_lmk), P(ws.offs), P(ws.ocrops), ws.ocrop_cap, P(ws.status), st)
        mark("onet")
        _lib.call("vnfr_stage3_faces", B, cap3, P(ws.s3_count), P(ws.s3_box), P(ws.s3_prob), P(ws.s3_reg), P(ws.s3_lmk), t2,
                  1 if sl else 0, capf, P(ws.out_count), P(ws.out_box), P(ws.out_pts), P(ws.status), st)
        mark("stage3_nms")
        ws.frames = frames_u8
        return ws

    def detect_device_chunked(self, frames_dev, ready_events, bounds, slot=0, wait_current=True):
        """The cascade over the sub-batches ``bounds`` = [(b0, b1), ...] of ``frames_dev`` (CUDA u8 (B,H,W,3)); sub-batch
        i starts as soon as ``ready_events[i]`` (recorded on the copy stream after its H2D) has fired.  Returns a
        ResultWorkspace with the detections of the whole batch.

        ``slot`` selects one of two ResultWorkspaces and ``wait_current=False`` drops the dependency of the cascade on the
        work already enqueued on the current stream (only ``ready_events`` gate it): together they let the cascade of
        batch i+1 run while the encoder of batch i is still busy on the current stream (FacePipeline pipelining).  The
        caller then guarantees that the previous user of this slot has finished (its face crops have been read)."""
        B, H, W, _ = frames_dev.shape
        dev = frames_dev.device
        key = ("result", B, H, W, tuple(self.caps), dev, slot)
        full = self._ws.get(key)
        if full is None:
            full = self._ws[key] = ResultWorkspace(B, H, W, tuple(self.caps), dev)
        cur = torch.cuda.current_stream(dev)
        # consecutive sub-batches run on two alternating streams (each with its own workspace): the low-occupancy stage
        # kernels of one sub-batch (one CTA per image NMS, persistent-kernel tails) overlap the P-Net / R-Net of the next
        if getattr(self, "_chunk_streams", None) is None or self._chunk_streams[0].device != dev:
            self._chunk_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s0, s1 = self._chunk_streams
        if wait_current:
            s0.wait_stream(cur)
            s1.wait_stream(cur)
        with torch.cuda.stream(s0):
            if ready_events is not None and not wait_current:
                s0.wait_event(ready_events[0])            # orders the reset after the previous user of this slot
            full.status.zero_()
            zeroed = torch.cuda.Event()
            zeroed.record(s0)
        s1.wait_event(zeroed)
        for i, (b0, b1) in enumerate(bounds):
            st = self._chunk_streams[i & 1]
            with torch.cuda.stream(st):
                if ready_events is not None:
                    st.wait_event(ready_events[i])
                ws = self.detect_device(frames_dev[b0:b1], slot=i & 1)
                full.out_count[b0:b1].copy_(ws.out_count)
                full.out_box[b0:b1].copy_(ws.out_box)
                full.out_pts[b0:b1].copy_(ws.out_pts)
                full.status.bitwise_or_(ws.status)
        for s_ in self._chunk_streams:
            cur.wait_stream(s_)
        full.frames = frames_dev
        return full

    def grow_crop_workspace(self):
        """Doubles the per-frame sizing of the R-/O-Net crop workspaces (called after a crop-workspace overflow) and drops
        the cached workspaces so that the next pass allocates the larger ones."""
        self.crop_ws_per_frame = tuple(2 * v for v in self.crop_ws_per_frame)
        self._ws.clear()

    @staticmethod
    def check_status(status):
        if status & 32 and not (status & 31):
            raise CropWorkspaceOverflow("R-/O-Net crop workspace exceeded (MTCNN.crop_ws_per_frame)")
        if status:
            names = ["cap1 (P-Net candidates per image/level)", "cap2 (boxes per image into R-Net)",
                     "cap3 (boxes per image into O-Net)", "capf (faces per image)", "max_faces",
                     "crop workspace (MTCNN.crop_ws_per_frame)"]
            over = [n for i, n in enumerate(names) if status & (1 << i)]
            raise _lib.VnfrError("detection capacity exceeded: %s.  MTCNN.caps = %s can be raised up to MTCNN.CAPS_CEILING = %s (the "
                                 "per-CTA shared-memory NMS holds 4096 boxes; capf / max_faces_per_frame have no ceiling); a frame "
                                 "beyond that must be split into tiles by the caller" % (", ".join(over), MTCNN.caps, MTCNN.CAPS_CEILING))

    def face_crops_device(self, ws, mode, image_size, margin=0, template=None, half_dtype=None, max_faces=None,
                          want_u8=True):
        """Faces of the detections in ``ws`` as encoder inputs (see vnfr_face_crops).  Returns (face_u8 (F,S,S,3),
        face_half (F,ceil(S/2),ceil(S/2),16) -- the space-to-depth layout the encoder's first convolution reads --,
        face_img (F,), F_capacity) on device; the number of valid faces is ws.out_count.sum()."""
        from .. import encoder_plan
        dt = half_dtype or encoder_plan.HALF
        dev = ws.frames.device
        capf = ws.caps[3]
        max_faces = max_faces or ws.B * min(capf, 32)
        S = image_size
        key = ("faces", max_faces, S, dt)
        bufs = getattr(ws, "_face_bufs", {})
        if key not in bufs:
            bufs[key] = (torch.empty(max_faces, S, S, 3, dtype=torch.uint8, device=dev),
                         torch.zeros(max_faces, (S + 1) // 2, (S + 1) // 2, 16, dtype=dt, device=dev),
                         torch.empty(max_faces, dtype=torch.int32, device=dev))
            ws._face_bufs = bufs
        u8, half, fimg = bufs[key]
        tmpl = None
        if template is not None:
            tmpl = (C.c_float * 10)(*[float(v) for v in np.asarray(template, dtype=np.float32).reshape(-1)])
        with torch.cuda.device(dev):
            _lib.call("vnfr_face_crops", _lib.ptr(ws.frames), ws.B, ws.H, ws.W, capf, _lib.ptr(ws.out_count), _lib.ptr(ws.out_box),
                      _lib.ptr(ws.out_pts), mode, S, margin, tmpl, encoder_plan.dtype_code(dt), max_faces, _lib.ptr(ws.offs),
                      _lib.ptr(u8 if want_u8 else None), _lib.ptr(half), _lib.ptr(fimg), _lib.ptr(ws.status), 1, _lib.stream_ptr())
        return (u8 if want_u8 else None), half, fimg, max_faces

    # ---- reference API ------------------------------------------------------------------------------------------
    def _to_frames(self, img):
        """detect_face.py:26-41: PIL | ndarray (3-D / 4-D) | Tensor | list of equal-size images -> CUDA u8 (B,H,W,3)."""
        dev = self._cuda_device()
        if isinstance(img, (np.ndarray, torch.Tensor)):
            t = torch.as_tensor(img.copy() if isinstance(img, np.ndarray) else img)
            if t.dim() == 3:
                t = t.unsqueeze(0)
        else:
            if not isinstance(img, (list, tuple)):
                img = [img]
            if any(get_size(im) != get_size(img[0]) for im in img):
                raise Exception("MTCNN batch processing only compatible with equal-dimension images.")
            t = torch.as_tensor(np.stack([np.uint8(im) for im in img]))
        if t.dtype != torch.uint8:
            t = t.to(torch.uint8)
        return t.to(dev, non_blocking=True)

    @staticmethod
    def _is_batch(img):
        return (isinstance(img, (list, tuple)) or (isinstance(img, np.ndarray) and len(img.shape) == 4) or
                (isinstance(img, torch.Tensor) and len(img.shape) == 4))

    def detect(self, img, landmarks=False):
        """mtcnn.py:278-361.  Returns host numpy (boxes, probs[, points]); ragged batches come back as object arrays."""
        with torch.no_grad():
            frames = self._to_frames(img)
            for attempt in range(6):
                ws = self.detect_device(frames)
                cnt = ws.out_count.cpu().numpy()
                try:
                    self.check_status(int(ws.status.item()))
                    break
                except CropWorkspaceOverflow:
                    if attempt == 5:
                        raise
                    self.grow_crop_workspace()
            nmax = int(cnt.max()) if len(cnt) else 0
            box = ws.out_box[:, :max(nmax, 1)].cpu().numpy()
            pts = ws.out_pts[:, :max(nmax, 1)].cpu().numpy()
        boxes, probs, points = [], [], []
        for b in range(len(cnt)):
            n = int(cnt[b])
            if n == 0:
                boxes.append([]); probs.append([]); points.append([])
            else:
                boxes.append(box[b, :n, :4].copy()); probs.append(box[b, :n, 4].copy())
                points.append(pts[b, :n].reshape(n, 5, 2).copy())
        boxes, probs, points = _np_array(boxes), _np_array(probs), _np_array(points)
        if not self._is_batch(img):
            boxes, probs, points = boxes[0], probs[0], points[0]
        if landmarks:
            return boxes, probs, points
        return boxes, probs

    def inference(self, rgb_image, landmark=True):
        """mtcnn.py:511-513."""
        return self.detect(rgb_image, landmark)

    def forward(self, img, save_path=None, return_prob=False, extract_face=True):
        """mtcnn.py:229-276 (this fork returns the boxes as well)."""
        batch_boxes, batch_probs, batch_points = self.detect(img, landmarks=True)
        if not self.keep_all:
            batch_boxes, batch_probs, batch_points = self.select_boxes(batch_boxes, batch_probs, batch_points, img,
                                                                       method=self.selection_method)
        faces = self.extract(img, batch_boxes, save_path) if extract_face else None
        if return_prob:
            return faces, batch_boxes, batch_probs
        return faces, batch_boxes

    def select_boxes(self, all_boxes, all_probs, all_points, imgs, method="probability", threshold=0.9, center_weight=2.0):
        """mtcnn.py:363-456 (host numpy, trivial cost)."""
        batch_mode = True
        if not self._is_batch(imgs):
            imgs, all_boxes, all_probs, all_points = [imgs], [all_boxes], [all_probs], [all_points]
            batch_mode = False
        sel_b, sel_p, sel_pt = [], [], []
        for boxes, points, probs, img in zip(all_boxes, all_points, all_probs, imgs):
            boxes, probs, points = np.array(boxes), np.array(probs), np.array(points)
            if len(boxes) == 0:
                sel_b.append(None); sel_p.append([None]); sel_pt.append(None)
                continue
            elif method == "largest":
                order = np.argsort((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))[::-1]
            elif method == "probability":
                order = np.argsort(probs)[::-1]
            elif method == "center_weighted_size":
                sizes = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
                w, h = get_size(img)
                centers = np.array(list(zip((boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2)))
                off2 = np.sum(np.power(centers - (w / 2, h / 2), 2.0), 1)
                order = np.argsort(sizes - off2 * center_weight)[::-1]
            elif method == "largest_over_threshold":
                mask = probs > threshold
                boxes = boxes[mask]
                order = np.argsort((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))[::-1]
                if sum(mask) == 0:
                    sel_b.append(None); sel_p.append([None]); sel_pt.append(None)
                    continue
            sel_b.append(boxes[order][[0]]); sel_p.append(probs[order][[0]]); sel_pt.append(points[order][[0]])
        if batch_mode:
            return _np_array(sel_b), _np_array(sel_p), _np_array(sel_pt)
        return sel_b[0], sel_p[0][0], sel_pt[0]

    @staticmethod
    def _crop_mode(img):
        """crop_resize (detect_face.py:309-325) picks its resampler by input type: torch.Tensor -> adaptive average +
        ``.byte()`` (vnfr_face_crops mode 0), numpy.ndarray -> cv2.resize(INTER_AREA) (mode 2), PIL.Image ->
        Image.resize(BILINEAR) (mode 3)."""
        items = list(img) if isinstance(img, (list, tuple)) else [img]
        modes = {0 if isinstance(x, torch.Tensor) else (2 if isinstance(x, np.ndarray) else 3) for x in items}
        if len(modes) != 1:
            raise _lib.VnfrError("MTCNN.extract: a batch must hold one image type (Tensor, ndarray or PIL)")
        return modes.pop()

    def extract(self, img, batch_boxes, save_path):
        """mtcnn.py:458-509.  Crops run on the GPU with the resampler the reference uses for the input's type (see
        ``_crop_mode``); like the reference's, the returned faces are float CPU tensors."""
        batch_mode = self._is_batch(img)
        crop_mode = self._crop_mode(img)
        frames = self._to_frames(img)
        if not batch_mode:
            batch_boxes = [batch_boxes]
        save_path = [save_path] if isinstance(save_path, str) else (save_path or [None] * frames.shape[0])
        B, H, W, _ = frames.shape
        per_img = []
        for bx in batch_boxes:
            if bx is None or len(bx) == 0:
                per_img.append(np.zeros((0, 4), np.float32))
            else:
                bx = np.asarray(bx, dtype=np.float32).reshape(-1, 4)
                per_img.append(bx if self.keep_all else bx[[0]])
        capf = max(1, max(len(b) for b in per_img))
        cnt = torch.tensor([len(b) for b in per_img], dtype=torch.int32)
        box = torch.zeros(B, capf, 5)
        for i, b in enumerate(per_img):
            box[i, :len(b), :4] = torch.from_numpy(b)
        dev = frames.device
        total = int(cnt.sum())
        faces_f = None
        if total:
            from .. import encoder_plan
            u8 = torch.empty(total, self.image_size, self.image_size, 3, dtype=torch.uint8, device=dev)
            half = torch.empty(total, self.image_size, self.image_size, 8, dtype=encoder_plan.HALF, device=dev)
            offs = torch.zeros(B + 1, dtype=torch.int32, device=dev)
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            d_cnt, d_box = cnt.to(dev), box.to(dev)        # keep the device copies alive across the launch
            with torch.cuda.device(dev):
                _lib.call("vnfr_face_crops", _lib.ptr(frames), B, H, W, capf, _lib.ptr(d_cnt), _lib.ptr(d_box), None, crop_mode,
                          self.image_size, self.margin, None, encoder_plan.dtype_code(encoder_plan.HALF), total, _lib.ptr(offs),
                          _lib.ptr(u8), _lib.ptr(half), None, _lib.ptr(status), 0, _lib.stream_ptr())
            u8 = u8.cpu()                                                  # the reference crops on the host: CPU tensors out
            faces_f = u8.permute(0, 3, 1, 2).float()                       # F.to_tensor(np.float32(face)), detect_face.py:376
            if self.post_process:
                faces_f = fixed_image_standardization(faces_f)
        out, o = [], 0
        for i, b in enumerate(per_img):
            if batch_boxes[i] is None:
                out.append(None)
                continue
            if len(b) == 0 and self.keep_all:
                raise RuntimeError("stack expects a non-empty TensorList")    # mtcnn.py:499-500 on zero faces
            f = faces_f[o:o + len(b)]
            o += len(b)
            if save_path[i] is not None:
                _save_faces(u8[o - len(b):o], save_path[i])
            out.append(f if self.keep_all else f[0])
        return out if batch_mode else out[0]


def _save_faces(u8, path):
    """extract_face save_path handling (mtcnn.py:485-494; detect_face.py:328-332, :372-374)."""
    import cv2
    os.makedirs(os.path.dirname(path) + "/", exist_ok=True)
    name, ext = os.path.splitext(path)
    for i, f in enumerate(u8.cpu().numpy()):
        cv2.imwrite(path if i == 0 else name + "_" + str(i + 1) + ext, cv2.cvtColor(f, cv2.COLOR_RGB2BGR))


def _np_array(lst):
    """np.array(list) with numpy < 1.24 semantics: ragged -> 1-D object array (mtcnn.py:345-347, detect_face.py:183)."""
    try:
        return np.array(lst)
    except ValueError:
        out = np.empty(len(lst), dtype=object)
        for i, o in enumerate(lst):
            out[i] = o
        return out


def get_size(img):
    """detect_face.py:335-339."""
    if isinstance(img, (np.ndarray, torch.Tensor)):
        return tuple(img.shape[1::-1])
    return img.size


def fixed_image_standardization(image_tensor):
    """mtcnn.py:516-518."""
    return (image_tensor - 127.5) / 128.0


def prewhiten(x):
    """mtcnn.py:521-526."""
    mean = x.mean()
    std = x.std()
    std_adj = std.clamp(min=1.0 / (float(x.numel()) ** 0.5))
    return (x - mean) / std_adj
