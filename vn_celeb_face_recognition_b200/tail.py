"""Host side of the fused tail kernel (csrc/tail_fused.cu, ``vnfr_tail_prepare`` / ``vnfr_tail_run``):

    AdaptiveAvgPool2d(1) -> last_linear + last_bn -> F.normalize [-> dense_1 + ReLU -> dense_2 -> log_softmax -> argmax]

(models/inception_resnet_v1.py:294-302, models/mlp_model.py:10-15, demo_image.py:113-137) as ONE cooperative launch whose
contractions run in split precision (fp32-level accuracy: the predicted label must equal the fp32 reference's).
``pack_linear_split`` builds the weight layout; ``TailPlan`` owns the scratch buffers of one chain of layers for a row
capacity ``n_pad`` and is re-pointed at its inputs / outputs per call (no allocation, no host sync).
"""
import ctypes as C

import torch

from . import _lib

ROWOP = {"identity": 0, "relu": 1, "l2norm": 2, "logsoftmax": 3}
BLOCK_N = 128


def _ceil(a, b):
    return (a + b - 1) // b * b


class SplitLinear:
    """fp32 Linear (w [N][K], bias [N] or None) as two fp16 planes: rows [0,N_pad) = fp16(w), rows [N_pad,2N_pad) =
    fp16(w - hi); bias fp32 [N_pad]."""

    def __init__(self, w, bias, device):
        w = w.detach().to(device=device, dtype=torch.float32)
        self.N, self.K_real = w.shape
        self.K = _ceil(self.K_real, 64)
        if self.K != self.K_real:
            w = torch.nn.functional.pad(w, (0, self.K - self.K_real))
        self.N_pad = _ceil(self.N, BLOCK_N)
        hi = w.to(torch.float16)
        lo = (w - hi.float()).to(torch.float16)
        self.w = torch.zeros(2 * self.N_pad, self.K, dtype=torch.float16, device=device)
        self.w[:self.N] = hi
        self.w[self.N_pad:self.N_pad + self.N] = lo
        self.bias = torch.zeros(self.N_pad, dtype=torch.float32, device=device)
        if bias is not None:
            self.bias[:self.N] = bias.detach().to(device=device, dtype=torch.float32)


def pick_split_k(m_tiles, n_tiles, kb, sms=148):
    """K ranges per output tile so that one wave of tiles covers the SMs (each range keeps >= 2 K blocks)."""
    s = max(1, sms // max(1, m_tiles * n_tiles))
    return max(1, min(s, kb // 2 if kb >= 2 else 1, 8))


class TailPlan:
    """One chain ``layers`` = [(SplitLinear, rowop name), ...] (<= 3) for up to ``n_pad`` rows.

    in_mode 0: input = 16-bit NHWC activations (n, hw, pitch) averaged over hw (the encoder's block8 output);
    in_mode 1: input = fp32 rows (n, K)."""

    def __init__(self, layers, n_pad, device, in_mode=0):
        assert 1 <= len(layers) <= 3
        n_pad = _ceil(max(int(n_pad), 1), 128)
        self.n_pad, self.layers, self.device = n_pad, layers, device
        op = _lib.TailOp()
        op.n_layers, op.n_pad, op.in_mode = len(layers), n_pad, in_mode
        self.bar = torch.zeros(1, dtype=torch.int32, device=device)
        op.grid_barrier = self.bar.data_ptr()
        self.a = [torch.zeros(2 * n_pad, lin.K, dtype=torch.float16, device=device) for lin, _ in layers]
        self.partial = []
        m_tiles = n_pad // 128
        for l, (lin, rowop) in enumerate(layers):
            L = op.layer[l]
            L.K, L.N, L.N_pad = lin.K, lin.N, lin.N_pad
            L.split_k = pick_split_k(m_tiles, lin.N_pad // BLOCK_N, lin.K // 64)
            L.rowop = ROWOP[rowop]
            part = torch.empty(L.split_k, n_pad, lin.N_pad, dtype=torch.float32, device=device)
            self.partial.append(part)
            L.weights, L.bias, L.partial = lin.w.data_ptr(), lin.bias.data_ptr(), part.data_ptr()
            L.a_in = self.a[l].data_ptr()
            L.a_next = self.a[l + 1].data_ptr() if l + 1 < len(layers) else None
            if l + 1 < len(layers):
                assert layers[l + 1][0].K == lin.N, "layer widths must chain (and be multiples of 64)"
        _lib.call("vnfr_tail_prepare", C.byref(op))
        self.op = op
        self._keep = None

    def run(self, n, x=None, x_f32=None, out_vecs=None, emb_half=None, label=None, prob=None, payload=None, thr=0.0,
            thr_class=None, n_classes=0, count_cell=None, count_value=0):
        """``x``: (n, hw, pitch) 16-bit activations (in_mode 0) / ``x_f32``: (n, K) fp32 (in_mode 1).
        ``out_vecs``: per-layer fp32 destination (2-D, unit inner stride) or None.  ``payload``: (n, D+2) fp32 view of the send-buffer
        rows of this call: label / prob go to its columns D / D+1 (the embedding gets there through out_vecs);
        ``count_cell`` (one fp32 element) receives ``count_value``."""
        op = self.op
        assert n <= self.n_pad
        if op.in_mode == 0:
            assert x.dim() == 3 and x.stride(2) == 1 and x.stride(0) == x.shape[1] * x.stride(1)
            op.x, op.hw, op.x_pitch = x.data_ptr(), x.shape[1], x.stride(1)
            op.x_dtype = 1 if x.dtype == torch.float16 else 0
        else:
            assert x_f32.dtype == torch.float32 and x_f32.stride(1) == 1
            op.x_f32, op.x_f32_pitch, op.x_f32_cols = x_f32.data_ptr(), x_f32.stride(0), x_f32.shape[1]
        out_vecs = out_vecs or [None] * len(self.layers)
        for l, ov in enumerate(out_vecs):
            L = op.layer[l]
            if ov is None:
                L.out_vec, L.out_vec_pitch = None, 0
            else:
                assert ov.dtype == torch.float32 and ov.stride(1) == 1 and ov.shape[0] >= n
                L.out_vec, L.out_vec_pitch = ov.data_ptr(), ov.stride(0)
        op.emb_half = None if emb_half is None else emb_half.data_ptr()
        op.emb_half_dtype = 1 if (emb_half is not None and emb_half.dtype == torch.float16) else 0
        op.label = None if label is None else label.data_ptr()
        op.prob = None if prob is None else prob.data_ptr()
        op.label_f = op.prob_f = op.count_cell = None
        op.lp_pitch = 0
        if payload is not None:
            D = payload.shape[1] - 2
            base, pitch = payload.data_ptr(), payload.stride(0)
            op.label_f, op.prob_f, op.lp_pitch = base + 4 * D, base + 4 * (D + 1), pitch
        if count_cell is not None:
            op.count_cell, op.count_value = count_cell.data_ptr(), int(count_value)
        op.thr = float(thr or 0.0)
        op.thr_class = None if thr_class is None else thr_class.data_ptr()
        op.n_classes = int(n_classes)
        self._keep = (x, x_f32, out_vecs, emb_half, label, prob, payload, thr_class)
        _lib.call("vnfr_tail_run", C.byref(op), int(n), _lib.stream_ptr())
