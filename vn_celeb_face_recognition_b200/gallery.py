"""Cosine top-k of L2-normalised embeddings against a gallery shard (BASELINE.json config 5 / north star: "optional cosine
top-k against a sharded gallery"; the reference itself has no gallery search -- SURVEY.md section 8d).

Each rank holds a contiguous block of gallery rows as 16-bit [rows][512] "weights" of the tcgen05 GEMM kernel
(csrc/igemm_conv.cu): scores = Q . G^T come out tile by tile in fp32, ``vnfr_topk_rows`` keeps a running top-k per query,
and ``merge_topk`` combines the per-rank lists (one all_gather of k values + k global indices per query).
"""
import torch
import torch.distributed as dist

from . import _lib, encoder_plan


class GalleryShard:
    """``emb``: (g, d) float tensor of unit vectors on a CUDA device; ``index_offset``: global row index of row 0."""

    #: gallery rows scored per GEMM launch (score buffer = queries x tile fp32)
    tile = 32768
    #: queries per pass
    q_chunk = 1024

    def __init__(self, emb, index_offset=0, dtype=None):
        if not emb.is_cuda:
            raise _lib.VnfrError("GalleryShard needs CUDA tensors: this package has no CPU path")
        self.dtype = dtype or encoder_plan.HALF
        self.g, self.d = emb.shape
        assert self.d % 64 == 0, "embedding size must be a multiple of 64"
        self.index_offset = int(index_offset)
        dev = emb.device
        g_pad = -(-max(self.g, 1) // 256) * 256
        self.w = torch.zeros(g_pad, self.d, dtype=self.dtype, device=dev)
        self.w[:self.g] = emb.to(self.dtype)
        self.bias = torch.zeros(g_pad, dtype=torch.float32, device=dev)
        self._plans = {}

    def _plan(self, n, t0, t1):
        key = (n, t0, t1)
        if key not in self._plans:
            dev = self.w.device
            rows = t1 - t0                                          # multiple of 256 (padded rows score 0 and are masked)
            pc = encoder_plan.PackedConv(self.w[t0:t1], self.bias[t0:t1], 1, 1, self.d, rows, 256)
            x = torch.zeros(n, 1, 1, self.d, dtype=self.dtype, device=dev)
            scores = torch.empty(n, rows, dtype=torch.float32, device=dev)
            ol = encoder_plan.OpList()
            ol.conv(pc, encoder_plan.View(x), None, relu=False, out_f32=scores)
            self._plans[key] = (x, scores, ol)
        return self._plans[key]

    def topk(self, q, k=5):
        """q: (n, d) unit vectors (any float dtype, CUDA).  Returns (values fp32 (n,k), global indices int64 (n,k)) of the
        k most similar rows of THIS shard, best first; indices of missing entries (g < k) are -1."""
        assert 1 <= k <= 8
        n = q.shape[0]
        dev = q.device
        vals = torch.full((n, k), float("-inf"), dtype=torch.float32, device=dev)
        idx = torch.full((n, k), 0x7fffffff, dtype=torch.int32, device=dev)
        q16 = q.to(self.dtype).contiguous()
        for s in range(0, n, self.q_chunk):
            m = min(self.q_chunk, n - s)
            v_s, i_s = vals[s:s + m], idx[s:s + m]
            first = True
            for t0 in range(0, self.w.shape[0], self.tile):
                t1 = min(self.w.shape[0], t0 + self.tile)
                valid = min(self.g, t1) - t0
                if valid <= 0:
                    break
                x, scores, ol = self._plan(m, t0, t1)
                x.view(m, self.d).copy_(q16[s:s + m])
                ol.run()
                _lib.call("vnfr_topk_rows", _lib.ptr(scores), m, valid, scores.shape[1], k, self.index_offset + t0,
                          0 if first else 1, _lib.ptr(v_s), _lib.ptr(i_s), _lib.stream_ptr())
                first = False
        idx64 = idx.to(torch.int64)
        idx64[idx == 0x7fffffff] = -1
        return vals, idx64


def merge_topk(vals, idx, k=None, group=None):
    """Per-rank (n,k) lists -> the global top-k over all ranks (all_gather of values and global indices; NCCL on GPUs,
    gloo in the CPU tests).  Ties go to the lower global index, like the per-shard kernel."""
    k = k or vals.shape[1]
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        gv = [torch.empty_like(vals) for _ in range(world)]
        gi = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(gv, vals.contiguous(), group=group)
        dist.all_gather(gi, idx.contiguous(), group=group)
        vals, idx = torch.cat(gv, 1), torch.cat(gi, 1)
    # sort by (value desc, index asc): stable sort by index first, then by value
    order = torch.argsort(idx, dim=1, stable=True)
    vals, idx = torch.gather(vals, 1, order), torch.gather(idx, 1, order)
    order = torch.argsort(vals, dim=1, descending=True, stable=True)
    return torch.gather(vals, 1, order)[:, :k].contiguous(), torch.gather(idx, 1, order)[:, :k].contiguous()
