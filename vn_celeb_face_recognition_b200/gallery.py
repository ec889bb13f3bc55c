"""Cosine top-k of L2-normalised embeddings against a gallery shard (BASELINE.json config 5 / north star: "optional cosine
top-k against a sharded gallery"; the reference itself has no gallery search -- SURVEY.md section 8d).

Each rank holds a contiguous block of gallery rows as 16-bit [rows][512]; ``GalleryShard.topk`` is one launch of the fused
score-GEMM + top-k kernel (csrc/gallery_topk.cu: the score matrix never leaves TMEM), and ``merge_topk`` combines the
per-rank lists (one all_gather of k values + k global indices per query).
"""
import torch
import torch.distributed as dist

from . import _lib, encoder_plan


class GalleryShard:
    """``emb``: (g, 512) float tensor of unit vectors on a CUDA device; ``index_offset``: global row index of row 0."""

    def __init__(self, emb, index_offset=0, dtype=None):
        if not emb.is_cuda:
            raise _lib.VnfrError("GalleryShard needs CUDA tensors: this package has no CPU path")
        self.dtype = dtype or encoder_plan.HALF
        self.g, self.d = emb.shape
        assert self.d == 512, "embedding size must be 512 (InceptionResnetV1)"
        self.index_offset = int(index_offset)
        dev = emb.device
        g_pad = -(-max(self.g, 1) // 256) * 256
        self.w = torch.zeros(g_pad, self.d, dtype=self.dtype, device=dev)
        self.w[:self.g] = emb.to(self.dtype)

    def topk(self, q, k=5, sms=148):
        """q: (n, 512) unit vectors (any float dtype, CUDA).  Returns (values fp32 (n,k), global indices int64 (n,k)) of the
        k most similar rows of THIS shard, best first; indices of missing entries (g < k) are -1.  ONE launch of the fused
        score-GEMM + top-k kernel (csrc/gallery_topk.cu); with few queries the gallery is split across CTAs and the per-split
        lists are merged here."""
        assert 1 <= k <= 8
        n = q.shape[0]
        dev = q.device
        q16 = q.to(self.dtype).contiguous()
        g_tiles = self.w.shape[0] // 256
        m_tiles = -(-max(n, 1) // 128)
        splits = max(1, min(g_tiles, -(-sms // m_tiles))) if m_tiles < sms else 1
        vals = torch.empty(splits, n, 8, dtype=torch.float32, device=dev)
        idx = torch.empty(splits, n, 8, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("vnfr_gallery_topk", _lib.ptr(q16), n, _lib.ptr(self.w), self.g, self.w.shape[0],
                      encoder_plan.dtype_code(self.dtype), splits, self.index_offset, _lib.ptr(vals), _lib.ptr(idx), _lib.stream_ptr())
        v = vals.permute(1, 0, 2).reshape(n, splits * 8)
        i = idx.permute(1, 0, 2).reshape(n, splits * 8).to(torch.int64)
        if splits > 1:
            key = torch.where(i < 0, torch.full_like(i, 1 << 62), i)
            order = torch.argsort(key, dim=1, stable=True)                # (value desc, index asc): index first, then value
            v, i = torch.gather(v, 1, order), torch.gather(i, 1, order)
            order = torch.argsort(v, dim=1, descending=True, stable=True)
            v, i = torch.gather(v, 1, order), torch.gather(i, 1, order)
        return v[:, :k].contiguous(), i[:, :k].contiguous()


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
