"""world_size-2 gloo test of the only collective on the path: the ragged all-gather of per-face results, and of the
frame sharding.  CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vn_celeb_face_recognition_b200 import dist as vdist
    n_frames = 7
    lo, hi = vdist.shard_range(n_frames, rank, world)
    # per-frame face counts are ragged; embeddings are a deterministic function of the global face index
    counts = [3, 0, 2, 5, 1, 0, 4]
    start = sum(counts[:lo])
    n = sum(counts[lo:hi])
    idx = torch.arange(start, start + n)
    emb = torch.stack([torch.full((8,), float(i)) for i in idx]) if n else torch.zeros(0, 8)
    label = idx * 3
    prob = idx.float() / 100
    e, l, p, c = vdist.all_gather_faces(emb, label, prob)
    # the sync-free fixed-capacity variant must compact to exactly the same result
    payload, D = vdist.all_gather_faces_padded(emb, label, prob, cap=12)
    assert payload.shape == (world, 13, 10)
    e2, l2, p2, c2 = vdist.compact_faces(payload, D)
    assert torch.equal(e, e2) and torch.equal(l, l2) and torch.equal(p, p2) and torch.equal(c, c2)
    # the send-buffer form (what the fused tail kernel fills on the GPU): rows [emb | label | prob], count in [cap, 0]
    send = torch.full((13, 10), float("nan"))
    send[:n, :8], send[:n, 8], send[:n, 9], send[12, 0] = emb, label.float(), prob, float(n)
    pay3, D3, ev = vdist.all_gather_payload(send)
    assert ev is None and pay3.shape == (world, 13, 10)
    e3, l3, p3, c3 = vdist.compact_faces(pay3, D3)
    assert torch.equal(e, e3) and torch.equal(l, l3) and torch.equal(p, p3) and torch.equal(c, c3)
    try:                                                          # n > cap raises on every rank before the collective
        vdist.all_gather_faces_padded(emb, label, prob, cap=n - 1)
        raise AssertionError("capacity overflow must raise")
    except ValueError:
        pass
    if rank == 0:
        out.put((e.numpy(), l.numpy(), p.numpy(), c.numpy(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_all_gather_reproduces_single_process_order():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    e, l, p, c, rng = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    total = 15
    assert e.shape == (total, 8) and c.tolist() == [10, 5] and rng == (0, 4)
    np.testing.assert_array_equal(e[:, 0], np.arange(total, dtype=np.float32))      # rank-order concat = global order
    np.testing.assert_array_equal(l, np.arange(total) * 3)
    np.testing.assert_allclose(p, np.arange(total) / 100, atol=1e-7)


def test_shard_range_partitions_all_items():
    from vn_celeb_face_recognition_b200 import dist as vdist
    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [vdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _topk_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vn_celeb_face_recognition_b200 import gallery
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(6, 40, generator=g)                     # queries x global gallery rows (same on every rank)
    scores[0, 3] = scores[0, 27] = 9.0                            # a tie across shards: the lower global index must win
    lo, hi = rank * 20, (rank + 1) * 20                           # this rank's gallery shard
    v, i = torch.topk(scores[:, lo:hi], 5, dim=1)
    mv, mi = gallery.merge_topk(v, i + lo, k=5)
    if rank == 0:
        out.put((scores.numpy(), mv.numpy(), mi.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gallery_topk_merge():
    """merge_topk: per-shard top-5 lists all-gathered over 2 ranks equal the top-5 of the unsharded score matrix."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_topk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    scores, mv, mi = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    rv, ri = torch.topk(torch.from_numpy(scores), 5, dim=1)
    np.testing.assert_allclose(mv, rv.numpy())
    assert mi[0, :2].tolist() == [3, 27]
    np.testing.assert_array_equal(np.sort(mi, axis=1), np.sort(ri.numpy(), axis=1))
