"""Multi-GPU correctness of the sharded pipeline: N ranks, each running the whole path on its
contiguous slice of the frames, ONE all-gather of the send buffers the fused tail kernel filled -> the rank-order
concatenation must equal the single-GPU result on all frames.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _models(dev):
    from oracle import nets
    from conftest import golden_encoder_state_dict
    from vn_celeb_face_recognition_b200.models import MTCNN, InceptionResnetV1, MLPModel
    det = MTCNN(image_size=160, keep_all=True, min_face_size=50, device=dev)
    enc = InceptionResnetV1(pretrained=None, device=dev).eval()
    enc.load_state_dict(golden_encoder_state_dict())
    mlp = MLPModel(512, 1001).to(dev).eval()
    mlp.load_state_dict(nets.make_mlp_state_dict(1001, seed=0))
    return det, enc, mlp


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import synth
    from vn_celeb_face_recognition_b200 import pipeline, dist as vdist
    frames = np.concatenate([synth.frames("small", 5, first_seed=0), synth.frames("small", 4, first_seed=20)])    # 9 frames: uneven shards
    lo, hi = vdist.shard_range(len(frames), rank, world)
    det, enc, mlp = _models(dev)
    fp = pipeline.FacePipeline(det, enc, mlp, (160, 160), "similarity", max_faces_per_frame=8)
    side = torch.cuda.Stream(dev)
    for rep in range(2):                                   # second pass: alternate send buffer, side-stream exchange
        out = fp.run_device(torch.from_numpy(frames[lo:hi]).to(dev))
        # every rank must offer the same send-buffer shape: capacity = frames of the LARGEST shard x max_faces_per_frame
        cap = vdist.shard_range(len(frames), 0, world)[1] * fp.max_faces_per_frame
        payload = out["payload"]
        if payload.shape[0] != cap + 1:                    # smaller shard: re-home its rows in a buffer of the common shape
            full = torch.zeros(cap + 1, payload.shape[1], device=dev)
            n = out["n_faces"]
            full[:n] = payload[:n]
            full[cap, 0] = float(n)
            payload = full
        gathered, D, ev = vdist.all_gather_payload(payload, stream=side if rep else None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        emb, label, prob, counts = vdist.compact_faces(gathered, D)
    # the same exchange without a collective kernel: symmetric send buffer, copy-engine pull from the peers (dist.PeerGather)
    pg = vdist.PeerGather()
    sym = pg.alloc((cap + 1, payload.shape[1]), dev)
    for rep in range(3):                                   # repeated: the barriers' channels / the buffer are reused
        sym.copy_(payload)
        g2, D2, ev2 = pg.gather(sym, stream=side if rep & 1 else None)
        torch.cuda.current_stream().wait_event(ev2)
        assert D2 == D and torch.equal(g2, gathered), "PeerGather differs from the NCCL all_gather"
        sym.zero_()                                        # safe: ev2 = every peer has read the buffer
    if rank == 0:
        ref = fp(torch.from_numpy(frames).to(dev))         # single-GPU result on ALL frames
        q.put((emb.cpu().numpy(), label.cpu().numpy(), prob.cpu().numpy(), counts.cpu().numpy(),
               np.concatenate([r["emb"] for r in ref]), np.concatenate([r["labels"] for r in ref]),
               np.concatenate([r["probs"] for r in ref]), [len(r["labels"]) for r in ref], (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_pipeline_allgather_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    emb, label, prob, counts, emb1, label1, prob1, per_frame, span = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert span == (0, 5) and counts.sum() == len(label1) == sum(per_frame)
    assert counts.tolist() == [sum(per_frame[:5]), sum(per_frame[5:])]
    np.testing.assert_array_equal(label, label1)                       # rank-order concatenation = single-GPU order
    np.testing.assert_allclose(emb, emb1, atol=2e-6)
    np.testing.assert_allclose(prob, prob1, atol=1e-6)
