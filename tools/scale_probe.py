#!/usr/bin/env python
"""Where does the per-step time go when N > 1?  Runs the device-resident step in several variants and prints the
event-timed ms/step of each (max over ranks):
  plain          run_device only, no collective, no extra sync
  sync           run_device + torch.cuda.synchronize() after every step (what a blocking collective does to the host)
  gather         run_device + dist.all_gather_faces (the bench's step)
Launch: python tools/scale_probe.py            (1 GPU)   or   torchrun ... tools/scale_probe.py   (N GPUs)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import bench
from vn_celeb_face_recognition_b200 import pipeline, dist as vdist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = 64
    det, enc, cls = bench.build_models(dev)
    enc.chunk = 1024
    fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity")
    frames = torch.from_numpy(bench.make_frames(B, rank * B)).to(dev)
    steps = int(os.environ.get("STEPS", "10"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(name, fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        host = []
        for _ in range(steps):
            h0 = time.perf_counter()
            fn()
            host.append(time.perf_counter() - h0)
        e1.record()
        barrier()
        w1 = time.perf_counter()
        ms = e0.elapsed_time(e1) / steps
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = ms
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            print("%-12s ms/step per rank %s   wall %.3f   host-per-step min/med %.2f/%.2f ms" % (
                name, " ".join("%.3f" % v for v in t.tolist()), 1e3 * (w1 - w0) / steps, 1e3 * min(host),
                1e3 * sorted(host)[len(host) // 2]), flush=True)

    def plain():
        return fp.run_device(frames)

    def sync():
        out = fp.run_device(frames)
        torch.cuda.synchronize()
        return out

    def gather():
        out = fp.run_device(frames)
        return vdist.all_gather_faces(out["emb"], out["label"], out["prob"])

    def solo():
        # every rank in turn runs alone while the others wait: interference between ranks shows as plain > solo
        for r in range(world):
            if r == rank:
                for _ in range(2):
                    fp.run_device(frames)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fp.run_device(frames)
                e1.record()
                torch.cuda.synchronize()
                print("solo rank %d: %.3f ms/step" % (rank, e0.elapsed_time(e1) / steps), flush=True)
            barrier()

    def gather_padded():
        out = fp.run_device(frames)
        return vdist.all_gather_faces_padded(out["emb"], out["label"], out["prob"], B * fp.max_faces_per_frame)

    run("plain", plain)
    if world > 1 and os.environ.get("SOLO"):
        solo()
    if world > 1:
        run("gather_padded", gather_padded)
        run("gather", gather)
        run("gather_padded", gather_padded)
    if rank == 0:
        print("-- with bench.ClockSampler polling NVML on rank 0", flush=True)
    smp = bench.ClockSampler(local)
    if rank == 0:
        smp.start()
    run("plain", plain)
    if world > 1:
        run("gather_padded", gather_padded)
        run("gather", gather)
    if rank == 0:
        print(smp.stop(), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
