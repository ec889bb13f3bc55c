"""Scratch: does ONE resident foreign CTA (what a collective kernel waiting for its peer is) slow the persistent kernels of the
step down?  Times device-resident steps alone, and with a one-block spin kernel (torch.cuda._sleep) alive on a side stream."""
import sys, torch
sys.path.insert(0, ".")
import bench
from vn_celeb_face_recognition_b200 import pipeline
dev = torch.device("cuda:0")
det, enc, cls = bench.build_models(dev)
fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity")
frames = torch.from_numpy(bench.make_frames(64, 0)).to(dev)
side = torch.cuda.Stream(dev)

def run(n, spin_ms):
    for _ in range(3):
        fp.run_device(frames, pipelined=True)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        if spin_ms:
            with torch.cuda.stream(side):
                torch.cuda._sleep(int(spin_ms * 1.9e6))      # ~spin_ms milliseconds at 1.9 GHz, one block of one thread
        fp.run_device(frames, pipelined=True)
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n

for spin in (0, 0.3, 1.0, 3.0, 8.0, 0):
    print("spin %.1f ms per step on a side stream: %.3f ms per step" % (spin, run(20, spin)), flush=True)
