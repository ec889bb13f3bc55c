"""Scratch: end-to-end (pinned host frames) ms per batch with two batches in flight, for several sub-batch sizes."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from vn_celeb_face_recognition_b200 import pipeline
dev = torch.device("cuda:0")
det, enc, cls = bench.build_models(dev)
fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity")
pin = torch.from_numpy(bench.make_frames(64, 0)).pin_memory()

def run(n):
    pend = None
    for _ in range(n):
        nxt = fp.submit(pin)
        if pend is not None:
            pend.result()
        pend = nxt
    pend.result()

for sb, fsb in ((8, 4), (8, 8), (16, 4), (16, 8), (16, 16), (32, 8), (32, 32)):
    fp.sub_batch, fp.first_sub_batch = sb, fsb
    run(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(20)
    torch.cuda.synchronize()
    print("sub_batch %2d first %2d: %.3f ms per batch" % (sb, fsb, (time.perf_counter() - t0) / 20 * 1e3), flush=True)
