"""Scratch: where does the end-to-end (host frames) time go?"""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from vn_celeb_face_recognition_b200 import pipeline
dev = torch.device("cuda:0")
det, enc, cls = bench.build_models(dev)
enc.chunk = 1024
fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity")
fr = bench.make_frames(64, 0)
pin = torch.from_numpy(fr).pin_memory()
devf = pin.to(dev)
def timeit(fn, n=5, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
buf = torch.empty_like(devf)
print("H2D 398 MB alone: %.2f ms" % timeit(lambda: buf.copy_(pin, non_blocking=True)))
print("device-resident run_device: %.2f ms" % timeit(lambda: fp.run_device(devf)))
print("device-resident __call__ (incl. D2H + host post): %.2f ms" % timeit(lambda: fp(devf)))
for sb, fsb in ((64, 64), (32, 8), (16, 4), (16, 16), (8, 4), (8, 8)):
    fp.sub_batch, fp.first_sub_batch = sb, fsb
    if sb == 64:
        print("host frames, no chunking: %.2f ms" % timeit(lambda: fp(pin)))
    else:
        print("host frames, sub_batch %d first %d: %.2f ms" % (sb, fsb, timeit(lambda: fp(pin))))
# detection only, chunked vs whole
fp.sub_batch, fp.first_sub_batch = 16, 4
bounds = fp._sub_batches(64)
print("detect whole: %.2f ms" % timeit(lambda: det.detect_device(devf)))
print("detect chunked (device frames, no copies): %.2f ms" % timeit(lambda: det.detect_device_chunked(devf, None, bounds)))
for nch in (2, 4, 8):
    sz = 64 // nch
    b = [(i * sz, (i + 1) * sz) for i in range(nch)]
    print("detect %d chunks of %d on 2 streams: %.2f ms" % (nch, sz, timeit(lambda: det.detect_device_chunked(devf, None, b))))
