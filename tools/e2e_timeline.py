"""Scratch: timeline of the two-batches-in-flight end-to-end loop (pinned host frames): per batch, when the H2D copies of its
sub-batches land, when its cascade + face crops are done, when its encoder + read-back are done (CUDA events, ms since a base
event), and the host times of submit() / result()."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from vn_celeb_face_recognition_b200 import pipeline
dev = torch.device("cuda:0")
det, enc, cls = bench.build_models(dev)
nv12 = len(sys.argv) > 1 and sys.argv[1] == "nv12"
fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity", input_format="nv12" if nv12 else "rgb")
frames = bench.make_frames(64, 0)
if nv12:
    import cv2, numpy as np
    B, H, W = frames.shape[:3]
    buf = np.empty((B, H * 3 // 2, W), np.uint8)
    for i in range(B):
        i420 = cv2.cvtColor(frames[i], cv2.COLOR_RGB2YUV_I420)
        buf[i, :H] = i420[:H]
        buf[i, H:] = np.stack([i420[H:H + H // 4].reshape(H // 2, W // 2), i420[H + H // 4:].reshape(H // 2, W // 2)], axis=-1).reshape(H // 2, W)
    frames = buf
pin = torch.from_numpy(frames).pin_memory()
if len(sys.argv) > 2:
    fp.sub_batch, fp.first_sub_batch = int(sys.argv[2]), int(sys.argv[3])

# hook: keep the H2D events of every call
orig = fp.det.detect_device_chunked
h2d_events = []
def hooked(buf, events, bounds, **kw):
    h2d_events.append(list(events) if events else [])
    return orig(buf, events, bounds, **kw)
fp.det.detect_device_chunked = hooked

def loop(n, rec=None):
    pend = None
    for i in range(n):
        t0 = time.perf_counter()
        if rec is not None:
            ev = torch.cuda.Event(enable_timing=True); ev.record(); rec["sub_ev"].append(ev)
        nxt = fp.submit(pin)
        t1 = time.perf_counter()
        if rec is not None:
            rec["crops"].append(fp._host_crops_done[1 - fp._host_slot]); rec["done"].append(nxt.done)
        if pend is not None:
            pend.result()
        t2 = time.perf_counter()
        if rec is not None:
            rec["host"].append((t0, t1, t2))
        pend = nxt
    pend.result()

loop(4)
torch.cuda.synchronize()
# timing events need enable_timing: patch torch.cuda.Event default for this run
_Ev = torch.cuda.Event
torch.cuda.Event = lambda *a, **k: _Ev(enable_timing=True)
h2d_events.clear()
rec = {"sub_ev": [], "crops": [], "done": [], "host": []}
base = _Ev(enable_timing=True); base.record()
tb = time.perf_counter()
loop(8, rec)
torch.cuda.synchronize()
for i in range(8):
    h = rec["host"][i]
    print("batch %d: submit host %.2f..%.2f ms (result until %.2f) | GPU: submit marker %.2f, H2D landed %s, crops done %.2f, all done %.2f" % (
        i, (h[0] - tb) * 1e3, (h[1] - tb) * 1e3, (h[2] - tb) * 1e3, base.elapsed_time(rec["sub_ev"][i]),
        " ".join("%.2f" % base.elapsed_time(e) for e in h2d_events[i]), base.elapsed_time(rec["crops"][i]), base.elapsed_time(rec["done"][i])))
