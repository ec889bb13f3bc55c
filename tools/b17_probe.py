"""Fused Block17 alone (n images, 10 blocks back to back): time per block + the cycle breakdown of CTA 0."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vn_celeb_face_recognition_b200 import _lib, encoder_plan as ep, synthetic
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
sd = {k: v.to(dev) for k, v in synthetic.encoder_state_dict_seed0().items() if k.startswith("repeat_2.")}
dt = torch.float16
x = torch.relu(torch.randn(n, 8, 8, 896, device=dev)).to(dt)
ol = ep.OpList()
for i in range(10):
    p = "repeat_2.%d" % i
    ol.block17(ep.pack_basic(sd, [p + ".branch0", p + ".branch1.0"], dev, block_n=256, dtype=dt), ep.pack_basic(sd, [p + ".branch1.1"], dev, dtype=dt),
               ep.pack_basic(sd, [p + ".branch1.2"], dev, dtype=dt), ep.pack_projection(sd, p + ".conv2d", 0.10, dev, dtype=dt), x)
for _ in range(3):
    ol.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ol.run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 50
flop = 2.0 * n * 64 * (896 * 256 + 2 * 896 * 128 + 256 * 896)
print("n=%d: %.1f us per block, %.0f TFLOP/s" % (n, us, flop / us / 1e6))
dbg = torch.zeros(32, dtype=torch.int64, device=dev)
fn = _lib.lib().vnfr_b17_debug
fn.argtypes = [C.c_void_p]; fn.restype = C.c_int
fn(C.c_void_p(dbg.data_ptr()))
os.environ["VNFR_NO_GRAPH"] = "1"
ol2 = ep.OpList(); ol2.ops = ol.ops[:1]; ol2.keep = ol.keep
_lib.call("vnfr_run_ops", (_lib.Op * 1)(*ol.ops[:1]), 1, _lib.stream_ptr())
torch.cuda.synchronize()
fn(C.c_void_p(0))
d = dbg.cpu().tolist()
tiles = -(-((n + 1) // 2 - 0) // min(148, (n + 1) // 2))
names = {0: "kernel", 1: "mma wait ring (gemm1)", 2: "mma wait ring (convs)", 3: "mma wait ring (proj)", 4: "mma wait tpad", 5: "mma wait cat",
         6: "mma wait tmem drained", 8: "epi wait gemm1", 9: "E1 work", 10: "epi wait 1x7", 11: "E2 work", 12: "epi wait 7x1", 13: "E3 work",
         14: "epi wait proj chunk", 15: "epi wait residual", 16: "panel math", 17: "panel barrier+store"}
print("CTA 0, %d tiles:" % tiles)
for k in sorted(names):
    print("  %-26s %9d cycles  (%7d per tile)" % (names[k], d[k], d[k] // max(tiles, 1)))
