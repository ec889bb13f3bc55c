"""Scratch: time the encoder + MLP forward (CUDA events). usage: enc_bench.py [batch] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
from vn_celeb_face_recognition_b200 import _lib, encoder_plan

dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = InceptionResnetV1(device=dev).eval()
mlp = MLPModel(512, 1001).to(dev).eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
enc.chunk = n
x = torch.randn(n, 80, 80, 16, device=dev).to(encoder_plan.HALF)
for _ in range(2):
    e, e16 = enc.embed_s2d(x, 160)
    lab, pr = mlp.classify_half(e16)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = _lib.launch_count()
t0.record()
for _ in range(iters):
    e, e16 = enc.embed_s2d(x, 160)
    lab, pr = mlp.classify_half(e16)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / iters
print("batch %5d: %.3f ms  %.0f embeds/s  %.1f TFLOP/s  launches/fwd %d" % (
    n, ms, n / ms * 1e3, n * 2.8415e9 / ms / 1e9, (_lib.launch_count() - l0) // iters))
