"""Scratch: encoder + MLP time for 768 faces as a function of the internal chunk size."""
import sys, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
from vn_celeb_face_recognition_b200 import encoder_plan
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = InceptionResnetV1(device=dev).eval()
mlp = MLPModel(512, 1001).to(dev).eval()
n = 768
x = torch.randn(n, 80, 80, 16, device=dev).to(encoder_plan.HALF)
for chunk in (768, 384, 256, 192, 128):
    enc.chunk = chunk
    for _ in range(3):
        e, e16 = enc.embed_s2d(x, 160); mlp.classify_half(e16)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        e, e16 = enc.embed_s2d(x, 160); mlp.classify_half(e16)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print("chunk %4d: %.3f ms  %.1f TFLOP/s" % (chunk, ms, n * 2.8415e9 / ms / 1e9))
