"""Scratch: time the encoder + MLP forward at several batch sizes (CUDA events)."""
import sys, time
import torch
sys.path.insert(0, ".")
from oracle import nets
from vn_celeb_face_recognition_b200.models import InceptionResnetV1, MLPModel
from vn_celeb_face_recognition_b200 import _lib

dev = torch.device("cuda:0")
enc = InceptionResnetV1(device=dev).eval()
enc.load_state_dict(nets.make_encoder_state_dict(0, calibrate=False))
mlp = MLPModel(512, 1001).to(dev).eval()
for chunk in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    enc.chunk = chunk
    n = max(chunk, 1024)
    from vn_celeb_face_recognition_b200 import encoder_plan
    x = torch.randn(n, 160, 160, 8, device=dev).to(encoder_plan.HALF)
    for _ in range(2):
        e, e16 = enc.embed_nhwc8(x)
        lab, pr = mlp.classify_half(e16)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    t0.record()
    for _ in range(3):
        e, e16 = enc.embed_nhwc8(x)
        lab, pr = mlp.classify_half(e16)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print("chunk %4d batch %5d: %.2f ms  %.0f embeds/s  %.1f TFLOP/s  launches/fwd %d" % (
        chunk, n, ms, n / ms * 1e3, n * 2.8415e9 / ms / 1e9, (_lib.launch_count() - l0) // 3))
