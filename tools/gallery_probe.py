"""Scratch: kernel-only time of vnfr_gallery_topk (config-5 sizes) for several split counts, one-CTA vs two-CTA MMA
(VNFR_GALLERY_1CTA=1 selects the one-CTA kernel for the whole process)."""
import os, sys, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import _lib, encoder_plan
dev = torch.device("cuda:0")
m, g = 122880, 125000
g_pad = -(-g // 256) * 256
torch.manual_seed(0)
Q = torch.nn.functional.normalize(torch.randn(m, 512, device=dev), dim=1).half()
G = torch.zeros(g_pad, 512, device=dev, dtype=torch.float16)
G[:g] = torch.nn.functional.normalize(torch.randn(g, 512, device=dev), dim=1).half()
for splits in (1, 2, 4, 6, 8, 16):
    vals = torch.empty(splits, m, 8, dtype=torch.float32, device=dev)
    idx = torch.empty(splits, m, 8, dtype=torch.int32, device=dev)
    def run():
        _lib.call("vnfr_gallery_topk", _lib.ptr(Q), m, _lib.ptr(G), g, g_pad, 1, splits, 0, _lib.ptr(vals), _lib.ptr(idx), _lib.stream_ptr())
    run(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        run()
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print("%s splits %2d: %.2f ms  %.0f TFLOP/s" % ("1-CTA" if os.environ.get("VNFR_GALLERY_1CTA") else "2-CTA", splits, ms, 2.0 * m * g_pad * 512 / ms / 1e9), flush=True)
