"""Scratch: what bounds the Block17 projection conv (1x1 256 -> 896 + residual) ?"""
import sys, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import _lib, encoder_plan as ep
dev = torch.device("cuda:0")
dt = torch.float16
n = 768
def run(cin, cout, res, bn, label, split=None):
    x = torch.randn(n, 8, 8, cin, device=dev).to(dt)
    w = torch.randn(cout, cin, 1, 1) * 0.05
    pc = ep.pack_conv(w, None, torch.zeros(cout), dev, block_n=bn, dtype=dt)
    out = torch.empty(n, 8, 8, cout, dtype=dt, device=dev)
    r = torch.randn(n, 8, 8, cout, device=dev).to(dt) if res else None
    ol = ep.OpList()
    if split:
        o1 = torch.empty(n, 8, 8, cout - split, dtype=dt, device=dev)
        o0 = torch.empty(n, 8, 8, split, dtype=dt, device=dev)
        ol.conv(pc, ep.View(x), ep.View(o0), dst1=ep.View(o1), n_split=split)
    else:
        ol.conv(pc, ep.View(x), ep.View(out if not res else r), residual=ep.View(r) if res else None)
    for _ in range(3): ol.run()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): ol.run()
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 100
    c = ol.ops[0].conv
    import ctypes as C
    lib = _lib.lib(); lib.vnfr_ig_debug.argtypes = [C.c_void_p]
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.vnfr_ig_debug(C.c_void_p(buf.data_ptr())); ol.run(); torch.cuda.synchronize(); lib.vnfr_ig_debug(C.c_void_p(0))
    b = buf.tolist()
    print("     kernel %d | mma: wait full %d, wait tmem %d | epi: wait acc %d, wait res/cfree %d, work %d, store %d | prod: wait empty %d, wait cfree %d"
          % (b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8]))
    print("%-44s %7.1f us  %6.1f TFLOP/s  (a_mode %d epi %d)" % (label, us, 2.0 * n * 64 * cin * cout / us / 1e6, c.a_mode, c.epi_mode))
run(256, 896, True, 256, "256->896 +res in place, N=256")
run(256, 896, False, 256, "256->896 no res, N=256")
run(256, 896, True, 128, "256->896 +res in place, N=128")
run(256, 896, False, 128, "256->896 no res, N=128")
run(256, 256, False, 256, "256->256 no res (1 N tile)")
run(896, 256, False, 256, "896->256 no res (1 N tile)")
run(896, 256, False, 256, "896->256 split 128|128 (old epilogue)", split=128)
run(896, 896, False, 256, "896->896 no res")
run(1792, 1792, False, 256, "1792->1792 no res")
