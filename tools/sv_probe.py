"""Scratch: cycle breakdown of the shifted-view conv kernel (CTA 0) for a few layer shapes."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import _lib, encoder_plan as ep

dev = torch.device("cuda:0")
lib = _lib.lib()
lib.vnfr_sv_debug.argtypes = [C.c_void_p]
dt = torch.float16
cases = [("2a", 768, 79, 79, 32, 32, (3, 3), (0, 0)), ("2b", 768, 77, 77, 32, 64, (3, 3), (1, 1)),
         ("b35", 768, 17, 17, 32, 32, (3, 3), (1, 1)), ("1x7", 768, 8, 8, 128, 128, (1, 7), (0, 3)),
         ("7x1", 768, 8, 8, 128, 128, (7, 1), (3, 0))]
for name, n, h, w, cin, cout, k, pad in cases:
    for sv in ([32, 64] if cin == 32 else [64]):
        x = torch.randn(n, h, w, cin, device=dev).to(dt)
        wt = torch.randn(cout, cin, k[0], k[1]) * 0.05
        pc = ep.pack_conv(wt, None, torch.zeros(cout), dev, cin_pad=max(cin, sv), dtype=dt)
        oh, ow = h + 2 * pad[0] - k[0] + 1, w + 2 * pad[1] - k[1] + 1
        out = torch.empty(n, oh, ow, cout, dtype=dt, device=dev)
        ol = ep.OpList()
        ol.conv(pc, ep.View(x), ep.View(out), pad=pad, sv=sv)
        assert ol.ops[0].conv.a_mode == 3
        ol.run(); torch.cuda.synchronize()
        buf = torch.zeros(8, dtype=torch.int64, device=dev)
        lib.vnfr_sv_debug(C.c_void_p(buf.data_ptr()))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); ol.run(); t1.record(); torch.cuda.synchronize()
        lib.vnfr_sv_debug(C.c_void_p(0))
        b = buf.tolist()
        print("%-4s sv%d  %.1f us | kernel %d cyc | mma: wait A %d, wait tmem %d, wait W %d | epi: wait acc %d, work %d | prod wait %d"
              % (name, sv, t0.elapsed_time(t1) * 1e3, b[0], b[1], b[2], b[3], b[5], b[6], b[7]))
