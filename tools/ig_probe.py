"""Cycle breakdown of CTA 0 of igemm_conv_kernel for selected encoder ops (g_ig_dbg, csrc/igemm_conv.cu).
usage: ig_probe.py [batch] idx idx ...     (indices into the op list of tools/enc_layers.py)"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200.models import InceptionResnetV1
from vn_celeb_face_recognition_b200 import _lib

NAMES = ["kernel", "mma wait full", "mma wait tmem", "epi wait acc", "epi wait res/C", "epi work", "store", "prod wait empty",
         "prod wait C"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = InceptionResnetV1(device=dev).eval()
n = int(sys.argv[1])
idx = [int(a) for a in sys.argv[2:]]
plan = enc._plan(n, 160, 160, dev)
plan.x0.normal_()
plan.run()
plan.run()
torch.cuda.synchronize()
L = _lib.lib()
L.vnfr_ig_debug.argtypes = [C.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device=dev)
for i in idx:
    arr = (_lib.Op * 1)(plan.ol.ops[i])
    c = plan.ol.ops[i].conv
    for _ in range(3):
        _lib.call("vnfr_run_ops", arr, 1, _lib.stream_ptr())
    buf.zero_()
    L.vnfr_ig_debug(C.c_void_p(buf.data_ptr()))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    _lib.call("vnfr_run_ops", arr, 1, _lib.stream_ptr())
    t1.record()
    torch.cuda.synchronize()
    L.vnfr_ig_debug(None)
    v = buf.cpu().tolist()
    print("op %d: %dx%d s%d cin %d cout %d a_mode %d block_n %d: %.1f us" % (i, c.kh, c.kw, c.stride, c.cin, c.cout, c.a_mode, c.block_n,
                                                                           t0.elapsed_time(t1) * 1e3))
    for k, name in enumerate(NAMES):
        print("   %-16s %10d cyc  %5.1f %%" % (name, v[k], 100.0 * v[k] / max(v[0], 1)))
