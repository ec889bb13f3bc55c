"""Scratch: pyramid kernel time on 64 synthetic 1080p frames."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import _lib, synthetic
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
fr = torch.from_numpy(synthetic.frames("1080p", 8)).to(dev).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous()
p = _lib.Pyramid()
_lib.call("vnfr_pyramid_plan", B, 1080, 1920, 50, 0.709, C.byref(p))
levels = torch.empty(p.level_off[p.n_levels], device=dev)
for _ in range(3):
    _lib.call("vnfr_pyramid_resize_norm", C.byref(p), _lib.ptr(fr), _lib.ptr(levels), _lib.stream_ptr())
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(10):
    _lib.call("vnfr_pyramid_resize_norm", C.byref(p), _lib.ptr(fr), _lib.ptr(levels), _lib.stream_ptr())
t1.record(); torch.cuda.synchronize()
print("pyramid %dx1080p: %.3f ms" % (B, t0.elapsed_time(t1) / 10))
