"""Scratch: per-op device time of one encoder forward (each op timed alone, back to back repeats, CUDA events).
usage: enc_layers.py [batch] [reps]   -> table sorted in execution order + totals by class"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200.models import InceptionResnetV1
from vn_celeb_face_recognition_b200 import _lib, encoder_plan

dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = InceptionResnetV1(device=dev).eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
plan = enc._plan(n, 160, 160, dev)
plan.x0.normal_()
plan.run()
torch.cuda.synchronize()
rows = []
tot = 0.0
for i, op in enumerate(plan.ol.ops):
    arr = (_lib.Op * 1)(op)
    _lib.call("vnfr_run_ops", arr, 1, _lib.stream_ptr())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        _lib.call("vnfr_run_ops", arr, 1, _lib.stream_ptr())
    t1.record()
    torch.cuda.synchronize()
    us = t0.elapsed_time(t1) / reps * 1e3
    c = op.conv
    if op.kind == 3:
        fl = 2.0 * n * 64 * (896 * 256 + 2 * 896 * 128 + 256 * 896)
        rows.append((i, "block17", n * 64, 896, 896, 0, 9, us, fl / us / 1e6, 2.0 * 2 * n * 64 * 896 / us / 1e3))
    elif op.kind == 0:
        M = c.n_img * c.out_h * c.out_w
        K = c.kh * c.kw * c.cin
        fl = 2.0 * M * K * c.cout
        byt = 2.0 * (c.n_img * c.in_h * c.in_w * c.cin + M * c.cout + c.cout * K)
        rows.append((i, "conv%dx%d s%d" % (c.kh, c.kw, c.stride), M, c.cout, K, c.block_n, c.a_mode, us, fl / us / 1e6, byt / us / 1e3))
    else:
        rows.append((i, "pool%d" % op.kind, c.n_img * c.in_h * c.in_w, c.cin, 0, 0, 0, us, 0.0, 0.0))
    tot += us
print("%3s %-12s %9s %5s %5s %4s %2s %9s %8s %8s" % ("#", "op", "M", "N", "K", "bn", "am", "us", "TFLOP/s", "GB/s"))
for r in rows:
    print("%3d %-12s %9d %5d %5d %4d %2d %9.1f %8.1f %8.1f" % r)
print("sum of per-op times: %.1f us" % tot)
cls = {}
for r in rows:
    k = r[1] + (" tma" if r[6] == 1 else "")
    a = cls.setdefault(k, [0.0, 0.0, 0])
    a[0] += r[7]; a[1] += r[8] * r[7]; a[2] += 1
for k, (us, fl, cnt) in sorted(cls.items(), key=lambda kv: -kv[1][0]):
    print("%-16s n=%3d %9.1f us  %5.1f%%  avg %.1f TFLOP/s" % (k, cnt, us, 100 * us / tot, fl / us if us else 0))
