"""Fused tail kernel alone (n faces): time per launch + phase timestamps of CTA 0."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vn_celeb_face_recognition_b200 import _lib, tail
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
g = torch.Generator().manual_seed(0)
x8 = torch.relu(torch.randn(n, 9, 1792, generator=g)).half().to(dev)
mk = lambda N, K: tail.SplitLinear(torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g) * 0.1, dev)
layers = [(mk(512, 1792), "l2norm"), (mk(2048, 512), "relu"), (mk(1001, 2048), "logsoftmax")]
plan = tail.TailPlan(layers, n, dev, in_mode=0)
emb = torch.empty(n, 512, device=dev); lab = torch.empty(n, dtype=torch.int64, device=dev); pr = torch.empty(n, device=dev)
run = lambda: plan.run(n, x=x8, out_vecs=[emb, None, None], label=lab, prob=pr, n_classes=1001)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("n=%d: %.1f us per launch; split_k %s" % (n, e0.elapsed_time(e1) * 50, [plan.op.layer[l].split_k for l in range(3)]))
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
fn = _lib.lib().vnfr_tail_debug; fn.argtypes = [C.c_void_p]; fn.restype = C.c_int
fn(C.c_void_p(dbg.data_ptr())); run(); torch.cuda.synchronize(); fn(C.c_void_p(0))
t = dbg.cpu().tolist()
names = ["input (pool+split)", "GEMM 1", "row 1 (l2norm)", "GEMM 2", "row 2 (relu)", "GEMM 3", "row 3 (softmax)"]
for i, nm in enumerate(names):
    print("  %-22s %7.1f us" % (nm, (t[i + 1] - t[i]) / 1e3))
