"""Scratch: run selected encoder ops alone inside a cudaProfiler range (for `ncu --profile-from-start off --set full`).
usage: enc_ops_profile.py [batch] idx idx ...     (indices into the op list of tools/enc_layers.py)"""
import sys
import torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200.models import InceptionResnetV1
from vn_celeb_face_recognition_b200 import _lib

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
torch.cuda.profiler.start()
for i in idx:
    arr = (_lib.Op * 1)(plan.ol.ops[i])
    _lib.call("vnfr_run_ops", arr, 1, _lib.stream_ptr())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
