"""Scratch: per-phase cycles of onet_kernel CTA 0 on the bench workload."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import _lib, synthetic
from vn_celeb_face_recognition_b200.models import MTCNN
dev = torch.device("cuda:0")
lib = _lib.lib()
lib.vnfr_heads_debug.argtypes = [C.c_void_p]
det = MTCNN(image_size=160, keep_all=True, min_face_size=50, device=dev)
fr = torch.from_numpy(synthetic.frames("1080p", 16)).to(dev)
det.detect_device(fr); torch.cuda.synchronize()
buf = torch.zeros(32, dtype=torch.int64, device=dev)
lib.vnfr_heads_debug(C.c_void_p(buf.data_ptr()))
ws = det.detect_device(fr); torch.cuda.synchronize()
lib.vnfr_heads_debug(C.c_void_p(0))
names = ["locate", "crop", "conv1", "pool1", "conv2", "pool2", "conv3", "pool3", "conv4", "fc+heads"]
b = buf.tolist()
tot = sum(b[:10])
n3 = int(ws.s3_count.sum().item())
print("onet crops total", n3, "per CTA ~", n3 / 148.0)
for n, v in zip(names, b):
    print("%-9s %9d cyc  %5.1f%%" % (n, v, 100.0 * v / max(tot, 1)))

n2 = int(ws.s2_count.sum().item())
print("rnet crops total", n2, "per CTA ~", n2 / 148.0)
rn = ["locate", "crop", "conv1", "pool1", "conv2", "pool2", "conv3", "fc"]
rt = sum(b[16:24])
for n, v in zip(rn, b[16:24]):
    print("%-9s %9d cyc  %5.1f%%" % (n, v, 100.0 * v / max(rt, 1)))
