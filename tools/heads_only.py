"""Scratch: run only the detection cascade on 16 synthetic 1080p frames (for ncu captures of the head kernels)."""
import sys, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import synthetic
from vn_celeb_face_recognition_b200.models import MTCNN
dev = torch.device("cuda:0")
det = MTCNN(image_size=160, keep_all=True, min_face_size=50, device=dev)
fr = torch.from_numpy(synthetic.frames("1080p", 16)).to(dev)
for _ in range(2):
    ws = det.detect_device(fr)
torch.cuda.synchronize()
print("faces", int(ws.out_count.sum().item()))
