"""Scratch: BASELINE config 4 (crowded 3840x2160 frames, ~50 faces, min_face_size 20) through the fused pipeline."""
import sys, time, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200 import pipeline, synthetic
from vn_celeb_face_recognition_b200.models import MTCNN, InceptionResnetV1, MLPModel
dev = torch.device("cuda:0")
torch.manual_seed(0)
det = MTCNN(image_size=160, keep_all=True, min_face_size=20, device=dev)
enc = InceptionResnetV1(pretrained=None, device=dev).eval(); enc.chunk = 1024
cls = MLPModel(512, 1001).to(dev).eval()
fp = pipeline.FacePipeline(det, enc, cls, (160, 160), "similarity", max_faces_per_frame=96)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fr = torch.from_numpy(synthetic.frames("4k", B)).to(dev)
for _ in range(2):
    out = fp.run_device(fr)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(3):
    out = fp.run_device(fr)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 3
ws = out["ws"]
print("4K x %d frames: %.2f ms/step, %d faces (%.1f per frame), %.0f faces/s, %.1f frames/s" % (
    B, ms, out["n_faces"], out["n_faces"] / B, out["n_faces"] / ms * 1e3, B / ms * 1e3))
print("rnet crops/frame %.0f, onet crops/frame %.0f" % (ws.s2_count.sum().item() / B, ws.s3_count.sum().item() / B))
