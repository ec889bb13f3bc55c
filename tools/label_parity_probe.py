"""Where do label differences against the reference come from?  (run on a GPU box)
For the 8 config-3 frames of tests/golden/pipeline_1080p_labels.npz: (a) our encoder + tail on the ORACLE's aligned faces
(identical pixels), (b) the fused pipeline end to end; per face: pixel differences of the aligned crop, cosine to the
reference embedding, label, reference margin."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, golden_encoder_state_dict
from oracle import synth, nets, pipeline as opipe, align
from vn_celeb_face_recognition_b200 import pipeline
from vn_celeb_face_recognition_b200.models import MTCNN, InceptionResnetV1, MLPModel

dev = torch.device("cuda:0")
n_fr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g = load_golden("pipeline_1080p_labels")
fr = synth.frames("1080p", n_fr)
enc = InceptionResnetV1(pretrained=None, device=dev).eval(); enc.load_state_dict(golden_encoder_state_dict())
mlp = MLPModel(512, 1001).to(dev).eval(); mlp.load_state_dict(nets.make_mlp_state_dict(1001, seed=0))
det = MTCNN(image_size=160, keep_all=True, min_face_size=50, device=dev)
fp = pipeline.FacePipeline(det, enc, mlp, (160, 160), "similarity", return_faces_u8=True)
out = fp.run_device(torch.from_numpy(fr).to(dev))
torch.cuda.synchronize()
F = out["n_faces"]
ours_u8 = out["faces_u8"].cpu().numpy()
ours_emb = out["emb"].cpu().numpy(); ours_lab = out["label"].cpu().numpy()
faces, boxes = opipe.parallel_detect_and_align(list(fr), synth.mtcnn_state_dicts(), align.CENTER_POINTS[(160, 160)], (160, 160), min_face_size=50)
flat = np.stack([f for x in faces for f in x])
x = torch.stack([opipe.transforms_default(f) for f in flat]).to(dev)
with torch.no_grad():
    e_same = enc(x); lp_same = mlp(e_same)
e_same = e_same.cpu().numpy(); lab_same = lp_same.argmax(1).cpu().numpy()
gl, ge, gm = g["labels"][:F], g["emb"][:F], g["margin"][:F]
print("faces", F, "oracle faces", len(flat))
print("(a) identical pixels: label flips", int((lab_same != gl).sum()), "min cos %.6f" % (e_same * ge).sum(1).min())
print("(b) end to end:       label flips", int((ours_lab != gl).sum()), "min cos %.6f" % (ours_emb * ge).sum(1).min())
for k in range(F):
    d = np.abs(ours_u8[k].astype(int) - flat[k].astype(int))
    flag = "FLIP" if ours_lab[k] != gl[k] else ""
    if flag or d.max() > 1 or k < 4:
        print("face %2d: pix diff max %3d mean %.4f frac>1 %.4f | cos e2e %.6f same-pix %.6f | margin %.3e %s" % (
            k, d.max(), d.mean(), (d > 1).mean(), (ours_emb[k] * ge[k]).sum(), (e_same[k] * ge[k]).sum(), gm[k], flag))
