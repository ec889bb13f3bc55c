"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, import shims only).

Run in the build container only:  python -m oracle.make_golden
Provenance is recorded inside every file (torch / torchvision / numpy / cv2 versions, device = CPU).
"""
import os
import sys

import numpy as np
import torch

from . import ref_shims, synth, nets, align

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _prov():
    import torchvision, cv2
    return np.array("reference=/root/reference (unmodified, oracle/ref_shims.py) device=cpu torch=%s torchvision=%s numpy=%s "
                    "cv2=%s" % (torch.__version__, torchvision.__version__, np.__version__, cv2.__version__))


def _ragged(prefix, lst, d):
    d[prefix + "_n"] = np.array([len(x) for x in lst], dtype=np.int64)
    for i, x in enumerate(lst):
        d["%s_%d" % (prefix, i)] = np.asarray(x, dtype=np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.load_reference()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---------------- detection: the reference's MTCNN.detect (mtcnn.py:278-361)
    for name, kw, frames in [
        ("detect_small_min50", dict(min_face_size=50), synth.frames("small", 3)),
        ("detect_small_min20", dict(min_face_size=20), synth.frames("small", 2, first_seed=7)),
        ("detect_1080p_min50", dict(min_face_size=50), synth.frames("1080p", 2)),
        ("detect_4k_min20", dict(min_face_size=20), synth.frames("4k", 1)),
    ]:
        m = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", **kw)
        d = {"provenance": _prov(), "min_face_size": np.array(kw["min_face_size"])}
        for sl in (True, False):
            m.select_largest = sl
            b, p, l = m.detect(frames, landmarks=True)
            tag = "largest" if sl else "prob"
            _ragged("boxes_" + tag, b, d); _ragged("probs_" + tag, p, d); _ragged("points_" + tag, l, d)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, d["boxes_prob_n"])

    # the 20 bundled PNGs, one image at a time (sizes differ), min_face_size 50 (cfg/detection/mtcnn.json)
    m = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", min_face_size=50)
    d = {"provenance": _prov()}
    bl, pl, ll = [], [], []
    for _, img in synth.bundled_faces():
        b, p, l = m.detect(img, landmarks=True)
        bl.append(b); pl.append(p); ll.append(l)
    _ragged("boxes", bl, d); _ragged("probs", pl, d); _ragged("points", ll, d)
    np.savez_compressed(os.path.join(OUT, "detect_bundled.npz"), **d)
    print("bundled", d["boxes_n"])

    # ---------------- MTCNN.forward / extract (mtcnn.py:229-276, 458-509) for ndarray and Tensor inputs
    fr = synth.frames("small", 2)
    m = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", min_face_size=50)
    d = {"provenance": _prov()}
    faces_np, boxes_np = m(fr)                       # ndarray input -> cv2.INTER_AREA crops
    faces_t, _ = m(torch.from_numpy(fr))             # Tensor input -> area + .byte()
    for i in range(2):
        d["faces_ndarray_%d" % i] = faces_np[i].numpy()
        d["faces_tensor_%d" % i] = faces_t[i].numpy()
        d["boxes_%d" % i] = np.asarray(boxes_np[i], dtype=np.float32)
    m2 = ref.models.MTCNN(image_size=160, margin=14, keep_all=False, device="cpu", min_face_size=50)
    f2, b2, p2 = m2(torch.from_numpy(fr), return_prob=True)
    d["faces_margin14_single"] = torch.stack(f2).numpy(); d["probs_margin14_single"] = np.asarray(p2, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "extract_small.npz"), **d)

    # ---------------- encoder + MLP (inception_resnet_v1.py:272-303, mlp_model.py:10-15)
    enc_sd = nets.make_encoder_state_dict(seed=0)
    mlp_sd = nets.make_mlp_state_dict(1001, seed=0)
    enc = ref.models.InceptionResnetV1(pretrained=None, device="cpu").eval()
    enc.load_state_dict(enc_sd)
    mlp = ref.models.MLPModel(512, 1001).eval()
    mlp.load_state_dict(mlp_sd)
    x = synth.crops_160(8, seed=1)
    with torch.no_grad():
        e = enc(x)
        lp = mlp(e)
        x112 = torch.nn.functional.interpolate(x[:2], size=(112, 112), mode="bilinear", align_corners=False)
        e112 = enc(x112)
    d = {"provenance": _prov(), "emb": e.numpy(), "logp_argmax": lp.argmax(1).numpy(), "logp_max": lp.max(1)[0].numpy(),
         "logp_head": lp[:, :16].numpy(), "emb112": e112.numpy()}
    bn = {k: v.numpy() for k, v in enc_sd.items() if "running_" in k}
    d["bn_keys"] = np.array(sorted(bn))
    d["bn_vals"] = np.concatenate([bn[k].ravel() for k in sorted(bn)])
    np.savez_compressed(os.path.join(OUT, "encoder_seed0.npz"), **d)
    cos = torch.nn.functional.cosine_similarity(e[:, None], e[None], dim=2)
    print("encoder: pairwise cosine min/max offdiag", float(cos.min()), float((cos - torch.eye(8)).max()),
          "labels", lp.argmax(1).tolist())

    # ---------------- demo_video path: parallel_detect_and_align + recognize_celeb (demo_image.py:273-306, 50-76)
    import pandas as pd
    fr = synth.frames("small", 2, first_seed=3)
    det = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", min_face_size=50)
    cp = ref.align_face.center_point_dict["(160, 160)"]
    faces, boxes = ref.demo_image.parallel_detect_and_align(list(fr), det, cp, (160, 160))
    name_df = pd.DataFrame({"label": np.arange(1001), "name": ["id%d" % i for i in range(1001)]})
    names = ref.demo_image.recognize_celeb(faces, "cpu", enc, mlp, ref.data_loader.transforms_default, name_df, 0.0)
    d = {"provenance": _prov()}
    for i in range(2):
        d["aligned_%d" % i] = np.stack(faces[i]) if faces[i] else np.zeros((0, 160, 160, 3), np.uint8)
        d["boxes_%d" % i] = np.asarray(boxes[i], dtype=np.float32)
        d["labels_%d" % i] = np.array([int(n[2:]) for n in names[i]], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "demo_video_small.npz"), **d)
    print("demo_video labels", names)

    # ---------------- find_embedding.cal_embedding (find_embedding.py:45-59) on the bundled PNGs, Resize(160)
    import tempfile, torchvision.transforms as tf
    tr = tf.Compose([tf.Resize(160), ref.data_loader.transforms_default])
    with tempfile.TemporaryDirectory() as td:
        ref.find_embedding.cal_embedding(os.path.join(synth.ASSETS, "faces"), 64, enc, tr, td, "cpu")
        files = sorted(os.listdir(td))
        embs = np.stack([np.load(os.path.join(td, f))["arr_0"] for f in files])
    np.savez_compressed(os.path.join(OUT, "find_embedding_bundled.npz"), provenance=_prov(), files=np.array(files),
                        emb=embs)
    print("find_embedding", len(files), embs.shape)


if __name__ == "__main__":
    main()
