"""More golden vectors from the UNMODIFIED reference (build container only):  python -m oracle.make_golden_extras

  * tests/golden/extract_small.npz      += faces_pil_{0,1}: MTCNN.forward on PIL images (crop_resize's PIL.BILINEAR path,
                                            detect_face.py:322-323); existing arrays are kept byte for byte
  * tests/golden/select_boxes_small.npz    MTCNN.select_boxes (mtcnn.py:363-456) for all four methods on the reference's own
                                            detections of 3 small frames (PIL inputs: center_weighted_size needs .width)
  * tests/golden/heads_seed0.npz           InceptionResnetV1(classify=True, num_classes=10).forward (inception_resnet_v1.py:
                                            298-300) and identify_person with a per-class threshold dict (demo_image.py:113-147)
"""
import os

import numpy as np
import torch

from . import ref_shims, synth, nets
from .make_golden import _prov, OUT


def main():
    from PIL import Image
    import pandas as pd
    ref = ref_shims.load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    # ---- PIL extraction
    path = os.path.join(OUT, "extract_small.npz")
    g = dict(np.load(path, allow_pickle=False))
    fr = synth.frames("small", 2)
    m = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", min_face_size=50)
    faces_pil, boxes_pil = m([Image.fromarray(f) for f in fr])
    for i in range(2):
        assert np.allclose(np.asarray(boxes_pil[i], np.float32), g["boxes_%d" % i])
        g["faces_pil_%d" % i] = faces_pil[i].numpy()
    np.savez_compressed(path, **g)
    # ---- select_boxes, four methods
    fr3 = synth.frames("small", 3)
    imgs = [Image.fromarray(f) for f in fr3]
    b, p, l = m.detect(imgs, landmarks=True)
    d = {"provenance": _prov()}
    for i in range(3):
        d["boxes_%d" % i], d["probs_%d" % i], d["points_%d" % i] = (np.asarray(b[i], np.float32), np.asarray(p[i], np.float32),
                                                                 np.asarray(l[i], np.float32))
    thr = float(np.median(np.concatenate([np.asarray(x, np.float32) for x in p])))
    d["threshold"] = np.array(thr, np.float32)
    for method in ["probability", "largest", "center_weighted_size", "largest_over_threshold"]:
        sb, sp, spt = m.select_boxes(b, p, l, imgs, method=method, threshold=thr)
        for i in range(3):
            none = sb[i] is None
            d["%s_none_%d" % (method, i)] = np.array(none)
            d["%s_box_%d" % (method, i)] = np.zeros((1, 4), np.float32) if none else np.asarray(sb[i], np.float32)
            d["%s_prob_%d" % (method, i)] = np.zeros(1, np.float32) if none else np.asarray(sp[i], np.float32)
            d["%s_point_%d" % (method, i)] = np.zeros((1, 5, 2), np.float32) if none else np.asarray(spt[i], np.float32)
    # un-batched call (single image): returns (box (1,4), prob scalar, point (1,5,2))
    sb, sp, spt = m.select_boxes(b[0], p[0], l[0], imgs[0], method="probability")
    d["single_box"], d["single_prob"], d["single_point"] = np.asarray(sb, np.float32), np.asarray(sp, np.float32), np.asarray(spt, np.float32)
    np.savez_compressed(os.path.join(OUT, "select_boxes_small.npz"), **d)
    print("select_boxes: faces per frame", [len(x) for x in b], "threshold", thr)
    # ---- classify head + per-class thresholds
    sd = nets.make_encoder_state_dict(seed=0)
    gl = torch.Generator().manual_seed(77)
    sd["logits.weight"] = torch.randn(10, 512, generator=gl) * 0.05
    sd["logits.bias"] = torch.randn(10, generator=gl) * 0.1
    enc = ref.models.InceptionResnetV1(pretrained=None, classify=True, num_classes=10, device="cpu").eval()
    enc.load_state_dict(sd)
    x = synth.crops_160(4, seed=2)
    import warnings
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = enc(x)
    h = {"provenance": _prov(), "classify_logp": out.numpy(), "logits_weight": sd["logits.weight"].numpy(),
         "logits_bias": sd["logits.bias"].numpy()}
    mlp_sd = nets.make_mlp_state_dict(1001, seed=0)
    mlp = ref.models.MLPModel(512, 1001).eval()
    mlp.load_state_dict(mlp_sd)
    emb = torch.from_numpy(np.load(os.path.join(OUT, "encoder_seed0.npz"))["emb"])
    with torch.no_grad():
        probs = torch.exp(mlp(emb)).max(1)[0].numpy()
    rng = np.random.RandomState(5)
    thr_vals = rng.uniform(probs.min() * 0.5, probs.max() * 1.2, size=1001).astype(np.float32)
    thr_dict = {str(i): float(thr_vals[i]) for i in range(1001)}
    name_df = pd.DataFrame({"label": np.arange(1001), "name": ["id%d" % i for i in range(1001)]})
    names = ref.demo_image.identify_person(emb, mlp, name_df, thr_dict)
    h["thr_vals"] = thr_vals
    h["names_per_class_thr"] = np.array(names)
    h["names_scalar_thr"] = np.array(ref.demo_image.identify_person(emb, mlp, name_df, float(np.median(probs))))
    h["scalar_thr"] = np.array(float(np.median(probs)), np.float32)
    h["max_probs"] = probs
    np.savez_compressed(os.path.join(OUT, "heads_seed0.npz"), **h)
    print("identify_person per-class:", names, "scalar:", h["names_scalar_thr"].tolist())


if __name__ == "__main__":
    main()
