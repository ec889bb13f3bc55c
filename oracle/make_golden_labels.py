"""tests/golden/pipeline_1080p_labels.npz: the UNMODIFIED reference's demo_video path on BASELINE config 3 frames.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Run:  python -m oracle.make_golden_labels

8 synthetic 1080p frames (seeds 0..7 = the first frames of bench.py's rank 0 batch, 12 faces each = 96 faces) go through
the reference's parallel_detect_and_align (demo_image.py:273-306) and recognize_celeb's model calls (demo_image.py:50-76:
transforms_default -> InceptionResnetV1 -> MLPModel -> argmax) with the seed-0 random-init encoder / MLP state dicts of
oracle/nets.py.  Stored per face: frame index, box, label, the full log-probability row (so a parity test can print the
reference's own margin for any face whose label differs) and the embedding."""
import os

import numpy as np
import torch

from . import ref_shims, nets, synth

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pipeline_1080p_labels.npz")
N_FRAMES = 8


def main():
    import torchvision, cv2
    ref = ref_shims.load_reference()
    torch.manual_seed(0)
    fr = synth.frames("1080p", N_FRAMES)
    det = ref.models.MTCNN(image_size=160, keep_all=True, device="cpu", min_face_size=50)
    cp = ref.align_face.center_point_dict["(160, 160)"]
    faces, boxes = ref.demo_image.parallel_detect_and_align(list(fr), det, cp, (160, 160))
    enc = ref.models.InceptionResnetV1(pretrained=None, device="cpu").eval()
    enc.load_state_dict(nets.make_encoder_state_dict(seed=0))
    mlp = ref.models.MLPModel(512, 1001).eval()
    mlp.load_state_dict(nets.make_mlp_state_dict(1001, seed=0))
    flat = [f for x in faces for f in x]
    x = torch.stack([ref.data_loader.transforms_default(f) for f in flat])
    with torch.no_grad():
        emb = ref.demo_image.find_embedding(x, enc)
        lp = mlp(emb)
    top2 = torch.topk(lp, 2, dim=1)[0]
    margin = (top2[:, 0] - top2[:, 1]).numpy()
    d = {"provenance": np.array("reference=/root/reference (unmodified, oracle/ref_shims.py) device=cpu torch=%s torchvision=%s numpy=%s "
                                "cv2=%s" % (torch.__version__, torchvision.__version__, np.__version__, cv2.__version__)),
         "count": np.array([len(x) for x in faces], dtype=np.int64),
         "boxes": np.concatenate([np.asarray(b, dtype=np.float32).reshape(-1, 4) for b in boxes]),
         "labels": lp.argmax(1).numpy().astype(np.int64), "logp": lp.numpy().astype(np.float32),
         "emb": emb.numpy().astype(np.float32), "margin": margin.astype(np.float32)}
    np.savez_compressed(PATH, **d)
    print("faces per frame", d["count"].tolist(), "| top-1 minus top-2 log-prob: min %.2e median %.2e" % (margin.min(), np.median(margin)))
    print("sorted margins (first 12):", np.sort(margin)[:12])


if __name__ == "__main__":
    main()
