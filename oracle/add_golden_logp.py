"""Adds the reference classifier's full log-probability vectors to tests/golden/demo_video_small.npz.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Run:  python -m oracle.add_golden_logp

The labels in that golden come from a RANDOM-INIT MLPModel(512, 1001) (BASELINE.json configs use random-init weights), so
the top two classes of a face can be separated by less than the encoder tolerance the north star allows (cosine >= 0.999).
With the reference's own log-probabilities in the golden, the parity test can tell a wrong label from a near-tie: the
predicted label must be the reference's, or a class whose reference log-probability is within the stated tolerance of
the reference's maximum.  The existing arrays of the file are kept byte for byte; the reference is re-run on its own
aligned faces (demo_image.py:50-76: transforms_default -> InceptionResnetV1 -> MLPModel) and must reproduce the stored
labels before anything is written."""
import os

import numpy as np
import torch

from . import ref_shims, nets

PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "demo_video_small.npz")


def main():
    ref = ref_shims.load_reference()
    g = dict(np.load(PATH, allow_pickle=True))
    enc = ref.models.InceptionResnetV1(pretrained=None, device="cpu").eval()
    enc.load_state_dict(nets.make_encoder_state_dict(seed=0))
    mlp = ref.models.MLPModel(512, 1001).eval()
    mlp.load_state_dict(nets.make_mlp_state_dict(1001, seed=0))
    for i in range(2):
        faces = g["aligned_%d" % i]
        x = torch.stack([ref.data_loader.transforms_default(f) for f in faces])
        with torch.no_grad():
            lp = mlp(enc(x))
        assert lp.argmax(1).tolist() == g["labels_%d" % i].tolist(), "the reference does not reproduce its stored labels"
        g["logp_%d" % i] = lp.numpy().astype(np.float32)
        top2 = torch.topk(lp, 2, dim=1)[0]
        print("frame %d: labels %s, top-1 minus top-2 log-prob %s" % (i, g["labels_%d" % i].tolist(), (top2[:, 0] - top2[:, 1]).tolist()))
    np.savez_compressed(PATH, **g)


if __name__ == "__main__":
    main()
