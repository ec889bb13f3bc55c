"""tests/golden/tracker_rows.npz: the tracker CSV rows of demo_video.py, produced by EXECUTING the reference's own lines.

TEST INFRASTRUCTURE ONLY (build container: needs /root/reference).  Run:  python -m oracle.make_golden_tracker

demo_video.py builds the rows inline in its main loop (lines 154-181), so the lines are read from the reference file at
generation time and exec'd with prepared locals (names, boxes, frames_info, frames_queue, args) -- nothing is copied into
the repository; only inputs and the resulting text are stored."""
import os
import textwrap
import types

import numpy as np

from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tracker_rows.npz")


def main():
    src = open(os.path.join(ref_shims.REF_ROOT, "demo_video.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.strip() == "logged_rows = []")
    end = next(i for i, l in enumerate(src) if l.strip() == "str_logged_rows = ''.join(logged_rows)")
    code = textwrap.dedent("\n".join(src[start:end + 1]))
    rng = np.random.RandomState(3)
    h, w = 1080, 1920
    frames_queue = [np.zeros((h, w, 3), np.uint8)] * 3
    frames_info = [[(i + 1) / 29.97, i + 1] for i in range(3)]
    bth_names = [["id12", "Unknown"], [], ["id7"]]
    bth_chosen_boxes = [[(rng.rand(4) * [w, h, w, h]).astype(np.float32) for _ in range(2)], [],
                        [(rng.rand(4) * [w, h, w, h]).astype(np.float32)]]
    env = dict(np=np, bth_names=bth_names, bth_chosen_boxes=bth_chosen_boxes, frames_info=frames_info, frames_queue=frames_queue,
               args=types.SimpleNamespace(recog_emotion=False))
    exec(code, env)
    text = env["str_logged_rows"]
    flat = np.concatenate([np.stack(b) if len(b) else np.zeros((0, 4), np.float32) for b in bth_chosen_boxes])
    np.savez_compressed(OUT, text=np.array(text), boxes=flat, counts=np.array([len(b) for b in bth_chosen_boxes]),
                        names=np.array([n for x in bth_names for n in x]), times=np.array([t for t, _ in frames_info]),
                        numpy_version=np.array(np.__version__), frame_shape=np.array([h, w, 3]),
                        provenance=np.array("demo_video.py lines %d-%d of the unmodified reference, exec'd" % (start + 1, end + 1)))
    print(text)


if __name__ == "__main__":
    main()
