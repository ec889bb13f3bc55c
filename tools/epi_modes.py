import sys, torch
sys.path.insert(0, ".")
from vn_celeb_face_recognition_b200.models import InceptionResnetV1
dev = torch.device("cuda:0")
enc = InceptionResnetV1(device=dev).eval()
plan = enc._plan(768, 160, 160, dev)
for i, op in enumerate(plan.ol.ops):
    if op.kind == 0:
        c = op.conv
        if c.epi_mode == 0 and c.a_mode != 3:
            print(i, "k%dx%d s%d cin %d cout %d a_mode %d bn %d n_split %d pitch0 %d pitch1 %d" % (c.kh, c.kw, c.stride, c.cin, c.cout, c.a_mode, c.block_n, c.n_split, c.out0_pitch, c.out1_pitch))
        if c.a_mode == 3:
            print(i, "sv   k%dx%d cin %d cout %d pitch0 %d n_split %d" % (c.kh, c.kw, c.cin, c.cout, c.out0_pitch, c.n_split))
