"""Seeded synthetic inputs for the parity tests: re-export of the package's data generator
(vn_celeb_face_recognition_b200/synthetic.py) so that tests and the oracle share one definition of every config."""
from vn_celeb_face_recognition_b200.synthetic import *  # noqa: F401,F403
from vn_celeb_face_recognition_b200.synthetic import ASSETS, WEIGHTS, _resize_u8  # noqa: F401
