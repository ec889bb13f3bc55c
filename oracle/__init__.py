"""oracle/ -- CPU restatement of the reference's detect -> align -> embed -> classify path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or the reported CPU baseline -- never as the thing
shipped or measured as the product.  The product package (``vn_celeb_face_recognition_b200``) never imports it and
fails loudly when its CUDA extension is missing.

Parity status: PINNED against the reference itself.  The reference (/root/reference) is pure Python and ships no tests
or golden vectors, so the oracle is pinned by running the UNMODIFIED reference (import shims only,
``oracle/ref_shims.py``) in the build container on the bundled images / seeded synthetic frames and comparing, see
``oracle/make_golden.py`` -> ``tests/golden/*.npz`` and ``tests/test_oracle_vs_golden.py``.  Exceptions, marked
"parity unpinned" where they occur: ``skimage.transform.SimilarityTransform`` (absent from the image; restated from
Umeyama 1991 in ``oracle/align.py``).

Modules:
  nets.py      P/R/O-Net, InceptionResnetV1, MLPModel as functions of a state_dict  (models/mtcnn.py, inception_resnet_v1.py, mlp_model.py)
  detect.py    detect_face and helpers, NMS restatements, extract_face               (models/mtcnn_utils/detect_face.py, models/mtcnn.py)
  align.py     Umeyama similarity + cv2.warpAffine restatement, templates            (align_face.py, demo_image.py)
  pipeline.py  parallel_detect_and_align / recognize_celeb / cal_embedding           (demo_image.py, find_embedding.py)
  synth.py     seeded synthetic frames that contain faces, weight loading            
  ref_shims.py import the real reference (build container only)
"""
