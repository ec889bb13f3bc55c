"""B200-native (sm_100a) face detect -> align -> embed -> classify: a drop-in for the hot path of
votnhan/VN_celeb_face_recognition.  ``models`` mirrors the reference's ``models`` package (MTCNN, InceptionResnetV1,
MLPModel); ``pipeline`` mirrors the glue of demo_image.py / find_embedding.py and adds the fused device pipeline.
Everything computes through libvnfr_b200.so (include/vnfr_b200.h); there is no CPU fallback."""
__all__ = ["models", "pipeline"]
