"""Host-side mirror of the reference's ``models`` package for the hot path (models/__init__.py:1-3 of the reference):
``MTCNN``, ``InceptionResnetV1``, ``MLPModel`` -- same names, constructor arguments, ``state_dict`` keys and return
conventions; every forward runs hand-written sm_100a kernels through the C-ABI of include/vnfr_b200.h."""
from .inception_resnet_v1 import InceptionResnetV1
from .mlp_model import MLPModel
from .mtcnn import MTCNN, fixed_image_standardization, prewhiten
