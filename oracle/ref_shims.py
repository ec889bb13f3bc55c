"""Import the UNMODIFIED reference (/root/reference) with the minimum shims that make its hot path importable.

TEST INFRASTRUCTURE ONLY.  This module is used in the build container to (a) validate the restatement in
``oracle/*.py`` against the real reference and (b) generate the golden vectors under ``tests/golden/`` (see
``oracle/make_golden.py``).  ``/root/reference`` does not exist on the GPU box, so nothing under ``tests/ -m gpu``,
``bench.py`` or ``__graft_entry__.smoke()`` imports this file.

Shims (none of them touches arithmetic on the path):
  * ``torchvision.models.utils`` was removed from torchvision -> fake module (``models/iresnet_encoder.py:4``).
  * ``np.array(ragged)`` raises under numpy >= 1.24 (``detect_face.py:183``, ``mtcnn.py:345-347``) -> the ``np`` global of
    those two modules is replaced by a proxy whose ``array()`` falls back to a 1-D object array.
  * ``face_alignment`` / ``imgaug`` / ``pafy`` / ``matplotlib`` / ``skimage.io`` are absent -> MagicMock modules.
  * ``skimage.transform.SimilarityTransform`` is absent -> Umeyama-with-scale restatement (``oracle.align``).
  * ``torchvision.transforms.RandomRotation(resample=...)`` keyword was removed -> swallow it.
"""
import os
import sys
import types
import importlib
from unittest import mock

import numpy as np
import torch

def _find_reference_root():
    """VNFR_REFERENCE_ROOT, else /root/reference (build container), else <repo>/baseline/_ref (a tree a user placed there)."""
    cands = [os.environ.get("VNFR_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "models")) and os.path.exists(os.path.join(c, "demo_image.py")):
            return c
    return cands[0] or "/root/reference"


REF_ROOT = _find_reference_root()


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


class _NpProxy:
    """numpy proxy: ``array`` tolerates ragged nested sequences the way numpy < 1.24 did."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def array(self, obj, *a, **k):
        try:
            return self._real.array(obj, *a, **k)
        except ValueError:
            out = self._real.empty(len(obj), dtype=object)
            for i, o in enumerate(obj):
                out[i] = o
            return out


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's ``models``, ``detect_face`` module, ``demo_image``, ``align_face``,
    ``data_loader`` and ``find_embedding`` modules."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    # 1. torchvision.models.utils
    import torchvision  # noqa
    fake = types.ModuleType("torchvision.models.utils")
    fake.load_state_dict_from_url = torch.hub.load_state_dict_from_url
    sys.modules.setdefault("torchvision.models.utils", fake)

    # 2. absent third-party modules
    for name in ["face_alignment", "imgaug", "imgaug.augmenters", "pafy", "matplotlib", "matplotlib.pyplot",
                 "skimage.io"]:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()

    # 3. skimage.transform.SimilarityTransform
    try:
        import skimage.transform  # noqa
    except Exception:
        from oracle.align import SimilarityTransform
        sk = types.ModuleType("skimage")
        skt = types.ModuleType("skimage.transform")
        skt.SimilarityTransform = SimilarityTransform
        sk.transform = skt
        sk.io = sys.modules["skimage.io"]
        sys.modules["skimage"] = sk
        sys.modules["skimage.transform"] = skt

    # 4. RandomRotation(resample=...)
    import torchvision.transforms as tvt
    if not getattr(tvt.RandomRotation, "_vnfr_patched", False):
        _orig = tvt.RandomRotation

        class RandomRotation(_orig):
            _vnfr_patched = True

            def __init__(self, *a, resample=None, **k):
                super().__init__(*a, **k)

        tvt.RandomRotation = RandomRotation

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # The reference's top-level packages are called ``models``, ``utils`` ... ; make sure none of ours shadows them.
    for clash in ["models", "utils", "data_loader", "find_embedding", "demo_image", "align_face"]:
        if clash in sys.modules and not getattr(sys.modules[clash], "__file__", "").startswith(REF_ROOT):
            del sys.modules[clash]

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        models = importlib.import_module("models")
        df_mod = sys.modules["models.mtcnn_utils.detect_face"]
        mtcnn_mod = sys.modules["models.mtcnn"]
        df_mod.np = _NpProxy(np)
        mtcnn_mod.np = _NpProxy(np)
        data_loader = importlib.import_module("data_loader")
        data_loader.transforms = data_loader.transforms_default
        align_face = importlib.import_module("align_face")
        demo_image = importlib.import_module("demo_image")
        find_embedding = importlib.import_module("find_embedding")

    _loaded.update(models=models, detect_face_mod=df_mod, mtcnn_mod=mtcnn_mod, data_loader=data_loader,
                   align_face=align_face, demo_image=demo_image, find_embedding=find_embedding)
    return types.SimpleNamespace(**_loaded)
