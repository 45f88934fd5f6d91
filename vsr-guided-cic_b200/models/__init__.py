"""Drop-in `models` package for the role-shift captioning decoder path.

Mirrors the import surface of the reference's models/__init__.py:1-4 so that
coco_scripts/eval_coco.py:6,10 and flickr_scripts/eval_flickr.py:6 keep working when this
directory's parent is placed on sys.path ahead of the reference checkout.  Only the
captioning decoder and (first piece) the R-level SSP network `SinkhornNet`
are re-implemented (B200-native, via libvsrdec); the S-level SSP transformer `S_SSP` is out
of this path's scope and is forwarded, unmodified, to a reference checkout named by
$VSR_REFERENCE_ROOT.
"""
import importlib.util
import os
import sys

from .CaptioningModel import CaptioningModel as _CaptioningModel
from .controllable_captioning import ControllableCaptioningModel
from .sinkhorn_network import SinkhornNet

_PASS_THROUGH = {"S_SSP": "sort_model"}


def _load_reference_module(stem):
    root = os.environ.get("VSR_REFERENCE_ROOT")
    if not root:
        raise ImportError(f"models.{stem} is outside the accelerated decoder path; set "
                          f"VSR_REFERENCE_ROOT to a VSR-guided-CIC checkout to use the original")
    name = f"_vsr_reference_models.{stem}"
    if name in sys.modules:
        return sys.modules[name]
    pkg_name = "_vsr_reference_models"
    if pkg_name not in sys.modules:
        pkg_spec = importlib.util.spec_from_loader(pkg_name, loader=None, is_package=True)
        pkg = importlib.util.module_from_spec(pkg_spec)
        pkg.__path__ = [os.path.join(root, "models")]
        sys.modules[pkg_name] = pkg
    # the SSP modules import their siblings as `from models.xxx import ...`
    saved = sys.modules.get("models")
    sys.modules["models"] = sys.modules[pkg_name]
    try:
        mod = importlib.import_module(name)
    finally:
        if saved is not None:
            sys.modules["models"] = saved
    return mod


def __getattr__(attr):
    if attr in _PASS_THROUGH:
        return getattr(_load_reference_module(_PASS_THROUGH[attr]), attr)
    raise AttributeError(attr)
