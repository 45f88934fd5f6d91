"""Drop-in `models` package for the role-shift captioning decoder path.

Mirrors the import surface of the reference's models/__init__.py:1-4 so that
coco_scripts/eval_coco.py:6,10 and flickr_scripts/eval_flickr.py:6 keep working when this
directory's parent is placed on sys.path ahead of the reference checkout.  The captioning
decoder and the two networks of the eval pre-step — the R-level SSP network
`SinkhornNet` and the S-level SSP transformer `S_SSP` — are re-implemented B200-native, via
libvsrdec; nothing is forwarded to a reference checkout any more.
"""
from .CaptioningModel import CaptioningModel as _CaptioningModel
from .controllable_captioning import ControllableCaptioningModel
from .sinkhorn_network import SinkhornNet
from .sort_model import S_SSP
