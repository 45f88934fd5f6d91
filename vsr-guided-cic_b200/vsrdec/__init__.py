"""Host-side binding of libvsrdec.so (C ABI: include/vsrdec.h) for the role-shift decoder path."""
from ._lib import load_library, library_path, VsrError, EXPORTED_SYMBOLS  # noqa: F401
from .engine import DecoderEngine, PARAM_NAMES  # noqa: F401
from .sharding import shard_range, decode_sharded  # noqa: F401
from .pipeline import DecodePipeline  # noqa: F401
from .preorder import RoleOrderer, permute_slot_index, permute_slot_tiles  # noqa: F401
from .evalflow import EvalFlow  # noqa: F401
