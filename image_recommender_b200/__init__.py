"""B200-native exact top-k engine behind the image_recommender retrieval hot path.

Importing this package loads libb2k.so (hand-written sm_100a CUDA behind the C ABI in
include/b2k.h).  There is no CPU fallback: a missing extension is an ImportError, a missing
GPU is a B2KError at the first compute call.
"""
from .index import (B2KError, FlatShard, device_count, file_info, load_ids, merge_topk_device,  # noqa: F401
                    normalize_L2, parse_f32_blob)
from .group import ShardGroup  # noqa: F401
from . import _capi  # noqa: F401

__version__ = "0.1.0"
