"""torch_tts_b200 -- B200-native (sm_100a) replacement for the VITS2 alignment
hot path of kgoba/torch-tts: the neg_cent cost block of SynthesizerTrn.forward
(vits2/models.py:1224-1256) and monotonic_align.maximum_path
(vits2/monotonic_align/).  Python here is a thin host layer over the C ABI in
include/mas_b200.h; there is no CPU fallback.
"""
from .monotonic_align import maximum_path, maximum_path_compact, lengths_from_mask  # noqa: F401
from .align import align, neg_cent, CompactAlignment  # noqa: F401
from .sharded import shard_bounds, gather_compact, expand_path, align_sharded  # noqa: F401
from .expand import expand_prior, logw, idx_from_durations, generate_path  # noqa: F401
from .plan import AlignPlan  # noqa: F401

__all__ = [
    "maximum_path", "maximum_path_compact", "lengths_from_mask", "align", "neg_cent", "CompactAlignment", "AlignPlan",
    "shard_bounds", "gather_compact", "expand_path", "align_sharded",
    "expand_prior", "logw", "idx_from_durations", "generate_path",
]
