"""brainmagick_b200: B200-native (sm_100a) drop-in for brainmagick's contrastive training step --
`SimpleConv` (bm/models/simpleconv.py) and `ClipLoss` (bm/losses.py) behind the reference's module surface -- and for the
callers either side of it: batch preparation (`norm`: bm/norm.py), the DeepMel feature model (`features`:
bm/models/features.py) and the retrieval evaluation (`retrieval`: scripts/run_eval_probs.py, bm/wer.py)."""
from .simpleconv import SimpleConv  # noqa: F401
from .losses import ClipLoss  # noqa: F401
from .common import ChannelMerger, ConvSequence, FourierEmb, PositionGetter, SubjectLayers  # noqa: F401
from .features import DeepMel  # noqa: F401
from .norm import BatchScaler, ScaleReject  # noqa: F401

__all__ = ["SimpleConv", "ClipLoss", "ChannelMerger", "ConvSequence", "FourierEmb", "PositionGetter",
           "SubjectLayers", "DeepMel", "BatchScaler", "ScaleReject"]
