"""brainmagick_b200: B200-native (sm_100a) drop-in for brainmagick's contrastive training step --
`SimpleConv` (bm/models/simpleconv.py) and `ClipLoss` (bm/losses.py) behind the reference's module surface."""
from .simpleconv import SimpleConv  # noqa: F401
from .losses import ClipLoss  # noqa: F401
from .common import ChannelMerger, ConvSequence, FourierEmb, PositionGetter, SubjectLayers  # noqa: F401

__all__ = ["SimpleConv", "ClipLoss", "ChannelMerger", "ConvSequence", "FourierEmb", "PositionGetter",
           "SubjectLayers"]
