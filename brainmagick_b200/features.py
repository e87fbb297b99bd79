"""Models applied to the features before the contrastive loss (reference: bm/models/features.py).

`DeepMel` is a `ConvSequence` over the Mel spectrogram (conf/feature_model/deep_mel.yaml: 10 layers of 320 channels,
768 out, kernel 3, dilation period 5, BatchNorm, LeakyReLU(0), skip, GLU every 2 layers, no activation on the last
layer).  Its candidates are trained jointly with the brain encoder (bm/solver.py:304-320), so `ClipLoss` back-propagates
into them (`bm_clip_loss_bwd_cand`).  Forward and backward run in CUDA through `convseq.conv_sequence`.
"""
from __future__ import annotations

from .common import ConvSequence


class DeepMel(ConvSequence):
    """n_in_channels -> (n_hidden_layers - 1) x n_hidden_channels -> n_out_channels; `kwargs` go to ConvSequence."""

    def __init__(self, n_in_channels: int, n_hidden_channels: int, n_hidden_layers: int, n_out_channels: int,
                 **kwargs):
        widths = [n_in_channels] + (n_hidden_layers - 1) * [n_hidden_channels] + [n_out_channels]
        super().__init__(widths, **kwargs)
