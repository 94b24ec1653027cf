"""17-layer DnCNN with real spectral normalisation on every conv — drop-in for the reference's
networks/provable/model/realSN_models.py:4-22 (`--denoiser RealSN_DnCNN`, video_sci_proxgrad.py:178-180):
SN-conv + ReLU, 15 x (SN-conv + BatchNorm + ReLU), SN-conv; state_dict keys
`dncnn.N.{weight_orig,weight,weight_u}` and the BatchNorm entries.  Same Sequential as the other DnCNN
mirrors, so inference runs on the native conv stack (stored `weight` buffers, BatchNorm folded)."""
import torch.nn as nn

from .SimpleCNN_models import DnCNN as _SequentialDnCNN
from .Spectral_Normalize_chen import spectral_norm


class DnCNN(_SequentialDnCNN):
    def __init__(self, channels, num_of_layers=17, tag='denoiser'):
        super().__init__(channels, num_of_layers=num_of_layers, lip=0.0, no_bn=False, adaptive=False, tag=tag,
                         _conv_layer=lambda cin, cout: spectral_norm(
                             nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False)))
