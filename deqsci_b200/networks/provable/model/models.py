"""17-layer DnCNN — drop-in for the reference's networks/provable/model/models.py:5-22
(`--denoiser DnCNN`, video_sci_proxgrad.py:172-174): conv+ReLU, 15 x (conv + BatchNorm + ReLU), conv;
64 features, 3x3, no bias; state_dict keys `dncnn.N.*`.  It is the BatchNorm variant of the same
Sequential the SimpleCNN mirror builds, so it runs on the same native conv stack (eval-mode BatchNorm
folded to a per-channel affine); the reference ships denoiser-only weights for it
(networks/provable/Pretrained_models/DnCNN_noise*.pth) but no DEQ checkpoint."""
from .SimpleCNN_models import DnCNN as _SequentialDnCNN


class DnCNN(_SequentialDnCNN):
    def __init__(self, channels, num_of_layers=17, tag='denoiser'):
        super().__init__(channels, num_of_layers=num_of_layers, lip=0.0, no_bn=False, adaptive=False, tag=tag)
