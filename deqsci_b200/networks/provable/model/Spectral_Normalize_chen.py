"""Real spectral normalisation, the variant behind `--denoiser RealSN_DnCNN` — state-compatible with the
reference's networks/provable/model/Spectral_Normalize_chen.py:24-118: parameter `weight_orig`, buffers
`weight` and `weight_u` ([1, Cout in {1,64}, 40, 40]).  Differences from conv_sn_chen.py that matter for
parity: the adjoint step is a full correlation (padding 2) cropped by one pixel (:59-60), and the
normalised weight carries the fixed factor 0.3**(1/17) (:70).

eval mode (inference): the stored `weight` buffer is used as is (:87-89) and is what the native plan packs."""
import torch
import torch.nn.functional as F

from .conv_sn_chen import SpectralNormConv2d, _unit

LIP_FACTOR = pow(0.3, 1.0 / 17.0)


class RealSNConv2d(SpectralNormConv2d):
    def __init__(self, cin, cout, n_power_iterations=1, eps=1e-12):
        super().__init__(cin, cout, sigma=1.0, n_power_iterations=n_power_iterations, eps=eps)

    def _power_iteration(self):
        W = self.weight_orig
        u = self.weight_u
        with torch.no_grad():
            for _ in range(self.n_power_iterations):
                v = _unit(F.conv2d(u.flip(2, 3), W.permute(1, 0, 2, 3), padding=2), self.eps).flip(2, 3)[:, :, 1:-1, 1:-1]
                u = _unit(F.conv2d(v, W, padding=1), self.eps)
        sigma = torch.sum(u * F.conv2d(v, W, padding=1))
        return W / sigma * LIP_FACTOR, u


def spectral_norm(module, name='weight', n_power_iterations=1, eps=1e-12, dim=None):
    """Reference-compatible factory: a bias-free 3x3 nn.Conv2d -> the spectrally normalised layer."""
    sn = RealSNConv2d(module.in_channels, module.out_channels, n_power_iterations, eps)
    with torch.no_grad():
        sn.weight_orig.copy_(module.weight)
        sn.weight.copy_(module.weight)
    return sn
