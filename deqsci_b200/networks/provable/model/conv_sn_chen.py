"""Real spectral normalisation of a conv layer (Ryu et al.), state-compatible with the reference's
networks/provable/model/conv_sn_chen.py:16-93: parameter `weight_orig`, buffers `weight` and
`weight_u` ([1, Cout in {1,64}, 40, 40] power-iteration probe).

eval mode (inference, the hot path): the stored `weight` buffer is used as is (reference :65-67) —
that buffer is what the native conv plan packs.
train mode: one power iteration per forward on the real conv operator (reference :29-50)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _unit(t, eps):
    n = float(torch.sqrt(torch.sum(t * t)))
    return t / max(n, eps)


class SpectralNormConv2d(nn.Module):
    """3x3, padding 1, bias-free conv whose weight is weight_orig / sigma_hat * sigma."""

    def __init__(self, cin, cout, sigma=1.0, n_power_iterations=1, eps=1e-12):
        super().__init__()
        if n_power_iterations <= 0:
            raise ValueError('Expected n_power_iterations to be positive, but got n_power_iterations={}'.format(
                n_power_iterations))
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.padding = (3, 3), (1, 1)
        self.sigma, self.n_power_iterations, self.eps = sigma, n_power_iterations, eps
        proto = nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False)      # same init as the reference
        self.weight_orig = nn.Parameter(proto.weight.detach().clone())
        self.register_buffer("weight", self.weight_orig.detach().clone())
        probe_c = 1 if cout == 1 else 64
        self.register_buffer("weight_u", _unit(torch.randn(1, probe_c, 40, 40), eps))
        self.bias = None

    def plan_weight(self):
        return self.weight

    def _power_iteration(self):
        W = self.weight_orig
        u = self.weight_u
        with torch.no_grad():
            for _ in range(self.n_power_iterations):
                # v = W^T u (adjoint conv = conv with flipped, transposed kernel), u = W v
                v = _unit(F.conv2d(u.flip(2, 3), W.permute(1, 0, 2, 3), padding=1), self.eps).flip(2, 3)
                u = _unit(F.conv2d(v, W, padding=1), self.eps)
        cur_sigma = torch.sum(u * F.conv2d(v, W, padding=1))
        return W / cur_sigma * self.sigma, u

    def forward(self, x):
        if self.training:
            w, u = self._power_iteration()
            self.weight_u = u.detach()
            self.weight = w.detach()
            return F.conv2d(x, w, padding=1)
        return F.conv2d(x, self.weight, padding=1)


def conv_spectral_norm(module, name='weight', sigma=1.0, n_power_iterations=1, eps=1e-12, dim=None):
    """Reference-compatible factory: takes a bias-free 3x3 nn.Conv2d and returns the spectrally
    normalised layer (its weight becomes `weight_orig`)."""
    sn = SpectralNormConv2d(module.in_channels, module.out_channels, sigma, n_power_iterations, eps)
    with torch.no_grad():
        sn.weight_orig.copy_(module.weight)
        sn.weight.copy_(module.weight)
    return sn
