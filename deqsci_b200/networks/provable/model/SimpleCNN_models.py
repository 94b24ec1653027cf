"""DnCNN-style CNN — drop-in for the reference's networks/provable/model/SimpleCNN_models.py:6-61
(`DnCNN(channels, num_of_layers, lip, no_bn, adaptive, tag)`), used by cnn.ckpt (lip=0, 4 layers,
no BN) and rsn_cnn.ckpt (lip=1: spectrally normalised convs).  state_dict keys: `dncnn.N.weight`
or `dncnn.N.{weight_orig,weight,weight_u}`.  At inference forward runs libdeqsci's conv stack."""
import os

import torch
import torch.nn as nn

from ...._lib import DeqsciError
from ....native import NativePlanCache
from ...ffdnet.models import sequential_bn_slots, sequential_live_weights, sequential_to_plan_layers
from .conv_sn_chen import conv_spectral_norm


class DnCNN(nn.Module, NativePlanCache):
    def __init__(self, channels, num_of_layers=17, lip=1.0, no_bn=False, adaptive=False, tag='denoiser',
                 _conv_layer=None):
        super().__init__()
        self.tag = tag
        self.channels = channels
        features = 64
        if lip > 0.0:
            sigmas = [pow(lip, 1.0 / num_of_layers)] * num_of_layers
        else:
            sigmas = [0.0] * num_of_layers
        if adaptive:
            sigmas = [5.0, 2.0, 1.0, 0.681, 0.464, 0.316]
            assert len(sigmas) == num_of_layers, "Length of SN list uncompatible with num of layers."

        def conv_layer(cin, cout, sigma):
            if _conv_layer is not None:          # subclasses with their own conv flavour (realSN_models)
                return _conv_layer(cin, cout)
            conv = nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False)
            return conv_spectral_norm(conv, sigma=sigma) if sigma > 0.0 else conv

        mods = [conv_layer(channels, features, sigmas[0]), nn.ReLU(inplace=True)]
        for i in range(1, num_of_layers - 1):
            mods.append(conv_layer(features, features, sigmas[i]))
            if not no_bn:
                mods.append(nn.BatchNorm2d(features))
            mods.append(nn.ReLU(inplace=True))
        mods.append(conv_layer(features, channels, sigmas[-1]))
        self.dncnn = nn.Sequential(*mods)

    def _plan_layers(self, train=False):
        if self.channels != 1:
            raise DeqsciError("the native DnCNN path covers single-channel frames (the SCI path)")
        return "dncnn", sequential_to_plan_layers(self.dncnn, fold_bn=not train)

    def bn_slots(self):
        return sequential_bn_slots(self.dncnn)

    def _plan_live_weights(self, train=False):
        return sequential_live_weights(self.dncnn, train)

    def native_train_ok(self, z):
        """Train-mode forward solve (no_grad) with BatchNorm on the native kernels (plain conv / BatchNorm /
        ReLU stacks only: the spectral-norm power iteration is per-call state the plan does not express)."""
        from ....native import default_precision
        H, W = int(z.shape[1]), int(z.shape[2])
        plain = all(isinstance(m, (nn.Conv2d, nn.ReLU, nn.BatchNorm2d)) for m in self.dncnn)
        return (z.is_cuda and self.training and not torch.is_grad_enabled() and self.channels == 1 and plain
                and any(isinstance(m, nn.BatchNorm2d) for m in self.dncnn)
                and (getattr(self, "precision", None) or default_precision()) == "tc_split"
                and os.environ.get("DEQSCI_TC_PAIR", "1") != "0"      # the train path lives in the CTA-pair kernel
                and W > 64)

    # -- the stack's VJP on the same kernels (backward solve of the implicit-differentiation hook) ---------
    def native_adjoint_ok(self, z):
        """Plain conv / ReLU stack (no BatchNorm, no spectral-norm hook: J_D^T is then the same stack with transposed,
        flipped weights and the ReLUs replaced by the saved activations' signs), on the CTA-pair / first-layer
        tensor-core kernels: cube [B,H,W,T] wider than 64 pixels, precision tc_split."""
        from ....native import default_precision
        convs = [m for m in self.dncnn if isinstance(m, nn.Conv2d)]
        return (z.is_cuda and self.channels == 1 and self._stateless_in_train_mode() and len(convs) >= 3
                and all(c.bias is None and tuple(c.weight.shape[2:]) == (3, 3) for c in convs)
                and (getattr(self, "precision", None) or default_precision()) == "tc_split"
                and os.environ.get("DEQSCI_TC_PAIR", "1") != "0" and os.environ.get("DEQSCI_TC_FIRST", "1") != "0"
                and int(z.shape[2]) > 64)

    def _plan_kind(self):
        return "dncnn"

    def _adjoint_conv_weights(self):
        if not all(isinstance(m, (nn.Conv2d, nn.ReLU, nn.BatchNorm2d)) for m in self.dncnn):
            return None
        return [m.weight for m in self.dncnn if isinstance(m, nn.Conv2d)]

    def _stateless_in_train_mode(self):
        return all(isinstance(m, (nn.Conv2d, nn.ReLU)) for m in self.dncnn)

    def uses_native(self, x):
        # eval mode always; train mode only under no_grad for plain conv/ReLU stacks (BatchNorm batch
        # statistics and the spectral-norm power iteration are per-call state the plan does not express)
        if not (x.is_cuda and self.channels == 1):
            return False
        from ....native import graph_needed
        if graph_needed(self, x):
            return False
        return (not self.training) or (not torch.is_grad_enabled() and self._stateless_in_train_mode())

    def forward(self, x):
        if self.uses_native(x):
            N, C, H, W = x.shape
            frames = x.detach().reshape(N, H, W, 1).contiguous()      # cube with T = 1
            out = self.native_plan(x.device).denoise_residual(frames, 0.0)
            return (frames - out).reshape(N, C, H, W)                 # network output (predicted noise)
        if not x.is_cuda and not self.training:
            raise DeqsciError("DnCNN inference on %s: deqsci_b200 has no CPU path" % x.device)
        return self.dncnn(x)
