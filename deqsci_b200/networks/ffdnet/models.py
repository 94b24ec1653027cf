"""FFDNet — drop-in for the reference's networks/ffdnet/models.py:27-108.

Same constructor, attribute names and state_dict keys (including the upstream spelling
`intermediate_dncnn.itermediate_dncnn.N.*`), so the reference's checkpoints load with strict=True.
The torch layers below are parameter containers; at inference (eval mode or grad disabled) forward
runs the hand-written CUDA conv stack of libdeqsci — pixel-unshuffle + noise map folded into the
first conv, BatchNorm folded into a per-channel affine, pixel-shuffle folded into the last conv.
The layer-by-layer torch evaluation is kept only for the graph-attached training call."""
import os

import torch
import torch.nn as nn

from ..._lib import DeqsciError
from ...native import NativePlanCache
from .functions import concatenate_input_noise_map, upsamplefeatures


class UpSampleFeatures(nn.Module):
    def forward(self, x):
        return upsamplefeatures(x)


class IntermediateDnCNN(nn.Module):
    """conv(in->mid)+ReLU, (num_conv_layers-2) x [conv(mid->mid)+BN+ReLU], conv(mid->out);
    3x3, padding 1, no bias (reference networks/ffdnet/models.py:27-68)."""

    def __init__(self, input_features, middle_features, num_conv_layers):
        super().__init__()
        self.kernel_size = 3
        self.padding = 1
        self.input_features = input_features
        self.num_conv_layers = num_conv_layers
        self.middle_features = middle_features
        if input_features == 5:
            self.output_features = 4      # grayscale
        elif input_features == 15:
            self.output_features = 12     # RGB
        else:
            raise Exception('Invalid number of input features')

        def conv(cin, cout):
            return nn.Conv2d(cin, cout, kernel_size=3, padding=1, bias=False)

        mods = [conv(input_features, middle_features), nn.ReLU(inplace=True)]
        for _ in range(num_conv_layers - 2):
            mods += [conv(middle_features, middle_features), nn.BatchNorm2d(middle_features), nn.ReLU(inplace=True)]
        mods.append(conv(middle_features, self.output_features))
        self.itermediate_dncnn = nn.Sequential(*mods)   # (sic) upstream attribute name = checkpoint key

    def forward(self, x):
        return self.itermediate_dncnn(x)


def sequential_bn_slots(seq):
    """[BatchNorm2d or None] per conv layer of the Sequential (the BatchNorm that follows it)."""
    slots = []
    for mod in seq:
        if isinstance(mod, nn.Conv2d) or hasattr(mod, "plan_weight"):
            slots.append(None)
        elif isinstance(mod, nn.BatchNorm2d):
            slots[-1] = mod
    return slots


def sequential_to_plan_layers(seq, fold_bn=True):
    """nn.Sequential of Conv2d / BatchNorm2d / ReLU -> layer dicts for NativeDenoiser, with
    eval-mode BatchNorm folded to out = conv*scale + bias (computed in fp64, stored fp32).
    fold_bn=False (train plan): BatchNorm layers are left to deqsci_iterate_train."""
    layers = []
    for mod in seq:
        if isinstance(mod, nn.Conv2d) or hasattr(mod, "plan_weight"):
            w = mod.plan_weight() if hasattr(mod, "plan_weight") else mod.weight
            if tuple(w.shape[2:]) != (3, 3) or getattr(mod, "bias", None) is not None:
                raise DeqsciError("native conv stack supports 3x3 bias-free convolutions only")
            layers.append({"weight": w.detach().float().cpu(), "scale": None, "bias": None, "relu": False})
        elif isinstance(mod, nn.BatchNorm2d):
            if not fold_bn:
                continue
            var = mod.running_var.detach().double().cpu()
            mean = mod.running_mean.detach().double().cpu()
            gamma = mod.weight.detach().double().cpu() if mod.affine else torch.ones_like(var)
            beta = mod.bias.detach().double().cpu() if mod.affine else torch.zeros_like(var)
            scale = gamma / torch.sqrt(var + mod.eps)
            layers[-1]["scale"] = scale.float()
            layers[-1]["bias"] = (beta - mean * scale).float()
        elif isinstance(mod, nn.ReLU):
            layers[-1]["relu"] = True
        else:
            raise DeqsciError("native conv stack cannot express layer %r" % (mod,))
    return layers


def sequential_live_weights(seq, train):
    """Conv weights of a plain Conv2d / BatchNorm2d / ReLU stack, for the device-side plan refresh; None
    when the plan also depends on host-folded BatchNorm (eval plans of stacks with BatchNorm) or on
    derived weights (spectral norm)."""
    if not all(isinstance(m, (nn.Conv2d, nn.BatchNorm2d, nn.ReLU)) for m in seq):
        return None
    if any(hasattr(m, "plan_weight") or hasattr(m, "weight_orig") for m in seq):
        return None
    if not train and any(isinstance(m, nn.BatchNorm2d) for m in seq):
        return None
    ws = [m.weight for m in seq if isinstance(m, nn.Conv2d)]
    return ws if all(w.dtype == torch.float32 and w.is_contiguous() for w in ws) else None


class FFDNet(nn.Module, NativePlanCache):
    """FFDNet(num_input_channels, tag); forward(x [N,C,H,W], noise_sigma [N]) -> predicted noise.
    The input is detached before the network, as in the reference (models.py:103-104)."""

    def __init__(self, num_input_channels, tag):
        super().__init__()
        self.num_input_channels = num_input_channels
        self.tag = tag
        if num_input_channels == 1:
            self.num_feature_maps, self.num_conv_layers = 64, 15
            self.downsampled_channels, self.output_features = 5, 4
        elif num_input_channels == 3:
            self.num_feature_maps, self.num_conv_layers = 96, 12
            self.downsampled_channels, self.output_features = 15, 12
        else:
            raise Exception('Invalid number of input features')
        self.intermediate_dncnn = IntermediateDnCNN(input_features=self.downsampled_channels,
                                                    middle_features=self.num_feature_maps,
                                                    num_conv_layers=self.num_conv_layers)
        self.upsamplefeatures = UpSampleFeatures()

    # -- native plan -----------------------------------------------------------------------
    def _plan_layers(self, train=False):
        if self.num_input_channels != 1:
            raise DeqsciError("the native FFDNet path covers the grayscale network (the SCI path); RGB is not on it")
        return "ffdnet", sequential_to_plan_layers(self.intermediate_dncnn.itermediate_dncnn, fold_bn=not train)

    def bn_slots(self):
        return sequential_bn_slots(self.intermediate_dncnn.itermediate_dncnn)

    def _plan_live_weights(self, train=False):
        return sequential_live_weights(self.intermediate_dncnn.itermediate_dncnn, train)

    def _plan_kind(self):
        return "ffdnet"

    def _adjoint_conv_weights(self):
        return [m.weight for m in self.intermediate_dncnn.itermediate_dncnn if isinstance(m, nn.Conv2d)]

    def native_train_ok(self, z):
        """Train-mode forward solve (no_grad) on the native kernels: cube [B,H,W,T] whose half-resolution
        frames are wider than 64 pixels (the CTA-pair conv kernel's tiles)."""
        from ...native import default_precision
        H, W = int(z.shape[1]), int(z.shape[2])
        return (z.is_cuda and self.training and not torch.is_grad_enabled() and self.num_input_channels == 1
                and (getattr(self, "precision", None) or default_precision()) == "tc_split"
                and os.environ.get("DEQSCI_TC_PAIR", "1") != "0"      # the train path lives in the CTA-pair kernel
                and H % 2 == 0 and W % 2 == 0 and W // 2 > 64)

    def uses_native(self, x):
        # train mode means batch-statistics BatchNorm (and running-stat updates) on every call, which
        # the folded-BN plan does not express: the native path is the eval-mode network.
        # (the input is detached before the network, models.py:103-104: only the parameters can need a graph)
        from ...native import graph_needed
        return x.is_cuda and self.num_input_channels == 1 and not self.training and not graph_needed(self)

    def forward(self, x, noise_sigma):
        if self.uses_native(x):
            return self._native_forward(x, noise_sigma)
        if not x.is_cuda and not self.training:
            raise DeqsciError("FFDNet inference on %s: deqsci_b200 has no CPU path" % x.device)
        concat_noise_x = concatenate_input_noise_map(x.detach(), noise_sigma)
        return self.upsamplefeatures(self.intermediate_dncnn(concat_noise_x))

    def _native_forward(self, x, noise_sigma):
        N, C, H, W = x.shape
        plan = self.native_plan(x.device)
        sig = torch.as_tensor(noise_sigma, dtype=torch.float32).reshape(-1).cpu()
        if sig.numel() == 1:
            sig = sig.expand(N)
        frames = x.detach().reshape(N, H, W, 1).contiguous()          # cube with T = 1
        out = torch.empty_like(frames)
        for s in torch.unique(sig).tolist():                          # one launch set per distinct sigma
            idx = torch.nonzero(sig == s).flatten().to(x.device)
            if idx.numel() == N:
                plan.denoise_residual(frames, s, out=out)
            else:
                out[idx] = plan.denoise_residual(frames[idx].contiguous(), s)
        return (frames - out).reshape(N, C, H, W)                     # predicted noise = z - (z - noise)
