"""FFDNet's input / output rearrangements (reference networks/ffdnet/functions.py:16-53,63-81) as
differentiable torch expressions.  Only the autograd (training) path uses them: at inference both are
folded into the first / last conv kernels of libdeqsci (conv_cc.cu)."""
import torch
import torch.nn.functional as F


def concatenate_input_noise_map(input, noise_sigma):
    """[N,C,H,W], sigma [N] -> [N, C + 4C, H/2, W/2]: C constant noise-map channels, then the 2x2
    pixel-unshuffle of the input.  The reference writes sub-image idx = 2r+c of input channel k to
    channel 4k+idx, which is exactly torch's pixel_unshuffle order."""
    N, C, H, W = input.shape
    noise_map = noise_sigma.reshape(N, 1, 1, 1).to(input.dtype).expand(N, C, H // 2, W // 2)
    return torch.cat((noise_map, F.pixel_unshuffle(input, 2)), 1)


def upsamplefeatures(input):
    """[N,4C,H,W] -> [N,C,2H,2W], inverse of the unshuffle above (differentiable)."""
    return F.pixel_shuffle(input, 2)
