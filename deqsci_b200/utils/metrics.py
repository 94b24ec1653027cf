"""PSNR / SSIM exactly as the reference evaluates them (not part of the hot path).

psnr: skimage.metrics.peak_signal_noise_ratio on float arrays, as called at
ref training/sci_equilibrium_training.py:182-183 (data_range 1 when the true image is >= 0, else 2).
ssim: ref pytorch_ssim/__init__.py:7-37,65-73 (11x11 gaussian window, sigma 1.5, mean SSIM map)."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def peak_signal_noise_ratio(image_true, image_test, data_range=None):
    t = np.asarray(image_true, np.float64)
    u = np.asarray(image_test, np.float64)
    if data_range is None:
        data_range = 1.0 if t.min() >= 0 else 2.0
    return float(10 * np.log10(data_range * data_range / np.mean((t - u) ** 2)))


def _window(window_size, channel, like):
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous().to(like)


def ssim(img1, img2, window_size=11, size_average=True):
    """img [N,C,H,W] tensors."""
    channel = img1.shape[1]
    win = _window(window_size, channel, img1)
    p = window_size // 2
    mu1 = F.conv2d(img1, win, padding=p, groups=channel)
    mu2 = F.conv2d(img2, win, padding=p, groups=channel)
    s1 = F.conv2d(img1 * img1, win, padding=p, groups=channel) - mu1 * mu1
    s2 = F.conv2d(img2 * img2, win, padding=p, groups=channel) - mu2 * mu2
    s12 = F.conv2d(img1 * img2, win, padding=p, groups=channel) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)
