"""Plugins around the net (reference components.py:162-236)."""
from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .unet import B200UNet, UNetV0  # noqa: F401  (UNetV0 re-exported like the reference)
from .utils import default


class _Appended(nn.Module):
    """Generic AppendChannelsPlugin for a net that is not the B200 U-Net: concatenate, call."""

    def __init__(self, net: nn.Module):
        super().__init__()
        self.net = net

    def forward(self, x: Tensor, *args, append_channels: Tensor, **kwargs):
        return self.net(torch.cat([x, append_channels], dim=1), *args, **kwargs)


def AppendChannelsPlugin(net_t: Callable, channels: int):
    """reference components.py:162-180.  With the B200 U-Net the concatenation is never
    materialised: the stem kernels read `x` and `append_channels` through two base pointers."""

    def Net(in_channels: int, out_channels: Optional[int] = None, **kwargs) -> nn.Module:
        out_channels = default(out_channels, in_channels)
        if net_t is UNetV0 or net_t is B200UNet:
            return B200UNet(in_channels=in_channels + channels, out_channels=out_channels,
                            append_channels=channels, **kwargs)
        return _Appended(net_t(in_channels=in_channels + channels, out_channels=out_channels,
                               **kwargs))

    return Net


class MelSpectrogram(nn.Module):
    """reference components.py:188-236 (DiffusionVocoder training front-end; runs once per
    call outside the step loop, torchaudio STFT + mel filterbank)."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int, sample_rate: int,
                 n_mel_channels: int, center: bool = False, normalize: bool = False,
                 normalize_log: bool = False):
        super().__init__()
        from torchaudio import transforms
        self.padding = (n_fft - hop_length) // 2
        self.normalize, self.normalize_log, self.hop_length = normalize, normalize_log, hop_length
        self.to_spectrogram = transforms.Spectrogram(
            n_fft=n_fft, hop_length=hop_length, win_length=win_length, center=center, power=None)
        self.to_mel_scale = transforms.MelScale(
            n_mels=n_mel_channels, n_stft=n_fft // 2 + 1, sample_rate=sample_rate)

    def forward(self, waveform: Tensor) -> Tensor:
        lead, t = waveform.shape[:-1], waveform.shape[-1]
        flat = F.pad(waveform.reshape(-1, t), [self.padding] * 2, mode="reflect")
        mel = self.to_mel_scale(torch.abs(self.to_spectrogram(flat)))
        if self.normalize:
            mel = mel / torch.max(mel)
            mel = 2 * torch.pow(mel, 0.25) - 1
        if self.normalize_log:
            mel = torch.log(torch.clamp(mel, min=1e-5))
        return mel.reshape(*lead, *mel.shape[-2:])
