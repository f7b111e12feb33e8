"""User-facing model classes with the reference API (reference models.py:22-45, :134-224)."""
from math import floor
from typing import Callable, Optional

import torch
from torch import Generator, Tensor, nn

from .components import AppendChannelsPlugin, MelSpectrogram
from .diffusion import VDiffusion, VSampler
from .unet import UNetV0
from .utils import default, downsample, groupby, randn_like, upsample


class DiffusionModel(nn.Module):
    """reference models.py:22-45: `net_t`, `diffusion_t`, `sampler_t` plugin slots; kwargs with
    the `diffusion_` / `sampler_` prefixes are routed to those, the rest builds the net, and
    the same net object is shared by all three."""

    def __init__(self, net_t: Callable = UNetV0, diffusion_t: Callable = VDiffusion,
                 sampler_t: Callable = VSampler, loss_fn: Callable = torch.nn.functional.mse_loss,
                 dim: int = 1, **kwargs):
        super().__init__()
        diffusion_kwargs, kwargs = groupby("diffusion_", kwargs)
        sampler_kwargs, kwargs = groupby("sampler_", kwargs)
        self.net = net_t(dim=dim, **kwargs)
        self.diffusion = diffusion_t(net=self.net, loss_fn=loss_fn, **diffusion_kwargs)
        self.sampler = sampler_t(net=self.net, **sampler_kwargs)

    def forward(self, *args, **kwargs) -> Tensor:
        return self.diffusion(*args, **kwargs)

    @torch.no_grad()
    def sample(self, *args, **kwargs) -> Tensor:
        return self.sampler(*args, **kwargs)


class DiffusionUpsampler(DiffusionModel):
    """reference models.py:134-165"""

    def __init__(self, in_channels: int, upsample_factor: int, net_t: Callable = UNetV0, **kwargs):
        self.upsample_factor = upsample_factor
        super().__init__(net_t=AppendChannelsPlugin(net_t, channels=in_channels),
                         in_channels=in_channels, **kwargs)

    def reupsample(self, x: Tensor) -> Tensor:
        return upsample(downsample(x.clone(), factor=self.upsample_factor),
                        factor=self.upsample_factor)

    def forward(self, x: Tensor, *args, **kwargs) -> Tensor:
        return super().forward(x, *args, append_channels=self.reupsample(x), **kwargs)

    @torch.no_grad()
    def sample(self, downsampled: Tensor, generator: Optional[Generator] = None, **kwargs) -> Tensor:
        reupsampled = upsample(downsampled, factor=self.upsample_factor)
        noise = randn_like(reupsampled, generator=generator)
        return super().sample(noise, append_channels=reupsampled, **kwargs)


class DiffusionVocoder(DiffusionModel):
    """reference models.py:168-224: audio channels are folded into the batch; the mel
    spectrogram is unrolled to a waveform-rate conditioning channel by a transposed conv."""

    def __init__(self, net_t: Callable = UNetV0, mel_channels: int = 80, mel_n_fft: int = 1024,
                 mel_hop_length: Optional[int] = None, mel_win_length: Optional[int] = None,
                 in_channels: int = 1, **kwargs):
        mel_hop_length = default(mel_hop_length, floor(mel_n_fft) // 4)
        mel_win_length = default(mel_win_length, mel_n_fft)
        mel_kwargs, kwargs = groupby("mel_", kwargs)
        super().__init__(net_t=AppendChannelsPlugin(net_t, channels=1), in_channels=1, **kwargs)
        self.to_spectrogram = MelSpectrogram(n_fft=mel_n_fft, hop_length=mel_hop_length,
                                             win_length=mel_win_length,
                                             n_mel_channels=mel_channels, **mel_kwargs)
        self.to_flat = nn.ConvTranspose1d(in_channels=mel_channels, out_channels=1,
                                          kernel_size=mel_win_length, stride=mel_hop_length,
                                          padding=(mel_win_length - mel_hop_length) // 2,
                                          bias=False)

    def forward(self, x: Tensor, *args, **kwargs) -> Tensor:
        spec = self.to_spectrogram(x)                                   # [b, c, f, l]
        flat = self.to_flat(spec.reshape(-1, *spec.shape[-2:]))         # [(b c), 1, t]
        x = x.reshape(-1, 1, x.shape[-1])
        return super().forward(x, *args, append_channels=flat, **kwargs)

    @torch.no_grad()
    def sample(self, spectrogram: Tensor, generator: Optional[Generator] = None, **kwargs) -> Tensor:
        lead = spectrogram.shape[:-2]
        flat = self.to_flat(spectrogram.reshape(-1, *spectrogram.shape[-2:]))
        noise = randn_like(flat, generator=generator)
        wave = super().sample(noise, append_channels=flat, **kwargs)
        return wave.reshape(*lead, wave.shape[-1])
