"""ORACLE / TEST INFRASTRUCTURE -- not product code.

CPU (or eager-GPU) restatement, in plain PyTorch fp32, of the reference's hot
path: UNetV0 assembly, VDiffusion loss, VSampler loop and the three model
wrappers.  /root/reference does not exist on the GPU box, so this port is what
travels; `oracle/make_golden.py` (run in the build container, where the
reference *is* importable) proves it equal to the unmodified reference files
running on the same `oracle/a_unet` shim and commits fixtures under
tests/golden/.

Every function cites the reference file:line it follows
(paths relative to /root/reference/audio_diffusion_pytorch/).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg and
--impl reference) may import this module.
"""
import math
import os
import sys
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Generator, Tensor, nn

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:  # makes `import a_unet` resolve to oracle/a_unet
    sys.path.insert(0, _HERE)

import a_unet  # noqa: E402
from a_unet import apex  # noqa: E402


# ----------------------------------------------------------------------------- utils.py
def split_prefixed(prefix: str, kwargs: Dict) -> Tuple[Dict, Dict]:
    """utils.py:48-70 (group_dict_by_prefix + groupby, keep_prefix=False)."""
    hit = {k[len(prefix):]: v for k, v in kwargs.items() if k.startswith(prefix)}
    rest = {k: v for k, v in kwargs.items() if not k.startswith(prefix)}
    return hit, rest


def sinc_resample(wave: Tensor, factor_in: int, factor_out: int, rolloff: float = 0.99,
                  lowpass_filter_width: int = 6) -> Tensor:
    """utils.py:82-109: windowed-sinc polyphase resampler (torchaudio-style)."""
    b, c, length = wave.shape
    target = int(factor_out * length / factor_in)
    opts = dict(device=wave.device, dtype=wave.dtype)
    base = min(factor_in, factor_out) * rolloff
    width = math.ceil(lowpass_filter_width * factor_in / base)
    idx = torch.arange(-width, width + factor_in, **opts)[None, None] / factor_in
    t = torch.arange(0, -factor_out, step=-1, **opts)[:, None, None] / factor_out + idx
    t = (t * base).clamp(-lowpass_filter_width, lowpass_filter_width) * math.pi
    window = torch.cos(t / lowpass_filter_width / 2) ** 2
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels = kernels * (window * (base / factor_in))
    flat = F.pad(wave.reshape(b * c, length), (width, width + factor_in))
    out = F.conv1d(flat[:, None], kernels, stride=factor_in)       # [(b c), k, l]
    out = out.reshape(b, c, factor_out, -1).permute(0, 1, 3, 2).reshape(b, c, -1)
    return out[..., :target]


def sinc_downsample(wave: Tensor, factor: int) -> Tensor:      # utils.py:112-113
    return sinc_resample(wave, factor_in=factor, factor_out=1)


def sinc_upsample(wave: Tensor, factor: int) -> Tensor:        # utils.py:116-117
    return sinc_resample(wave, factor_in=1, factor_out=factor)


def cpu_randn_like(t: Tensor, generator: Optional[Generator] = None) -> Tensor:
    """utils.py:123-125: drawn on the CPU generator, then moved to t's device/dtype."""
    return torch.randn(t.shape, generator=generator).to(t)


# ------------------------------------------------------------------------ components.py
def build_unet_v0(dim: int, in_channels: int, channels: Sequence[int], factors: Sequence[int],
                  items: Sequence[int], attentions: Optional[Sequence[int]] = None,
                  cross_attentions: Optional[Sequence[int]] = None,
                  context_channels: Optional[Sequence[int]] = None,
                  attention_features: Optional[int] = None,
                  attention_heads: Optional[int] = None,
                  embedding_features: Optional[int] = None, resnet_groups: int = 8,
                  use_modulation: bool = True, modulation_features: int = 1024,
                  embedding_max_length: Optional[int] = None,
                  use_time_conditioning: bool = True, use_embedding_cfg: bool = False,
                  use_text_conditioning: bool = False,
                  out_channels: Optional[int] = None) -> nn.Module:
    """components.py:34-105 (UNetV0): same defaults, asserts, plugin nesting and item lists."""
    n = len(channels)
    attentions = a_unet.default(attentions, [0] * n)
    cross_attentions = a_unet.default(cross_attentions, [0] * n)
    context_channels = a_unet.default(context_channels, [0] * n)
    per_level = (channels, factors, items, attentions, cross_attentions, context_channels)
    assert all(len(v) == n for v in per_level)                                  # :61

    net_t: Callable = apex.XUNet
    if use_embedding_cfg:                                                        # :66-69
        assert a_unet.exists(embedding_max_length), \
            "use_embedding_cfg requires embedding_max_length"
        net_t = a_unet.ClassifierFreeGuidancePlugin(net_t, embedding_max_length)
    if use_text_conditioning:                                                    # :71-72
        net_t = a_unet.TextConditioningPlugin(net_t)
    if use_time_conditioning:                                                    # :74-76
        assert use_modulation, "use_time_conditioning requires use_modulation=True"
        net_t = a_unet.TimeConditioningPlugin(net_t)

    blocks = []
    for ch, fac, n_items, att, cross, ctx in zip(*per_level):                    # :84-97
        one = ([apex.ResnetItem] + [apex.ModulationItem] * use_modulation
               + [apex.InjectChannelsItem] * (ctx > 0) + [apex.AttentionItem] * att
               + [apex.CrossAttentionItem] * cross)
        blocks.append(apex.XBlock(channels=ch, factor=fac, context_channels=ctx,
                                  items=one * n_items))
    return net_t(dim=dim, in_channels=in_channels, out_channels=out_channels, blocks=blocks,
                 skip_t=apex.SkipModulate if use_modulation else apex.SkipCat,       # :99
                 attention_features=attention_features, attention_heads=attention_heads,
                 embedding_features=embedding_features,
                 modulation_features=modulation_features, resnet_groups=resnet_groups)


def lt_plugin(net_t: Callable, num_filters: int, window_length: int, stride: int) -> Callable:
    """components.py:113-157 (LTPlugin): learned strided filterbank in front of / behind the net."""

    def make(dim: int, in_channels: int, out_channels: Optional[int] = None, **kwargs) -> nn.Module:
        out_channels = a_unet.default(out_channels, in_channels)                     # :121
        wide_in, wide_out = in_channels * num_filters, out_channels * num_filters    # :122-123
        padding = window_length // 2 - stride // 2                                   # :125
        encode = nn.Conv1d(in_channels, wide_in, window_length, stride=stride, padding=padding,
                           padding_mode="reflect", bias=False)                       # :126-135
        decode = nn.ConvTranspose1d(wide_out, out_channels, window_length, stride=stride,
                                    padding=padding, bias=False)                     # :136-143
        net = net_t(dim=dim, in_channels=wide_in, out_channels=wide_out, **kwargs)   # :144-149
        return a_unet.Module([encode, decode, net],
                             lambda x, *a, **kw: decode(net(encode(x), *a, **kw)))    # :151-157

    return make


def append_channels_plugin(net_t: Callable, channels: int) -> Callable:
    """components.py:162-180 (AppendChannelsPlugin)."""

    def make(in_channels: int, out_channels: Optional[int] = None, **kwargs) -> nn.Module:
        net = net_t(in_channels=in_channels + channels,
                    out_channels=a_unet.default(out_channels, in_channels), **kwargs)

        def forward(x: Tensor, *args, append_channels: Tensor, **kw):
            return net(torch.cat([x, append_channels], dim=1), *args, **kw)          # :175

        return a_unet.Module([net], forward)

    return make


# ------------------------------------------------------------------------- diffusion.py
def half_circle(sigmas: Tensor) -> Tuple[Tensor, Tensor]:
    """diffusion.py:77-80 / :167-170 (get_alpha_beta)."""
    angle = sigmas * math.pi / 2
    return torch.cos(angle), torch.sin(angle)


def right_pad_dims(x: Tensor, ndim: int) -> Tensor:
    """diffusion.py:57-59 (extend_dim)."""
    return x.view(*x.shape, *((1,) * (ndim - x.ndim)))


class UniformSigma:
    """diffusion.py:23-30 (UniformDistribution)."""

    def __init__(self, vmin: float = 0.0, vmax: float = 1.0):
        self.vmin, self.vmax = vmin, vmax

    def __call__(self, num_samples: int, device=torch.device("cpu")) -> Tensor:
        return (self.vmax - self.vmin) * torch.rand(num_samples, device=device) + self.vmin


class VDiffusionPort(nn.Module):
    """diffusion.py:68-95 (VDiffusion): RNG order is rand(B) then randn_like(x)."""

    def __init__(self, net: nn.Module, sigma_distribution=None, loss_fn=F.mse_loss):
        super().__init__()
        self.net = net
        self.sigma_distribution = sigma_distribution or UniformSigma()
        self.loss_fn = loss_fn

    def forward(self, x: Tensor, **kwargs) -> Tensor:
        sigmas = self.sigma_distribution(num_samples=x.shape[0], device=x.device)    # :85
        noise = torch.randn_like(x)                                                  # :88
        alpha, beta = half_circle(right_pad_dims(sigmas, x.ndim))                    # :90
        x_noisy = alpha * x + beta * noise                                           # :91
        v_target = alpha * noise - beta * x                                          # :92
        return self.loss_fn(self.net(x_noisy, sigmas, **kwargs), v_target)           # :94-95


class VSamplerPort(nn.Module):
    """diffusion.py:158-190 (VSampler) with LinearSchedule (diffusion.py:142-148)."""

    def __init__(self, net: nn.Module, start: float = 1.0, end: float = 0.0):
        super().__init__()
        self.net, self.start, self.end = net, start, end

    @torch.no_grad()
    def forward(self, x_noisy: Tensor, num_steps: int, show_progress: bool = False,
                **kwargs) -> Tensor:
        b = x_noisy.shape[0]
        sigmas = torch.linspace(self.start, self.end, num_steps + 1, device=x_noisy.device)
        sigmas = sigmas[:, None].expand(-1, b)                                        # :178
        alphas, betas = half_circle(right_pad_dims(sigmas, x_noisy.ndim + 1))         # :179-180
        for i in range(num_steps):                                                    # :183
            v = self.net(x_noisy, sigmas[i], **kwargs)                                # :184
            x_pred = alphas[i] * x_noisy - betas[i] * v                               # :185
            n_pred = betas[i] * x_noisy + alphas[i] * v                               # :186
            x_noisy = alphas[i + 1] * x_pred + betas[i + 1] * n_pred                  # :187
        return x_noisy


class VInpainterPort(nn.Module):
    """diffusion.py:306-354 (VInpainter) with LinearSchedule."""

    def __init__(self, net: nn.Module, start: float = 1.0, end: float = 0.0):
        super().__init__()
        self.net, self.start, self.end = net, start, end

    @torch.no_grad()
    def forward(self, source: Tensor, mask: Tensor, num_steps: int, num_resamples: int,
                show_progress: bool = False, x_noisy: Optional[Tensor] = None, **kwargs) -> Tensor:
        x_noisy = torch.randn_like(source) if x_noisy is None else x_noisy           # :331
        b = x_noisy.shape[0]
        sigmas = torch.linspace(self.start, self.end, num_steps + 1, device=x_noisy.device)
        sigmas = sigmas[:, None].expand(-1, b)                                        # :334
        alphas, betas = half_circle(right_pad_dims(sigmas, x_noisy.ndim + 1))         # :335-336
        for i in range(num_steps):                                                    # :339
            for r in range(num_resamples):                                            # :340
                v = self.net(x_noisy, sigmas[i], **kwargs)                            # :341
                x_pred = alphas[i] * x_noisy - betas[i] * v                           # :342
                n_pred = betas[i] * x_noisy + alphas[i] * v                           # :343
                j = r == num_resamples - 1                                            # :345
                x_noisy = alphas[i + j] * x_pred + betas[i + j] * n_pred              # :346
                s_noisy = alphas[i + j] * source + betas[i + j] * torch.randn_like(source)
                x_noisy = s_noisy * mask + x_noisy * ~mask                            # :350
        return x_noisy


class ARVDiffusionPort(nn.Module):
    """diffusion.py:98-130 (ARVDiffusion): one sigma per split, carried as an extra input channel.
    RNG order: rand((B,1,num_splits)) then randn_like(x)."""

    def __init__(self, net: nn.Module, length: int, num_splits: int, loss_fn=F.mse_loss):
        super().__init__()
        assert length % num_splits == 0, "length must be divisible by num_splits"       # :101
        self.net, self.length, self.num_splits, self.loss_fn = net, length, num_splits, loss_fn
        self.split_length = length // num_splits

    def forward(self, x: Tensor, **kwargs) -> Tensor:
        b, _, t = x.shape
        assert t == self.length, "input length must match length"                       # :116
        sigmas = torch.rand((b, 1, self.num_splits), device=x.device, dtype=x.dtype)    # :118
        sigmas = sigmas.repeat_interleave(self.split_length, dim=2)                     # :119
        noise = torch.randn_like(x)                                                     # :121
        alphas, betas = half_circle(sigmas)                                             # :123
        x_noisy = alphas * x + betas * noise                                            # :124
        v_target = alphas * noise - betas * x                                           # :125
        v_pred = self.net(torch.cat([x_noisy, sigmas], dim=1), **kwargs)                # :127-129
        return self.loss_fn(v_pred, v_target)                                           # :130


class ARVSamplerPort(nn.Module):
    """diffusion.py:193-298 (ARVSampler)."""

    def __init__(self, net: nn.Module, in_channels: int, length: int, num_splits: int):
        super().__init__()
        assert length % num_splits == 0, "length must be divisible by num_splits"       # :196
        self.net, self.in_channels, self.length, self.num_splits = net, in_channels, length, num_splits
        self.split_length = length // num_splits

    @property
    def device(self):
        return next(self.net.parameters()).device

    def sigmas_ladder(self, num_items: int, num_steps_per_split: int) -> Tensor:       # :213-221
        b, n_half, l, k = num_items, self.num_splits // 2, self.split_length, num_steps_per_split
        line = torch.linspace(1, 0, k * n_half, device=self.device)                     # :216
        out = torch.zeros(k + 1, b, 1, n_half * l, device=self.device)                  # (+ index k, :219)
        for step in range(k):                 # "(n i) -> i b 1 (n l)" then flip of the last axis (:217-218)
            for split in range(n_half):
                hi = n_half * l - split * l
                out[step, :, :, hi - l:hi] = line[split * k + step]
        out[-1, :, :, l:] = out[0, :, :, :-l]                                           # :220
        return torch.cat([torch.zeros_like(out), out], dim=-1)                          # :221

    def sample_loop(self, current: Tensor, sigmas: Tensor, **kwargs) -> Tensor:         # :223-238
        alphas, betas = half_circle(sigmas)
        for i in range(sigmas.shape[0] - 1):
            v = self.net(torch.cat([current, sigmas[i]], dim=1), **kwargs)              # :231-232
            x_pred = alphas[i] * current - betas[i] * v                                 # :233
            n_pred = betas[i] * current + alphas[i] * v                                 # :234
            current = alphas[i + 1] * x_pred + betas[i + 1] * n_pred                    # :235
        return current

    def sample_start(self, num_items: int, num_steps: int, **kwargs) -> Tensor:         # :240-247
        b, c, t = num_items, self.in_channels, self.length
        sigmas = torch.linspace(1, 0, num_steps + 1, device=self.device)
        sigmas = sigmas.view(-1, 1, 1, 1).expand(-1, b, 1, t)
        noise = torch.randn((b, c, t), device=self.device) * sigmas[0]
        return self.sample_loop(current=noise, sigmas=sigmas, **kwargs)

    @torch.no_grad()
    def forward(self, num_items: int, num_chunks: int, num_steps: int, start=None,
                show_progress: bool = False, **kwargs) -> Tensor:                       # :250-298
        n = self.num_splits
        assert num_chunks >= n, f"required at least {n} chunks"
        start = self.sample_start(num_items=num_items, num_steps=num_steps, **kwargs)
        if num_chunks == n:
            return start
        assert num_steps >= n, "num_steps must be greater than num_splits"
        sigmas = self.sigmas_ladder(num_items, num_steps // n)
        alphas, betas = half_circle(sigmas)
        start_noise = alphas[0] * start + betas[0] * torch.randn_like(start)            # :278
        chunks = list(start_noise.chunk(chunks=n, dim=-1))
        for _ in range(num_chunks):                                                     # :282
            updated = self.sample_loop(current=torch.cat(chunks[-n:], dim=-1), sigmas=sigmas, **kwargs)
            chunks[-n:] = list(updated.chunk(chunks=n, dim=-1))
            chunks += [torch.randn((num_items, self.in_channels, self.split_length), device=self.device)]
        return torch.cat(chunks[:num_chunks], dim=-1)                                   # :298


# ---------------------------------------------------------------------------- models.py
class DiffusionModelPort(nn.Module):
    """models.py:22-45 (DiffusionModel): one net shared by diffusion and sampler."""

    def __init__(self, net_t: Callable = build_unet_v0, loss_fn=F.mse_loss, dim: int = 1,
                 diffusion_t: Callable = VDiffusionPort, sampler_t: Callable = VSamplerPort, **kwargs):
        super().__init__()
        diffusion_kw, kwargs = split_prefixed("diffusion_", kwargs)                   # :33
        sampler_kw, kwargs = split_prefixed("sampler_", kwargs)                       # :34
        self.net = net_t(dim=dim, **kwargs)                                           # :36
        self.diffusion = diffusion_t(net=self.net, loss_fn=loss_fn, **diffusion_kw)
        self.sampler = sampler_t(net=self.net, **sampler_kw)

    def forward(self, *args, **kwargs) -> Tensor:
        return self.diffusion(*args, **kwargs)

    @torch.no_grad()
    def sample(self, *args, **kwargs) -> Tensor:
        return self.sampler(*args, **kwargs)


class DiffusionAEPort(DiffusionModelPort):
    """models.py:70-131 (DiffusionAE): latent injected at `inject_depth` (InjectChannelsItem)."""

    def __init__(self, in_channels: int, channels: Sequence[int], encoder: nn.Module,
                 inject_depth: int, latent_factor: Optional[int] = None, adapter=None, **kwargs):
        context_channels = [0] * len(channels)
        context_channels[inject_depth] = encoder.out_channels                        # :84-85
        super().__init__(in_channels=in_channels, channels=channels,
                         context_channels=context_channels, **kwargs)
        self.in_channels, self.encoder, self.inject_depth = in_channels, encoder, inject_depth
        self.latent_factor = a_unet.default(latent_factor, encoder.downsample_factor)   # :96
        self.adapter = adapter.requires_grad_(False) if adapter is not None else None

    def forward(self, x: Tensor, with_info: bool = False, **kwargs):                # :99-110
        latent, info = self.encoder(x, with_info=True)
        channels = [None] * self.inject_depth + [latent]
        x = self.adapter.encode(x) if self.adapter is not None else x
        loss = super().forward(x, channels=channels, **kwargs)
        return (loss, info) if with_info else loss

    @torch.no_grad()
    def decode(self, latent: Tensor, generator: Optional[Generator] = None, **kwargs) -> Tensor:
        b = latent.shape[0]                                                          # :116-131
        exponent = math.log2(latent.shape[2] * self.latent_factor)
        lo, hi = math.floor(exponent), math.ceil(exponent)
        target = latent.shape[2] * self.latent_factor
        length = 2 ** int(min((lo, hi), key=lambda z: abs(target - 2 ** z)))         # utils.py:45-49
        noise = torch.randn((b, self.in_channels, length), device=latent.device,
                            dtype=latent.dtype, generator=generator)
        channels = [None] * self.inject_depth + [latent]
        out = super().sample(noise, channels=channels, **kwargs)
        return self.adapter.decode(out) if self.adapter is not None else out


class DiffusionARPort(DiffusionModelPort):
    """models.py:227-250 (DiffusionAR): net input = waveform + sigma channel, no time
    conditioning, no modulation (SkipCat merges)."""

    def __init__(self, in_channels: int, length: int, num_splits: int, **kwargs):
        super().__init__(in_channels=in_channels + 1, out_channels=in_channels,
                         diffusion_t=ARVDiffusionPort, diffusion_length=length,
                         diffusion_num_splits=num_splits, sampler_t=ARVSamplerPort,
                         sampler_in_channels=in_channels, sampler_length=length,
                         sampler_num_splits=num_splits, use_time_conditioning=False,
                         use_modulation=False, **kwargs)


class ToyEncoder(nn.Module):
    """Test fixture: a minimal encoder with the interface DiffusionAE needs (`out_channels`,
    `downsample_factor`, `forward(x, with_info)`); stands in for audio_encoders_pytorch.MelE1d."""

    def __init__(self, in_channels: int = 2, out_channels: int = 16, downsample_factor: int = 16):
        super().__init__()
        self.out_channels, self.downsample_factor = out_channels, downsample_factor
        self.conv = nn.Conv1d(in_channels, out_channels, downsample_factor, stride=downsample_factor)

    def forward(self, x: Tensor, with_info: bool = False):
        z = torch.tanh(self.conv(x))
        return (z, {"latent": z}) if with_info else z


class DiffusionUpsamplerPort(DiffusionModelPort):
    """models.py:134-165 (DiffusionUpsampler)."""

    def __init__(self, in_channels: int, upsample_factor: int,
                 net_t: Callable = build_unet_v0, **kwargs):
        self.upsample_factor = upsample_factor
        super().__init__(net_t=append_channels_plugin(net_t, channels=in_channels),
                         in_channels=in_channels, **kwargs)

    def reupsample(self, x: Tensor) -> Tensor:                                       # :149-153
        return sinc_upsample(sinc_downsample(x.clone(), self.upsample_factor),
                             self.upsample_factor)

    def forward(self, x: Tensor, *args, **kwargs) -> Tensor:                         # :155-157
        return super().forward(x, *args, append_channels=self.reupsample(x), **kwargs)

    @torch.no_grad()
    def sample(self, downsampled: Tensor, generator: Optional[Generator] = None,
               **kwargs) -> Tensor:                                                   # :159-165
        re = sinc_upsample(downsampled, self.upsample_factor)
        noise = cpu_randn_like(re, generator=generator)
        return super().sample(noise, append_channels=re, **kwargs)


class MelSpectrogramPort(nn.Module):
    """components.py:188-236 (MelSpectrogram); vocoder *training* only."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int, sample_rate: int,
                 n_mel_channels: int, center: bool = False, normalize: bool = False,
                 normalize_log: bool = False):
        super().__init__()
        from torchaudio import transforms
        self.padding = (n_fft - hop_length) // 2
        self.normalize, self.normalize_log = normalize, normalize_log
        self.to_spectrogram = transforms.Spectrogram(n_fft=n_fft, hop_length=hop_length,
                                                     win_length=win_length, center=center,
                                                     power=None)
        self.to_mel_scale = transforms.MelScale(n_mels=n_mel_channels, n_stft=n_fft // 2 + 1,
                                                sample_rate=sample_rate)

    def forward(self, wave: Tensor) -> Tensor:
        lead = wave.shape[:-1]
        flat = F.pad(wave.reshape(-1, wave.shape[-1]), [self.padding] * 2, mode="reflect")
        mel = self.to_mel_scale(torch.abs(self.to_spectrogram(flat)))
        if self.normalize:
            mel = mel / torch.max(mel)
            mel = 2 * torch.pow(mel, 0.25) - 1
        if self.normalize_log:
            mel = torch.log(torch.clamp(mel, min=1e-5))
        return mel.reshape(*lead, *mel.shape[-2:])


class DiffusionVocoderPort(DiffusionModelPort):
    """models.py:168-224 (DiffusionVocoder)."""

    def __init__(self, mel_channels: int, mel_n_fft: int, mel_hop_length: Optional[int] = None,
                 mel_win_length: Optional[int] = None, net_t: Callable = build_unet_v0,
                 in_channels: int = 1, **kwargs):
        hop = a_unet.default(mel_hop_length, math.floor(mel_n_fft) // 4)
        win = a_unet.default(mel_win_length, mel_n_fft)
        mel_kw, kwargs = split_prefixed("mel_", kwargs)
        super().__init__(net_t=append_channels_plugin(net_t, channels=1), in_channels=1,
                         **kwargs)
        self.to_spectrogram = MelSpectrogramPort(n_fft=mel_n_fft, hop_length=hop, win_length=win,
                                                 n_mel_channels=mel_channels, **mel_kw)
        self.to_flat = nn.ConvTranspose1d(mel_channels, 1, kernel_size=win, stride=hop,
                                          padding=(win - hop) // 2, bias=False)       # :194-201

    def forward(self, x: Tensor, *args, **kwargs) -> Tensor:                          # :203-209
        spec = self.to_spectrogram(x)
        flat = self.to_flat(spec.reshape(-1, *spec.shape[-2:]))
        x = x.reshape(-1, 1, x.shape[-1])
        return super().forward(x, *args, append_channels=flat, **kwargs)

    @torch.no_grad()
    def sample(self, spectrogram: Tensor, generator: Optional[Generator] = None,
               **kwargs) -> Tensor:                                                    # :211-224
        lead = spectrogram.shape[:-2]
        flat = self.to_flat(spectrogram.reshape(-1, *spectrogram.shape[-2:]))
        noise = cpu_randn_like(flat, generator=generator)
        wave = super().sample(noise, append_channels=flat, **kwargs)
        return wave.reshape(*lead, wave.shape[-1])


README_UNCONDITIONAL = dict(  # reference README.md:22-34
    in_channels=2,
    channels=[8, 32, 64, 128, 256, 512, 512, 1024, 1024],
    factors=[1, 4, 4, 4, 2, 2, 2, 2, 2],
    items=[1, 2, 2, 2, 2, 2, 2, 4, 4],
    attentions=[0, 0, 0, 0, 0, 1, 1, 1, 1],
    attention_heads=8,
    attention_features=64,
)
