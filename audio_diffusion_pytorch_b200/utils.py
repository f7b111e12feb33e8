"""Host-side helpers of the hot path (reference utils.py): kwarg routing, the windowed-sinc
resampler used once per DiffusionUpsampler call, CPU-generator randn."""
from math import ceil, pi
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Generator, Tensor


def exists(val) -> bool:
    return val is not None


def default(val, d):
    if exists(val):
        return val
    return d() if callable(d) and not isinstance(d, type) else d


def groupby(prefix: str, d: Dict, keep_prefix: bool = False) -> Tuple[Dict, Dict]:
    """Splits kwargs on a prefix (reference utils.py:48-70): (`prefix*` stripped, the rest)."""
    with_prefix = {k: v for k, v in d.items() if k.startswith(prefix)}
    rest = {k: v for k, v in d.items() if not k.startswith(prefix)}
    if not keep_prefix:
        with_prefix = {k[len(prefix):]: v for k, v in with_prefix.items()}
    return with_prefix, rest


def resample(waveforms: Tensor, factor_in: int, factor_out: int, rolloff: float = 0.99,
             lowpass_filter_width: int = 6) -> Tensor:
    """Windowed-sinc polyphase resampling (reference utils.py:82-109): one strided conv with
    `factor_out` FIR phases, interleaved.  Runs once per call, outside the step loop."""
    b, c, length = waveforms.shape
    n_out = int(factor_out * length / factor_in)
    kw = dict(device=waveforms.device, dtype=waveforms.dtype)
    base = min(factor_in, factor_out) * rolloff
    width = ceil(lowpass_filter_width * factor_in / base)
    taps = torch.arange(-width, width + factor_in, **kw)[None, None] / factor_in
    phase = torch.arange(0, -factor_out, step=-1, **kw)[:, None, None] / factor_out
    t = ((phase + taps) * base).clamp(-lowpass_filter_width, lowpass_filter_width) * pi
    window = torch.cos(t / lowpass_filter_width / 2) ** 2
    fir = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * (window * (base / factor_in))
    x = F.pad(waveforms.reshape(b * c, 1, length), (width, width + factor_in))
    y = F.conv1d(x, fir, stride=factor_in)                      # [(b c), factor_out, frames]
    y = y.transpose(1, 2).reshape(b, c, -1)                      # interleave the phases
    return y[..., :n_out]


def downsample(waveforms: Tensor, factor: int, **kwargs) -> Tensor:
    return resample(waveforms, factor_in=factor, factor_out=1, **kwargs)


def upsample(waveforms: Tensor, factor: int, **kwargs) -> Tensor:
    return resample(waveforms, factor_in=1, factor_out=factor, **kwargs)


def randn_like(tensor: Tensor, *args, generator: Optional[Generator] = None, **kwargs) -> Tensor:
    """reference utils.py:123-125: drawn on the CPU generator, then moved to `tensor`."""
    return torch.randn(tensor.shape, *args, generator=generator, **kwargs).to(tensor)
