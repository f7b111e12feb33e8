"""Host-side helpers around the hot path: keyword routing for the plugin slots, the polyphase
windowed-sinc resampler behind DiffusionUpsampler (one strided convolution per call, outside
the step loop; its filter bank is built once per (factors, dtype, device) and cached), and the
CPU-generator noise draw.  Behavioural spec: reference utils.py:17-125."""
import functools
import inspect
import math
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Generator, Tensor


def exists(val: Any) -> bool:
    return val is not None


def default(val: Any, fallback: Union[Any, Callable[[], Any]]) -> Any:
    """`val` unless it is None; a plain-function fallback (def / lambda) is evaluated lazily,
    any other object -- including callables such as nn.Module instances -- is returned as is
    (reference utils.py:27-30, `inspect.isfunction`)."""
    if val is not None:
        return val
    return fallback() if inspect.isfunction(fallback) else fallback


def closest_power_2(x: float) -> int:
    """The power of two nearest to x (reference utils.py:45-49; DiffusionAE.decode sizes its noise with it)."""
    lo = 2 ** int(math.floor(math.log2(x)))
    return lo if x - lo <= 2 * lo - x else 2 * lo


def groupby(prefix: str, d: Dict[str, Any], keep_prefix: bool = False) -> Tuple[Dict, Dict]:
    """Partition keyword arguments: ({keys that start with `prefix`, prefix removed unless
    keep_prefix}, {all others}) — how `diffusion_*`, `sampler_*` and `mel_*` options reach their
    plugin (reference utils.py:48-70)."""
    taken: Dict[str, Any] = {}
    others: Dict[str, Any] = {}
    cut = 0 if keep_prefix else len(prefix)
    for key, value in d.items():
        if key.startswith(prefix):
            taken[key[cut:]] = value
        else:
            others[key] = value
    return taken, others


@functools.lru_cache(maxsize=32)
def _polyphase_bank(factor_in: int, factor_out: int, rolloff: float, zeros: int, dtype: torch.dtype,
                    device: str) -> Tuple[Tensor, int]:
    """FIR bank [factor_out, 1, taps] of the Hann-windowed sinc low-pass at `rolloff` x Nyquist
    of the slower rate, one phase per output sample inside an input period, and the half width
    (in input samples) the signal has to be padded by.  The expression order follows the
    oracle's (bit-identical filters)."""
    opts = dict(device=torch.device(device), dtype=dtype)
    cutoff = min(factor_in, factor_out) * rolloff
    half = math.ceil(zeros * factor_in / cutoff)
    grid = torch.arange(-half, half + factor_in, **opts)[None, None] / factor_in
    theta = torch.arange(0, -factor_out, step=-1, **opts)[:, None, None] / factor_out + grid
    theta = (theta * cutoff).clamp(-zeros, zeros) * math.pi
    hann = torch.cos(theta / zeros / 2) ** 2
    sinc = torch.where(theta == 0, torch.tensor(1.0).to(theta), theta.sin() / theta)
    return sinc * (hann * (cutoff / factor_in)), half


class _FirResample(torch.autograd.Function):
    """The polyphase FIR as one kernel (adp_resample); backward = its transpose (adp_resample_adjoint)."""

    @staticmethod
    def forward(ctx, rows: Tensor, bank: Tensor, factor_in: int, factor_out: int, half: int, t_out: int):
        from . import ops
        ctx.bank, ctx.geom = bank, (factor_in, factor_out, half, t_out, rows.shape[1])
        return ops.fir_resample(rows.contiguous(), bank, factor_in, factor_out, half, t_out)

    @staticmethod
    def backward(ctx, dy: Tensor):
        from . import ops
        fi, fo, half, t_out, t = ctx.geom
        return ops.fir_resample(dy.contiguous(), ctx.bank, fi, fo, half, t_out, adjoint_of=t), None, None, \
            None, None, None


def resample(waveforms: Tensor, factor_in: int, factor_out: int, rolloff: float = 0.99,
             lowpass_filter_width: int = 6) -> Tensor:
    """[b, c, t] -> [b, c, t * factor_out / factor_in].  CUDA tensors go through the adp_resample
    kernel (one pass, no padded copy, no [rows, phases, frames] intermediate); host tensors through
    the same filter bank as a strided convolution."""
    b, c, t = waveforms.shape
    if waveforms.is_cuda:
        bank, half = _polyphase_bank(int(factor_in), int(factor_out), float(rolloff),
                                     int(lowpass_filter_width), torch.float32, str(waveforms.device))
        out = _FirResample.apply(waveforms.reshape(b * c, t).float(), bank[:, 0].contiguous(),
                                 int(factor_in), int(factor_out), half, int(factor_out * t / factor_in))
        return out.reshape(b, c, -1).to(waveforms.dtype)
    bank, half = _polyphase_bank(int(factor_in), int(factor_out), float(rolloff),
                                 int(lowpass_filter_width), waveforms.dtype, str(waveforms.device))
    rows = F.pad(waveforms.reshape(b * c, t), (half, half + factor_in))
    phases = F.conv1d(rows[:, None], bank, stride=factor_in)             # [(b c), factor_out, frames]
    woven = phases.reshape(b, c, factor_out, -1).permute(0, 1, 3, 2).reshape(b, c, -1)
    return woven[..., : int(factor_out * t / factor_in)]


def downsample(waveforms: Tensor, factor: int, **kwargs) -> Tensor:
    return resample(waveforms, factor_in=factor, factor_out=1, **kwargs)


def upsample(waveforms: Tensor, factor: int, **kwargs) -> Tensor:
    return resample(waveforms, factor_in=1, factor_out=factor, **kwargs)


def randn_like(tensor: Tensor, *args, generator: Optional[Generator] = None, **kwargs) -> Tensor:
    """Gaussian noise shaped like `tensor`, drawn on the CPU generator (so a seed reproduces
    the reference's samples on any device) and then moved to `tensor`'s device and dtype."""
    return torch.randn(tensor.shape, *args, generator=generator, **kwargs).to(tensor)
