"""ORACLE / TEST INFRASTRUCTURE -- not product code.

Pure-PyTorch stand-in for the third-party package ``a_unet`` (PyPI ``a-unet``,
un-pinned in the reference's setup.py:20, NOT vendored under /root/reference and
not installable offline).  The reference imports seven names from ``a_unet`` and
nine from ``a_unet.apex`` (reference audio_diffusion_pytorch/components.py:5-24).
Their published behaviour is restated here from SURVEY.md appendix A (recall of
upstream archinetai/a-unet ~v0.0.16).  PARITY UNPINNED: upstream a_unet source is
unavailable, so this shim *is* the definition of the oracle for the U-Net math.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg may import this package.
"""
from math import pi
from typing import Callable, Optional, Sequence

import torch
import torch.nn.functional as F
from einops import rearrange
from torch import Tensor, nn


def exists(val) -> bool:
    return val is not None


def default(val, d):
    return val if exists(val) else d


class T:
    """Type template: ``T(t)(*a, **ka)`` is a deferred constructor.  Keyword
    arguments given to the template are readable as attributes (the reference
    zips over ``XBlock(...)`` objects and a_unet reads ``block.channels``)."""

    def __init__(self, t: Callable, override: bool = True):
        self.t, self.override = t, override

    def __call__(self, *a, **ka):
        t, override = self.t, self.override

        class Template:
            def __init__(self):
                self.args = a
                self.__dict__.update(**ka)

            def __call__(self, *b, **kb):
                if override:
                    return t(*(*a, *b), **{**ka, **kb})
                return t(*(*b, *a), **{**kb, **ka})

        return Template()


def Module(modules: Sequence[nn.Module], forward_fn: Callable) -> nn.Module:
    """Closure-style module (used by the reference at components.py:157,178)."""

    class _Module(nn.Module):
        def __init__(self):
            super().__init__()
            self.blocks = nn.ModuleList(modules)

        def forward(self, *args, **kwargs):
            return forward_fn(*args, **kwargs)

    return _Module()


class Sequential(nn.Module):
    """nn.Sequential that forwards the extra positional args to every block."""

    def __init__(self, *blocks):
        super().__init__()
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x: Tensor, *args) -> Tensor:
        for block in self.blocks:
            x = block(x, *args)
        return x


def Select(args_fn: Callable) -> Callable:
    def wrap(block_t: Callable) -> Callable:
        def build(*a, **k) -> nn.Module:
            block = block_t(*a, **k)
            return Module([block], lambda *args, **kw: block(*args_fn(*args), **kw))

        return build

    return wrap


class Packed(Sequential):
    """Runs the blocks on [b, n, d] instead of [b, d, n]."""

    def forward(self, x: Tensor, *args) -> Tensor:
        x = rearrange(x, "b d n -> b n d")
        x = super().forward(x, *args)
        return rearrange(x, "b n d -> b d n")


def Repeat(m, times: int) -> nn.Module:
    """A module *instance* repeated => shared weights."""
    return nn.Sequential(*([m] * times))


def Conv(dim: int, *args, **kwargs) -> nn.Module:
    return [nn.Conv1d, nn.Conv2d, nn.Conv3d][dim - 1](*args, **kwargs)


def ConvTranspose(dim: int, *args, **kwargs) -> nn.Module:
    return [nn.ConvTranspose1d, nn.ConvTranspose2d, nn.ConvTranspose3d][dim - 1](*args, **kwargs)


def Downsample(dim: int, factor: int = 2, width: int = 1, conv_t=Conv, **kwargs) -> nn.Module:
    width = width if factor > 1 else 1
    return conv_t(
        dim=dim,
        kernel_size=factor * width,
        stride=factor,
        padding=(factor * width - factor) // 2,
        **kwargs,
    )


def Upsample(dim: int, factor: int = 2, mode: str = "nearest", in_channels: int = 0,
             out_channels: int = 0) -> nn.Module:
    assert mode == "nearest", "only the UNetV0 default (nearest + conv3) is restated"
    return nn.Sequential(
        nn.Upsample(scale_factor=factor, mode="nearest"),
        Conv(dim=dim, in_channels=in_channels, out_channels=out_channels, kernel_size=3, padding=1),
    )


def ConvBlock(dim: int, in_channels: int, out_channels: int, num_groups: int,
              kernel_size: int = 3) -> nn.Module:
    return nn.Sequential(
        nn.GroupNorm(num_groups=num_groups, num_channels=in_channels),
        nn.SiLU(),
        Conv(dim=dim, in_channels=in_channels, out_channels=out_channels,
             kernel_size=kernel_size, padding=(kernel_size - 1) // 2),
    )


def ResnetBlock(dim: int, in_channels: int, out_channels: int, num_groups: int,
                kernel_size: int = 3) -> nn.Module:
    block1 = ConvBlock(dim, in_channels, out_channels, num_groups, kernel_size)
    block2 = ConvBlock(dim, out_channels, out_channels, num_groups, kernel_size)
    to_out = (nn.Identity() if in_channels == out_channels
              else Conv(dim=dim, in_channels=in_channels, out_channels=out_channels, kernel_size=1))
    return Module([block1, block2, to_out], lambda x: block2(block1(x)) + to_out(x))


def Modulation(in_features: int, num_features: int) -> nn.Module:
    to_scale_shift = nn.Sequential(nn.SiLU(), nn.Linear(num_features, in_features * 2, bias=True))
    norm = nn.LayerNorm(in_features, elementwise_affine=False, eps=1e-6)

    def forward(x: Tensor, features: Tensor) -> Tensor:
        scale, shift = rearrange(to_scale_shift(features), "b d -> b 1 d").chunk(2, dim=-1)
        return norm(x) * (1 + scale) + shift

    return Module([to_scale_shift, norm], forward)


def MergeAdd() -> nn.Module:
    return Module([], lambda x, y, *_: x + y)


def MergeCat(dim: int, channels: int, scale: float = 2 ** -0.5) -> nn.Module:
    conv = Conv(dim=dim, in_channels=channels * 2, out_channels=channels, kernel_size=1)
    return Module([conv], lambda x, y, *_: conv(torch.cat([x * scale, y], dim=1)))


def MergeModulate(dim: int, channels: int, modulation_features: int) -> nn.Module:
    to_scale = nn.Sequential(nn.SiLU(), nn.Linear(modulation_features, channels, bias=True))

    def forward(x: Tensor, y: Tensor, features: Tensor, *_) -> Tensor:
        scale = to_scale(features).view(*features.shape[:1], channels, *((1,) * dim))
        return x + scale * y

    return Module([to_scale], forward)


def AttentionBase(features: int, head_features: int, num_heads: int) -> nn.Module:
    scale = head_features ** -0.5
    to_out = nn.Linear(head_features * num_heads, features, bias=False)

    def forward(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
        q, k, v = (rearrange(t, "b n (h d) -> b h n d", h=num_heads) for t in (q, k, v))
        sim = torch.einsum("b h n d, b h m d -> b h n m", q, k) * scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum("b h n m, b h m d -> b h n d", attn, v)
        return to_out(rearrange(out, "b h n d -> b n (h d)"))

    return Module([to_out], forward)


def Attention(features: int, head_features: int, num_heads: int,
              context_features: Optional[int] = None) -> nn.Module:
    context_features = default(context_features, features)
    mid = head_features * num_heads
    norm = nn.LayerNorm(features)
    norm_context = nn.LayerNorm(context_features)
    to_q = nn.Linear(features, mid, bias=False)
    to_kv = nn.Linear(context_features, mid * 2, bias=False)
    attention = AttentionBase(features, head_features, num_heads)

    def forward(x: Tensor, context: Optional[Tensor] = None) -> Tensor:
        skip = x
        context = default(context, x)
        x, context = norm(x), norm_context(context)
        k, v = to_kv(context).chunk(2, dim=-1)
        return skip + attention(to_q(x), k, v)

    return Module([norm, norm_context, to_q, to_kv, attention], forward)


class NumberEmbedder(nn.Module):
    def __init__(self, features: int, dim: int = 256):
        super().__init__()
        assert dim % 2 == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))
        self.to_out = nn.Linear(dim + 1, features)

    def forward(self, x: Tensor) -> Tensor:
        x = x.unsqueeze(-1)
        freqs = x * self.weights * 2 * pi
        fouriered = torch.cat([x, freqs.sin(), freqs.cos()], dim=-1)
        return self.to_out(fouriered)


class FixedEmbedding(nn.Module):
    def __init__(self, max_length: int, features: int):
        super().__init__()
        self.max_length = max_length
        self.embedding = nn.Embedding(max_length, features)

    def forward(self, x: Tensor) -> Tensor:
        b, n = x.shape[0], x.shape[1]
        assert n <= self.max_length, "Input sequence length must be <= max_length"
        position = torch.arange(n, device=x.device)
        return self.embedding(position).unsqueeze(0).expand(b, -1, -1)


"""Plugins (reference components.py:64-76 composes these inside-out)."""


def TimeConditioningPlugin(net_t: Callable, num_layers: int = 2) -> Callable:
    def Net(modulation_features: Optional[int] = None, **kwargs) -> nn.Module:
        assert exists(modulation_features), "TimeConditioningPlugin requires modulation_features"
        embedder = NumberEmbedder(features=modulation_features)
        mlp = Repeat(nn.Sequential(nn.Linear(modulation_features, modulation_features), nn.GELU()),
                     times=num_layers)
        net = net_t(modulation_features=modulation_features, **kwargs)

        def forward(x: Tensor, time: Optional[Tensor] = None,
                    features: Optional[Tensor] = None, **kw):
            assert exists(time), "time conditioning requires the time argument"
            t = mlp(F.gelu(embedder(time)))
            if t.ndim == 3:
                t = t.sum(dim=1)
            features = features + t if exists(features) else t
            return net(x, features=features, **kw)

        return Module([embedder, mlp, net], forward)

    return Net


def ClassifierFreeGuidancePlugin(net_t: Callable, embedding_max_length: int) -> Callable:
    def Net(embedding_features: int, **kwargs) -> nn.Module:
        fixed_embedding = FixedEmbedding(max_length=embedding_max_length, features=embedding_features)
        net = net_t(embedding_features=embedding_features, **kwargs)

        def forward(x: Tensor, embedding: Optional[Tensor] = None, embedding_scale: float = 1.0,
                    embedding_mask_proba: float = 0.0, **kw):
            assert exists(embedding), "ClassiferFreeGuidancePlugin requires embedding"
            b, device = embedding.shape[0], embedding.device
            embedding_mask = fixed_embedding(embedding)
            if embedding_mask_proba > 0.0:
                batch_mask = torch.bernoulli(
                    torch.full((b, 1, 1), embedding_mask_proba, device=device)).to(torch.bool)
                embedding = torch.where(batch_mask, embedding_mask, embedding)
            if embedding_scale != 1.0:
                out = net(x, embedding=embedding, **kw)
                out_masked = net(x, embedding=embedding_mask, **kw)
                return out_masked + (out - out_masked) * embedding_scale
            return net(x, embedding=embedding, **kw)

        return Module([fixed_embedding, net], forward)

    return Net


def TextConditioningPlugin(net_t: Callable, embedder: Optional[nn.Module] = None) -> Callable:
    def Net(**kwargs):
        raise RuntimeError("TextConditioningPlugin needs t5-base weights (no network); pass a "
                           "precomputed `embedding=` with use_text_conditioning=False instead")

    return Net
