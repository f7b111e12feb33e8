"""ORACLE / TEST INFRASTRUCTURE -- stand-in for ``a_unet.apex`` (see package docstring).

Names exported for the reference (components.py:14-24): AttentionItem,
CrossAttentionItem, InjectChannelsItem, ModulationItem, ResnetItem, SkipCat,
SkipModulate, XBlock, XUNet.  Behaviour per SURVEY.md appendix A.2/A.3.
"""
from typing import Callable, List, Optional, Sequence

import torch
from torch import Tensor, nn

from . import (Attention, Conv, Downsample, MergeAdd, MergeCat, MergeModulate, Modulation, Module,
               Packed, ResnetBlock, Select, Sequential, T, Upsample, default, exists)

# positional calling convention of every item: (x, features, embedding, channels)
SelectX = Select(lambda x, *_: (x,))
SelectXF = Select(lambda x, f, *_: (x, f))
SelectXE = Select(lambda x, f, e, *_: (x, e))
SelectXC = Select(lambda x, f, e, c, *_: (x, c))


def DownsampleItem(dim=None, factor=None, in_channels=None, channels=None,
                   downsample_width: int = 1, **kwargs) -> nn.Module:
    msg = "DownsampleItem requires dim, factor, in_channels, channels"
    assert all(exists(v) for v in (dim, factor, in_channels, channels)), msg
    return SelectX(Downsample)(dim=dim, factor=factor, width=downsample_width,
                               in_channels=in_channels, out_channels=channels)


def UpsampleItem(dim=None, factor=None, channels=None, out_channels=None,
                 upsample_mode: str = "nearest", **kwargs) -> nn.Module:
    msg = "UpsampleItem requires dim, factor, channels, out_channels"
    assert all(exists(v) for v in (dim, factor, channels, out_channels)), msg
    return SelectX(Upsample)(dim=dim, factor=factor, mode=upsample_mode,
                             in_channels=channels, out_channels=out_channels)


def ResnetItem(dim=None, channels=None, resnet_groups=None, resnet_kernel_size: int = 3,
               **kwargs) -> nn.Module:
    msg = "ResnetItem requires dim, channels, and resnet_groups"
    assert all(exists(v) for v in (dim, channels, resnet_groups)), msg
    return SelectX(ResnetBlock)(dim=dim, in_channels=channels, out_channels=channels,
                                num_groups=resnet_groups, kernel_size=resnet_kernel_size)


def ModulationItem(channels=None, modulation_features=None, **kwargs) -> nn.Module:
    msg = "ModulationItem requires channels, modulation_features"
    assert all(exists(v) for v in (channels, modulation_features)), msg
    mod = Modulation(in_features=channels, num_features=modulation_features)
    return SelectXF(lambda: Packed(mod))()


def AttentionItem(channels=None, attention_features=None, attention_heads=None,
                  **kwargs) -> nn.Module:
    msg = "AttentionItem requires channels, attention_features, attention_heads"
    assert all(exists(v) for v in (channels, attention_features, attention_heads)), msg
    att = Attention(features=channels, head_features=attention_features, num_heads=attention_heads)
    return SelectX(lambda: Packed(att))()


def CrossAttentionItem(channels=None, attention_features=None, attention_heads=None,
                       embedding_features=None, **kwargs) -> nn.Module:
    msg = "CrossAttentionItem requires channels, embedding_features, attention_*"
    assert all(exists(v) for v in (channels, attention_features, attention_heads,
                                   embedding_features)), msg
    att = Attention(features=channels, head_features=attention_features,
                    num_heads=attention_heads, context_features=embedding_features)
    return SelectXE(lambda: Packed(att))()


def InjectChannelsItem(dim=None, channels=None, depth=None, context_channels=None,
                       **kwargs) -> nn.Module:
    msg = "InjectChannelsItem requires dim, depth, channels, context_channels"
    assert all(exists(v) for v in (dim, depth, channels, context_channels)), msg
    conv = Conv(dim=dim, in_channels=channels + context_channels, out_channels=channels,
                kernel_size=1)

    def forward(x: Tensor, channels_list: Sequence[Tensor]) -> Tensor:
        msg_ = "context `channels` at depth must match resolution and context_channels"
        ctx = channels_list[depth]
        assert ctx.shape[1] == context_channels and ctx.shape[2:] == x.shape[2:], msg_
        return conv(torch.cat([x, ctx], dim=1)) + x

    return SelectXC(lambda: Module([conv], forward))()


def SkipAdapter(dim=None, in_channels=None, out_channels=None, **kwargs) -> nn.Module:
    assert all(exists(v) for v in (dim, in_channels, out_channels))
    if in_channels == out_channels:
        return SelectX(nn.Identity)()
    return SelectX(Conv)(dim=dim, in_channels=in_channels, out_channels=out_channels,
                         kernel_size=1)


def SkipAdd(**kwargs) -> nn.Module:
    return MergeAdd()


def SkipCat(dim=None, out_channels=None, skip_scale: float = 2 ** -0.5, **kwargs) -> nn.Module:
    assert all(exists(v) for v in (dim, out_channels))
    return MergeCat(dim=dim, channels=out_channels, scale=skip_scale)


def SkipModulate(dim=None, out_channels=None, modulation_features=None, **kwargs) -> nn.Module:
    assert all(exists(v) for v in (dim, out_channels, modulation_features))
    return MergeModulate(dim=dim, channels=out_channels, modulation_features=modulation_features)


class Block(nn.Module):
    """One U-Net level: skip=adapter(x); x=down,items,[inner],items_up,up; merge(skip, x, f)."""

    def __init__(self, in_channels: int, downsample_t: Callable = DownsampleItem,
                 upsample_t: Callable = UpsampleItem, skip_t: Callable = SkipAdd,
                 skip_adapter_t: Callable = SkipAdapter, items: Sequence[Callable] = (),
                 items_up: Optional[Sequence[Callable]] = None,
                 out_channels: Optional[int] = None, inner_block: Optional[nn.Module] = None,
                 **kwargs):
        super().__init__()
        out_channels = default(out_channels, in_channels)
        items_up = default(items_up, items)
        items_kwargs = dict(in_channels=in_channels, out_channels=out_channels, **kwargs)
        items_down_built = [item_t(**items_kwargs) for item_t in items]
        items_up_built = [item_t(**items_kwargs) for item_t in items_up]
        self.skip_adapter = skip_adapter_t(**items_kwargs)
        self.block = Sequential(
            downsample_t(**items_kwargs),
            *items_down_built,
            *([inner_block] if exists(inner_block) else []),
            *items_up_built,
            upsample_t(**items_kwargs),
        )
        self.skip_merge = skip_t(**items_kwargs)

    def forward(self, x: Tensor, features: Optional[Tensor] = None,
                embedding: Optional[Tensor] = None,
                channels: Optional[Sequence[Tensor]] = None) -> Tensor:
        skip = self.skip_adapter(x)
        x = self.block(x, features, embedding, channels)
        return self.skip_merge(skip, x, features)


XBlock = T(Block, override=False)


class XUNet(nn.Module):
    def __init__(self, in_channels: int, blocks: Sequence, out_channels: Optional[int] = None,
                 **kwargs):
        super().__init__()
        num_layers = len(blocks)
        out_channels = default(out_channels, in_channels)

        def Net(i: int) -> Optional[nn.Module]:
            if i == num_layers:
                return None
            block_t = blocks[i]
            in_ch = in_channels if i == 0 else blocks[i - 1].channels
            out_ch = out_channels if i == 0 else in_ch
            return block_t(in_channels=in_ch, out_channels=out_ch, depth=i,
                           inner_block=Net(i + 1), **kwargs)

        self.net = Net(0)

    def forward(self, x: Tensor, *, features: Optional[Tensor] = None,
                embedding: Optional[Tensor] = None,
                channels: Optional[Sequence[Tensor]] = None) -> Tensor:
        return self.net(x, features, embedding, channels)
