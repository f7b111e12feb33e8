"""audio_diffusion_pytorch on B200: the reference's public names (reference __init__.py:1-20)
for the UNetV0 + VDiffusion/VSampler hot path, executed by hand-written sm_100a kernels.

Out-of-scope names of the reference raise on use instead of silently
running something else."""
from .components import AppendChannelsPlugin, MelSpectrogram, UNetV0
from .diffusion import (ARVDiffusion, ARVSampler, Diffusion, Distribution, Inpainter, LinearSchedule,
                        Sampler, Schedule, UniformDistribution, VDiffusion, VInpainter, VSampler)
from .models import (AdapterBase, DiffusionAE, DiffusionAR, DiffusionModel, DiffusionUpsampler,
                     DiffusionVocoder, EncoderBase)
from .unet import B200UNet


def _out_of_scope(name: str, why: str):
    class _Missing:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the B200 hot path this package "
                                      f"replaces ({why}); use the reference implementation")
    _Missing.__name__ = name
    return _Missing


XUNet = B200UNet
LTPlugin = _out_of_scope("LTPlugin", "not used by any model class or config")

__all__ = ["UNetV0", "XUNet", "LTPlugin", "MelSpectrogram", "VDiffusion", "VSampler", "VInpainter",
           "LinearSchedule", "UniformDistribution", "Diffusion", "Distribution", "Sampler",
           "Schedule", "DiffusionModel", "DiffusionUpsampler", "DiffusionVocoder", "DiffusionAE",
           "DiffusionAR", "ARVDiffusion", "ARVSampler", "EncoderBase", "AdapterBase", "AppendChannelsPlugin", "B200UNet", "Inpainter"]
