"""Conditioning front-ends as kernels: adp_resample (+ adjoint), adp_mel_spectrogram,
adp_to_flat (+ backward) against the same modules' tensor-op route on the CPU in fp32 (the route
the CPU oracle tests pin to the reference: tests/test_host_cpu.py).  fp32 throughout: rel-L2 1e-5
for the FIR / transposed-conv kernels, 1e-4 for the FFT-based spectrogram."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("factor_in,factor_out,t", [(1, 16, 4096), (16, 1, 65536), (4, 1, 1000), (3, 2, 3000),
                                                     (1, 2, 37)])
def test_resample_kernel_and_its_adjoint(factor_in, factor_out, t):
    from audio_diffusion_pytorch_b200.utils import resample
    torch.manual_seed(0)
    x = torch.randn(3, 2, t)
    x_ref = x.clone().requires_grad_()
    want = resample(x_ref, factor_in, factor_out)                  # host route: strided convolution
    x_gpu = x.to(DEV).requires_grad_()
    got = resample(x_gpu, factor_in, factor_out)
    assert got.shape == want.shape
    assert rel_l2(got, want) <= 1e-5
    d = torch.randn_like(want)
    want.backward(d)
    got.backward(d.to(DEV))
    assert rel_l2(x_gpu.grad, x_ref.grad) <= 1e-5


@pytest.mark.parametrize("cfg", [
    dict(n_fft=1024, hop_length=256, win_length=1024, sample_rate=48000, n_mel_channels=80),
    dict(n_fft=256, hop_length=64, win_length=128, sample_rate=16000, n_mel_channels=16),
    dict(n_fft=2048, hop_length=300, win_length=2048, sample_rate=44100, n_mel_channels=128, normalize_log=True),
    dict(n_fft=64, hop_length=16, win_length=64, sample_rate=48000, n_mel_channels=8, normalize=True),
])
def test_mel_spectrogram_kernel(cfg):
    from audio_diffusion_pytorch_b200.components import MelSpectrogram
    torch.manual_seed(1)
    front = MelSpectrogram(**cfg)
    wave = torch.randn(2, 2, 2 ** 14) * torch.linspace(0.05, 1.0, 2 ** 14)
    want = front(wave)                                              # torchaudio STFT + MelScale on the CPU
    got = front.to(DEV)(wave.to(DEV))
    assert got.shape == want.shape
    if cfg.get("normalize_log") and not cfg.get("normalize"):
        # log(max(mel, 1e-5)): compare where the clamp is inactive, absolutely elsewhere
        assert float((got.cpu() - want).abs().max()) <= 2e-3
        assert rel_l2(got.exp(), want.exp()) <= 1e-4
    else:
        assert rel_l2(got, want) <= 1e-4
    odd = wave[..., : 2 ** 14 - 123]                                # frame count not a multiple of 8
    assert rel_l2(front(odd.to(DEV)), front.cpu()(odd)) <= (2e-3 if cfg.get("normalize_log") else 1e-4)


@pytest.mark.parametrize("mel,win,hop,frames", [(80, 1024, 256, 64), (16, 256, 64, 33), (8, 64, 16, 5)])
def test_to_flat_kernel_forward_and_gradients(mel, win, hop, frames):
    import audio_diffusion_pytorch_b200 as adp
    torch.manual_seed(2)
    voc = adp.DiffusionVocoder(net_t=adp.UNetV0, channels=[8, 32], factors=[1, 4], items=[1, 1],
                               mel_n_fft=win, mel_channels=mel, mel_sample_rate=48000, mel_hop_length=hop)
    spec = torch.randn(3, mel, frames)
    s_ref = spec.clone().requires_grad_()
    want, _ = voc._unroll(s_ref)                                    # nn.ConvTranspose1d on the CPU
    d = torch.randn_like(want)
    want.backward(d)
    w_grad_ref = voc.to_flat.weight.grad.clone()
    voc.zero_grad()
    voc = voc.to(DEV)
    s_gpu = spec.to(DEV).requires_grad_()
    got, _ = voc._unroll(s_gpu)
    assert got.shape == want.shape
    assert rel_l2(got, want) <= 1e-5
    got.backward(d.to(DEV))
    assert rel_l2(s_gpu.grad, s_ref.grad) <= 1e-5
    assert rel_l2(voc.to_flat.weight.grad, w_grad_ref) <= 1e-5
