"""Host-side logic that needs no GPU: weight packers (the algebra the kernels rely on), the
liveness pool, the product path's refusal to run without a CUDA device."""
import pytest
import torch
import torch.nn.functional as F


@pytest.fixture(scope="module")
def ops():
    from audio_diffusion_pytorch_b200 import ops
    return ops


def test_pack_conv_is_tap_major(ops):
    w = torch.randn(24, 16, 3)
    p = ops.pack_conv(w).float()
    assert p.shape == (32, 48)                         # rows padded to 16
    for tap in range(3):
        assert torch.equal(p[:24, tap * 16:(tap + 1) * 16], w[:, :, tap].to(torch.bfloat16).float())
    assert torch.count_nonzero(p[24:]) == 0


def test_pack_mid_conv_layout(ops):
    w = torch.randn(32, 32, 3)
    p = ops.pack_mid_conv(w).float()
    assert p.shape == (32, 96)
    for tap in range(3):
        assert torch.equal(p[:, tap * 32:(tap + 1) * 32], w[:, :, tap].to(torch.bfloat16).float())


@pytest.mark.parametrize("f", [2, 4])
def test_upsample_fold_equals_upsample_then_conv(ops, f):
    """nearest-upsample(f) + conv3 == f phases x <= 2 taps on the low-res signal with pre-summed
    weights (the a_unet Upsample block, evaluated the way adp_conv_gemm(up_factor=f) does)."""
    torch.manual_seed(0)
    co, ci, T = 8, 16, 12
    w = torch.randn(co, ci, 3).to(torch.bfloat16).float()      # bf16-exact values: sums exact enough
    x = torch.randn(1, ci, T)
    ref = F.conv1d(F.interpolate(x, scale_factor=f, mode="nearest"), w, padding=1)   # [1, co, T*f]
    packed = ops.pack_upsample_conv(w, f).float()
    n_pad = packed.shape[0] // f
    xp = F.pad(x, (1, 1))                                       # zero rows outside [0, T)
    out = torch.zeros(1, co, T * f)
    for p in range(f):
        wp = packed[p * n_pad:p * n_pad + co]                   # [co, 2*ci]: tap slot 0, 1
        if p == 0:
            offs = (-1, 0)
        elif p == f - 1:
            offs = (0, 1)
        else:
            offs = (0,)
        acc = torch.zeros(1, co, T)
        for slot, off in enumerate(offs):
            xs = xp[:, :, 1 + off:1 + off + T]
            acc += torch.einsum("oc,bct->bot", wp[:, slot * ci:(slot + 1) * ci], xs)
        out[:, :, p::f] = acc
    err = float((out - ref).norm() / ref.norm())                # pre-summed weights are rounded to bf16
    assert err <= 5e-3, err


def test_pack_conv_dgrad_is_the_transposed_conv(ops):
    """Data gradient of conv3 = conv3 of the output gradient with flipped, transposed weights."""
    torch.manual_seed(1)
    co, ci, T = 16, 16, 20
    w = torch.randn(co, ci, 3).to(torch.bfloat16).float()
    x = torch.randn(1, ci, T, requires_grad=True)
    g = torch.randn(1, co, T)
    F.conv1d(x, w, padding=1).backward(g)
    p = ops.pack_conv_dgrad(w).float()[:ci]                     # [ci, 3*co] tap-major
    wt = torch.stack([p[:, k * co:(k + 1) * co] for k in range(3)], dim=2)   # [ci, co, 3]
    dx = F.conv1d(g, wt, padding=1)
    assert torch.allclose(dx, x.grad, rtol=1e-4, atol=1e-4)


def test_product_path_refuses_to_run_on_cpu():
    """No CPU / PyTorch fallback: a CPU tensor must raise, never silently compute."""
    import audio_diffusion_pytorch_b200 as adp
    model = adp.DiffusionModel(net_t=adp.UNetV0, in_channels=2, channels=[8, 32], factors=[1, 4],
                               items=[1, 1], attentions=[0, 0])
    with pytest.raises((AssertionError, RuntimeError)):
        model.sample(torch.randn(1, 2, 64), num_steps=2)
