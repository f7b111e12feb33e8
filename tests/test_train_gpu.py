"""Training step parity: fused VDiffusion loss + hand-written backward of the B200 U-Net against
torch.autograd through the CPU oracle (same weights, same x / noise / sigma).  Attention-free
configs (BASELINE cfg4/cfg5 shape).  bf16 storage: loss within 2e-3 relative, every parameter
gradient within 6e-2 rel-L2 (and the global gradient direction within 1e-3 cosine distance)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
CFG = dict(in_channels=2, channels=[8, 32, 64], factors=[1, 4, 4], items=[1, 2, 2])


def oracle_loss(ref_net, x, noise, sigma, **kw):
    a, b = torch.cos(sigma * math.pi / 2)[:, None, None], torch.sin(sigma * math.pi / 2)[:, None, None]
    return F.mse_loss(ref_net(a * x + b * noise, sigma, **kw), a * noise - b * x)


def compare_grads(ref_params, got_params):
    worst, dots, n1, n2 = 0.0, 0.0, 0.0, 0.0
    # gradients that are analytically zero (a conv bias feeding a GroupNorm with one channel per
    # group) are compared on the scale of a typical parameter gradient, not on their own
    floor = 0.1 * float(torch.stack([p.grad.double().norm() for _, p in ref_params]).median())
    for (name, p), q in zip(ref_params, got_params):
        assert q.grad is not None, f"no gradient for {name}"
        g_ref, g = p.grad.double(), q.grad.double().cpu()
        rel = float((g - g_ref).norm() / g_ref.norm().clamp_min(floor))
        worst = max(worst, rel)
        dots += float((g * g_ref).sum()); n1 += float((g * g).sum()); n2 += float((g_ref * g_ref).sum())
        if rel > 6e-2:
            print(f"  {name:60s} shape {tuple(p.shape)} rel-L2 {rel:.3e}")
    cos = dots / math.sqrt(n1 * n2)
    print(f"worst per-parameter rel-L2 {worst:.3e}; global cosine {cos:.6f}")
    return worst, cos


@pytest.mark.parametrize("upsampler", [False, True])
def test_loss_and_gradients_match_oracle(oracle_port, upsampler):
    import audio_diffusion_pytorch_b200 as adp
    from audio_diffusion_pytorch_b200.training import fused_v_loss
    torch.manual_seed(0)
    if upsampler:
        kw = {k: v for k, v in CFG.items() if k != "in_channels"}
        ref = oracle_port.DiffusionUpsamplerPort(upsample_factor=16, in_channels=2, **kw)
        model = adp.DiffusionUpsampler(net_t=adp.UNetV0, upsample_factor=16, in_channels=2, **kw).to(DEV)
    else:
        ref = oracle_port.DiffusionModelPort(**CFG)
        model = adp.DiffusionModel(net_t=adp.UNetV0, **CFG).to(DEV)
    model.net.load_reference_parameters(ref.net)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, 4096, generator=g)
    noise = torch.randn(2, 2, 4096, generator=g)
    sigma = torch.rand(2, generator=g)
    extra_ref, extra = {}, {}
    if upsampler:
        app = ref.reupsample(x)
        extra_ref, extra = dict(append_channels=app), dict(append_channels=app.to(DEV))
    loss_ref = oracle_loss(ref.net, x, noise, sigma, **extra_ref)
    loss_ref.backward()
    for call in range(3):                       # eager, graph capture, graph replay
        model.zero_grad(set_to_none=True)
        loss = fused_v_loss(model.net, x.to(DEV), noise.to(DEV), sigma.to(DEV), **extra)
        loss.backward()
        rel = abs(float(loss) - float(loss_ref)) / float(loss_ref)
        print(f"call {call}: loss {float(loss):.6f} vs oracle {float(loss_ref):.6f} (rel {rel:.2e})")
        assert rel < 2e-3
        worst, cos = compare_grads(list(ref.net.named_parameters()), list(model.net.parameters()))
        assert worst < 6e-2 and cos > 1 - 1e-3


def test_optimizer_step_refreshes_packed_weights(oracle_port):
    """After optimizer.step() the next loss must see the new weights (in-place re-pack)."""
    import audio_diffusion_pytorch_b200 as adp
    torch.manual_seed(0)
    ref = oracle_port.DiffusionModelPort(**CFG)
    model = adp.DiffusionModel(net_t=adp.UNetV0, **CFG).to(DEV)
    model.net.load_reference_parameters(ref.net)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.05)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 2, 4096, generator=g)
    losses, losses_ref = [], []
    for step in range(4):
        torch.manual_seed(100 + step)
        sigma = torch.rand(2)
        noise = torch.randn(2, 2, 4096)
        from audio_diffusion_pytorch_b200.training import fused_v_loss
        opt.zero_grad(); opt_ref.zero_grad()
        loss = fused_v_loss(model.net, x.to(DEV), noise.to(DEV), sigma.to(DEV))
        loss.backward(); opt.step()
        lr_ = oracle_loss(ref.net, x, noise, sigma)
        lr_.backward(); opt_ref.step()
        losses.append(float(loss)); losses_ref.append(float(lr_))
    print("losses", losses, "oracle", losses_ref)
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) / b < 5e-3


def test_model_forward_is_the_reference_call(oracle_port):
    """`loss = model(x); loss.backward()` -- the reference's own training call (README.md:37-38)."""
    import audio_diffusion_pytorch_b200 as adp
    torch.manual_seed(0)
    model = adp.DiffusionModel(net_t=adp.UNetV0, **CFG).to(DEV)
    x = torch.randn(2, 2, 4096, device=DEV)
    loss = model(x)
    loss.backward()
    assert loss.ndim == 0 and torch.isfinite(loss)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
