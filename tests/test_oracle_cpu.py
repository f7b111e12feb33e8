"""The CPU oracle (oracle/reference_port.py over the oracle/a_unet shim) against the committed
golden vectors.  The vectors were produced by oracle/make_golden.py in the build container, where
the same script also proved the port bit-identical to the UNMODIFIED reference package; here
(and on the GPU box, where /root/reference does not exist) the oracle is re-checked against them
before any CUDA result is compared with it.  Same seeds as make_golden.py; tolerances are a few
fp32 ulps of accumulated reordering noise (thread count / BLAS blocking may differ between
machines), not bit-exactness."""
import numpy as np
import pytest
import torch

TINY = dict(in_channels=2, channels=[8, 32, 64], factors=[1, 4, 4], items=[1, 2, 2],
            attentions=[0, 0, 1], attention_heads=2, attention_features=64)
TINY_TEXT = dict(TINY, cross_attentions=[0, 1, 1], use_embedding_cfg=True,
                 embedding_max_length=8, embedding_features=32)
TINY_NOATT = dict(channels=[8, 32, 64], factors=[1, 4, 4], items=[1, 2, 2])
RTOL = 2e-5        # relative L2


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def fingerprint(module):
    ps = [p.detach().double() for p in module.parameters()]
    return np.array([sum(float(p.sum()) for p in ps), sum(float(p.abs().sum()) for p in ps),
                     float(sum(p.numel() for p in ps))])


def load(golden_dir, name):
    return {k: v for k, v in np.load(f"{golden_dir}/{name}").items()}


def t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(autouse=True)
def _threads():
    n = torch.get_num_threads()
    torch.set_num_threads(min(n, 8))
    yield
    torch.set_num_threads(n)


def test_sampler_algebra_toy_net(oracle_port, golden_dir):
    """VSampler update algebra (reference diffusion.py:183-188) with a closed-form net."""
    g = load(golden_dir, "sampler_toy.npz")

    class Toy(torch.nn.Module):
        def forward(self, x, t, **kw):
            return 0.5 * x * t.view(-1, 1, 1) + torch.sin(x)

    out = oracle_port.VSamplerPort(net=Toy())(t(g["x"]), num_steps=7)
    assert rel_l2(out, t(g["out7"])) <= 1e-6


def test_unconditional_forward_loss_grads_sampler(oracle_port, golden_dir):
    g = load(golden_dir, "tiny_unconditional.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionModelPort(**TINY)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    x, sig = t(g["x"]), t(g["sigma"])
    with torch.no_grad():
        v = m.net(x, sig)
    assert rel_l2(v, t(g["v"])) <= RTOL
    torch.manual_seed(2)
    loss = m(x)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    gnorm = torch.stack([p.grad.norm() for p in m.parameters()])
    assert rel_l2(gnorm, t(g["grad_norms"])) <= 1e-4
    grads = dict(m.named_parameters())
    for key in [k for k in g if k.startswith("grad:")]:
        got = grads[key[5:]].grad
        assert rel_l2(got, t(g[key])) <= 1e-3, key
    s = m.sample(t(g["noise"]), num_steps=5)
    assert rel_l2(s, t(g["sample5"])) <= 1e-4


def test_text_conditioning_and_cfg(oracle_port, golden_dir):
    g = load(golden_dir, "tiny_text_cfg.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionModelPort(**TINY_TEXT)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    x, sig, emb = t(g["x"]), t(g["sigma"]), t(g["embedding"])
    with torch.no_grad():
        v1 = m.net(x, sig, embedding=emb)
        v5 = m.net(x, sig, embedding=emb, embedding_scale=5.0)
    assert rel_l2(v1, t(g["v_scale1"])) <= RTOL
    assert rel_l2(v5, t(g["v_scale5"])) <= RTOL
    s = m.sample(t(g["noise"]), num_steps=3, embedding=emb, embedding_scale=5.0)
    assert rel_l2(s, t(g["sample3"])) <= 1e-4


def test_upsampler_sample_loss_reupsample(oracle_port, golden_dir):
    g = load(golden_dir, "tiny_upsampler.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionUpsamplerPort(upsample_factor=16, in_channels=2, **TINY_NOATT)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    s = m.sample(t(g["low"]), num_steps=3, generator=torch.Generator().manual_seed(5))
    assert rel_l2(s, t(g["sample3"])) <= 1e-4
    audio = t(g["audio"])
    torch.manual_seed(6)
    loss = m(audio)
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_l2(m.reupsample(audio), t(g["reupsampled"])) <= 1e-6


def test_vocoder_sample_and_loss(oracle_port, golden_dir):
    g = load(golden_dir, "tiny_vocoder.npz")
    kw = dict(mel_n_fft=64, mel_channels=8, mel_sample_rate=48000, mel_normalize_log=True, **TINY_NOATT)
    torch.manual_seed(0)
    m = oracle_port.DiffusionVocoderPort(**kw)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    s = m.sample(t(g["mel"]), num_steps=3, generator=torch.Generator().manual_seed(8))
    assert rel_l2(s, t(g["sample3"])) <= 1e-4
    torch.manual_seed(9)
    loss = m(t(g["audio"]))
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))


# ------------------------------------------------------------------ round-2 golden vectors
def test_oracle_50_step_sampler_and_inpainter(oracle_port, golden_dir):
    """tiny_sample50.npz / tiny_inpaint.npz (oracle/make_golden_r2.py): the headline metric is a
    50-step sample; VInpainter consumes torch.randn_like draws in a fixed order."""
    g = load(golden_dir, "tiny_sample50.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionModelPort(**TINY)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    noise = torch.randn(2, 2, 4096, generator=torch.Generator().manual_seed(int(g["noise_seed"])))
    assert rel_l2(m.sample(noise, num_steps=50), t(g["sample50"])) <= 20 * RTOL   # 50 steps compound

    g = load(golden_dir, "tiny_inpaint.npz")
    source = torch.randn(2, 2, 4096, generator=torch.Generator().manual_seed(int(g["source_seed"])))
    mask = torch.zeros(2, 2, 4096, dtype=torch.bool)
    for b_, lo, hi in g["mask_spans"]:
        mask[b_, :, lo:hi] = True
    torch.manual_seed(int(g["rng_seed"]))
    out = oracle_port.VInpainterPort(net=m.net)(source, mask, num_steps=int(g["num_steps"]),
                                                num_resamples=int(g["num_resamples"]))
    assert rel_l2(out, t(g["out"])) <= 10 * RTOL
    assert torch.equal(out[mask], source[mask])        # sigma = 0 at the end: the known region is the source


def test_oracle_readme_scale_goldens(oracle_port, golden_dir):
    """readme_full_size.npz (forward at [1,2,2^18], windows) and cfg3_readme_scale.npz (text +
    guidance at README scale): the oracle re-evaluated on this machine's CPU."""
    README = dict(in_channels=2, channels=[8, 32, 64, 128, 256, 512, 512, 1024, 1024],
                  factors=[1, 4, 4, 4, 2, 2, 2, 2, 2], items=[1, 2, 2, 2, 2, 2, 2, 4, 4],
                  attentions=[0, 0, 0, 0, 0, 1, 1, 1, 1], attention_heads=8, attention_features=64)
    g = load(golden_dir, "readme_full_size.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionModelPort(**README)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    x = torch.randn(1, 2, 2 ** 18, generator=torch.Generator().manual_seed(int(g["x_seed"])))
    with torch.no_grad():
        v = m.net(x, t(g["sigma"]))
    win = torch.stack([v[..., int(s):int(s) + 1024] for s in g["starts"]], dim=-2)
    assert rel_l2(win, t(g["v_windows"])) <= RTOL
    del m

    g = load(golden_dir, "cfg3_readme_scale.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionModelPort(**dict(README, cross_attentions=[0, 0, 0, 1, 1, 1, 1, 1, 1],
                                              use_embedding_cfg=True, embedding_max_length=64,
                                              embedding_features=768))
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    gen = torch.Generator().manual_seed(int(g["seed"]))
    x = torch.randn(1, 2, int(g["length"]), generator=gen)
    emb = torch.randn(1, 64, 768, generator=gen)
    with torch.no_grad():
        v5 = m.net(x, t(g["sigma"]), embedding=emb, embedding_scale=5.0)
    assert rel_l2(v5, t(g["v_scale5"])) <= RTOL


def test_oracle_autoencoder(oracle_port, golden_dir):
    """tiny_autoencoder.npz: DiffusionAE loss, encoder gradient and 3-step decode."""
    g = load(golden_dir, "tiny_autoencoder.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionAEPort(encoder=oracle_port.ToyEncoder(), **dict(TINY, inject_depth=2))
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    audio = torch.randn(2, 2, 4096, generator=torch.Generator().manual_seed(int(g["audio_seed"])))
    torch.manual_seed(int(g["loss_seed"]))
    loss = m(audio)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_l2(m.encoder.conv.weight.grad, t(g["enc_grad"])) <= 10 * RTOL
    latent = m.encoder(audio).detach()
    out = m.decode(latent, num_steps=3, generator=torch.Generator().manual_seed(int(g["decode_seed"])))
    assert rel_l2(out, t(g["decode3"])) <= 10 * RTOL


TINY_AR = dict(TINY, in_channels=2, length=4096, num_splits=4)


def test_oracle_autoregressive(oracle_port, golden_dir):
    """tiny_autoregressive.npz: DiffusionAR = use_modulation=False net (SkipCat merges) with
    ARVDiffusion loss / gradients and the ARVSampler ladder (start window + 2 shifts)."""
    g = load(golden_dir, "tiny_autoregressive.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionARPort(**TINY_AR)
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    audio = torch.randn(2, 2, 4096, generator=gen)
    chan = torch.cat([audio, torch.rand(2, 1, 4096, generator=gen)], dim=1)
    with torch.no_grad():
        assert rel_l2(m.net(chan), t(g["v"])) <= RTOL
    torch.manual_seed(int(g["loss_seed"]))
    loss = m(audio)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    norms = np.array([float(p.grad.norm()) for p in m.net.parameters()])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-3, atol=1e-7)
    assert rel_l2(list(m.net.parameters())[-2].grad, t(g["grad_net_skip_merge_blocks_0_weight"])) <= 10 * RTOL
    torch.manual_seed(int(g["sample_seed"]))
    out = m.sample(num_items=2, num_chunks=6, num_steps=4)
    assert out.shape == (2, 2, 6 * 1024)
    assert rel_l2(out, t(g["sample"])) <= 10 * RTOL
    assert torch.equal(m.sampler.sigmas_ladder(2, 3)[:, 0, 0], t(g["ladder3"]))


def test_package_sigma_ladder_matches_reference(golden_dir):
    """Host logic of the product's ARVSampler (no kernels): the staircase of per-split noise
    levels equals the reference's (reference diffusion.py:213-221), value for value."""
    import torch.nn as nn
    from audio_diffusion_pytorch_b200.diffusion import ARVSampler
    g = load(golden_dir, "tiny_autoregressive.npz")
    s = ARVSampler(net=nn.Linear(1, 1), in_channels=2, length=4096, num_splits=4)
    lad = s.get_sigmas_ladder(num_items=2, num_steps_per_split=3)
    assert lad.shape == (4, 2, 1, 4096)
    assert torch.equal(lad[:, 0, 0], t(g["ladder3"])) and torch.equal(lad[:, 1, 0], t(g["ladder3"]))


LT = dict(num_filters=4, window_length=8, stride=4)


def test_oracle_learned_transform(oracle_port, golden_dir):
    """tiny_learned_transform.npz: LTPlugin (reference components.py:113-157) around the net."""
    g = load(golden_dir, "tiny_learned_transform.npz")
    torch.manual_seed(0)
    m = oracle_port.DiffusionModelPort(net_t=oracle_port.lt_plugin(oracle_port.build_unet_v0, **LT),
                                       **dict(TINY, in_channels=1))
    np.testing.assert_allclose(fingerprint(m), g["param_fingerprint"], rtol=1e-9)
    gen = torch.Generator().manual_seed(int(g["seed"]))
    x = torch.randn(2, 1, 16384, generator=gen)
    sig = torch.rand(2, generator=gen)
    with torch.no_grad():
        assert rel_l2(m.net(x, sig), t(g["v"])) <= RTOL
    torch.manual_seed(int(g["loss_seed"]))
    loss = m(x)
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    ps = list(m.net.parameters())
    assert rel_l2(ps[0].grad, t(g["encode_grad"])) <= 10 * RTOL
    assert rel_l2(ps[1].grad, t(g["decode_grad"])) <= 10 * RTOL
    noise = torch.randn(2, 1, 16384, generator=gen)
    assert rel_l2(m.sample(noise, num_steps=3), t(g["sample3"])) <= 10 * RTOL
