"""ORACLE / TEST INFRASTRUCTURE.

Run in the BUILD container only (needs /root/reference):

    python oracle/make_golden.py

1. Imports the UNMODIFIED reference package from /root/reference with
   `oracle/a_unet` standing in for the absent third-party `a_unet`.
2. Proves `oracle/reference_port.py` bit-identical to it (same weights, same
   seeds) for: net forward, VDiffusion loss + gradients, VSampler, CFG sampling,
   DiffusionUpsampler.sample / forward, DiffusionVocoder.sample.
3. Writes the self-pinned golden vectors to tests/golden/*.npz (the reference
   ships none -- SURVEY.md section 4) together with the error of the same
   oracle evaluated under bf16 autocast, which defines the stated bf16
   tolerance of the CUDA path (DESIGN.md "Tolerance").
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)              # a_unet -> oracle/a_unet
sys.path.insert(0, "/root/reference")  # audio_diffusion_pytorch -> the real reference

import audio_diffusion_pytorch as ref  # noqa: E402

import reference_port as port  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

TINY = dict(in_channels=2, channels=[8, 32, 64], factors=[1, 4, 4], items=[1, 2, 2],
            attentions=[0, 0, 1], attention_heads=2, attention_features=64)
TINY_TEXT = dict(TINY, cross_attentions=[0, 1, 1], use_embedding_cfg=True,
                 embedding_max_length=8, embedding_features=32)
TINY_NOATT = dict(channels=[8, 32, 64], factors=[1, 4, 4], items=[1, 2, 2])
T_LEN, BATCH = 4096, 2


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def same(a, b, what):
    assert torch.equal(a, b), f"port != reference for {what}: max diff {(a - b).abs().max()}"
    print(f"  port == reference (bit-exact): {what}")


def param_fingerprint(module):
    ps = [p.detach().double() for p in module.parameters()]
    return np.array([sum(float(p.sum()) for p in ps), sum(float(p.abs().sum()) for p in ps),
                     float(sum(p.numel() for p in ps))])


def bf16_eval(fn):
    with torch.autocast("cpu", dtype=torch.bfloat16):
        return fn().float()


def case_unconditional():
    torch.manual_seed(0)
    m_ref = ref.DiffusionModel(net_t=ref.UNetV0, diffusion_t=ref.VDiffusion,
                               sampler_t=ref.VSampler, **TINY)
    torch.manual_seed(0)
    m_port = port.DiffusionModelPort(**TINY)
    m_port.load_state_dict(m_ref.state_dict(), strict=True)
    same(torch.cat([p.flatten() for p in m_ref.parameters()]),
         torch.cat([p.flatten() for p in m_port.parameters()]), "same-seed construction")

    g = torch.Generator().manual_seed(1)
    x = torch.randn(BATCH, 2, T_LEN, generator=g)
    sig = torch.rand(BATCH, generator=g)
    with torch.no_grad():
        v_ref, v_port = m_ref.net(x, sig), m_port.net(x, sig)
    same(v_port, v_ref, "UNetV0 forward")

    torch.manual_seed(2)
    loss_ref = m_ref(x)
    loss_ref.backward()
    torch.manual_seed(2)
    loss_port = m_port(x)
    loss_port.backward()
    same(loss_port.detach(), loss_ref.detach(), "VDiffusion loss")
    for (n, p), q in zip(m_ref.named_parameters(), m_port.parameters()):
        assert torch.equal(p.grad, q.grad), n
    print("  port == reference (bit-exact): every parameter gradient")
    grads = {n: p.grad for n, p in m_port.named_parameters()}
    gnorm = torch.stack([g_.norm() for g_ in grads.values()])
    pick = sorted(grads, key=lambda n: grads[n].numel())[:4] + \
        [n for n in grads if grads[n].numel() in (8 * 8 * 3, 32 * 8 * 4)][:2]

    noise = torch.randn(BATCH, 2, T_LEN, generator=g)
    s_ref = m_ref.sample(noise, num_steps=5)
    s_port = m_port.sample(noise, num_steps=5)
    same(s_port, s_ref, "VSampler 5 steps")

    with torch.no_grad():
        v_bf16 = bf16_eval(lambda: m_port.net(x, sig))
        s_bf16 = bf16_eval(lambda: m_port.sample(noise, num_steps=5))
    # the net output is skip(x) + gate*branch, dominated by x at init: also pin the branch
    print(f"  bf16-autocast oracle error: net {rel_l2(v_bf16, v_ref):.3e}  "
          f"branch(v-x) {rel_l2(v_bf16 - x, v_ref - x):.3e}  "
          f"sampler {rel_l2(s_bf16, s_ref):.3e}  |v-x|/|v| {rel_l2(v_ref - x, v_ref) :.3e}")
    np.savez_compressed(
        os.path.join(OUT, "tiny_unconditional.npz"),
        x=x.numpy(), sigma=sig.numpy(), v=v_ref.numpy(), loss=loss_ref.detach().numpy(),
        grad_norms=gnorm.numpy(), noise=noise.numpy(), sample5=s_ref.numpy(),
        param_fingerprint=param_fingerprint(m_port),
        bf16_err_net=rel_l2(v_bf16, v_ref), bf16_err_sample5=rel_l2(s_bf16, s_ref),
        bf16_err_branch=rel_l2(v_bf16 - x, v_ref - x),
        **{"grad:" + n: grads[n].numpy() for n in pick if n in grads})


def case_text_cfg():
    torch.manual_seed(0)
    m_ref = ref.DiffusionModel(net_t=ref.UNetV0, diffusion_t=ref.VDiffusion,
                               sampler_t=ref.VSampler, **TINY_TEXT)
    m_port = port.DiffusionModelPort(**TINY_TEXT)
    m_port.load_state_dict(m_ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(BATCH, 2, T_LEN, generator=g)
    sig = torch.rand(BATCH, generator=g)
    emb = torch.randn(BATCH, 8, 32, generator=g)
    with torch.no_grad():
        v1_ref = m_ref.net(x, sig, embedding=emb)
        v1_port = m_port.net(x, sig, embedding=emb)
        v5_ref = m_ref.net(x, sig, embedding=emb, embedding_scale=5.0)
        v5_port = m_port.net(x, sig, embedding=emb, embedding_scale=5.0)
    same(v1_port, v1_ref, "text-cond forward (scale 1)")
    same(v5_port, v5_ref, "text-cond forward (CFG scale 5)")
    noise = torch.randn(BATCH, 2, T_LEN, generator=g)
    s_ref = m_ref.sample(noise, num_steps=3, embedding=emb, embedding_scale=5.0)
    s_port = m_port.sample(noise, num_steps=3, embedding=emb, embedding_scale=5.0)
    same(s_port, s_ref, "CFG sampler 3 steps")
    with torch.no_grad():
        v5_bf16 = bf16_eval(lambda: m_port.net(x, sig, embedding=emb, embedding_scale=5.0))
    np.savez_compressed(
        os.path.join(OUT, "tiny_text_cfg.npz"), x=x.numpy(), sigma=sig.numpy(),
        embedding=emb.numpy(), v_scale1=v1_ref.numpy(), v_scale5=v5_ref.numpy(),
        noise=noise.numpy(), sample3=s_ref.numpy(), param_fingerprint=param_fingerprint(m_port),
        bf16_err_net=rel_l2(v5_bf16, v5_ref))


def case_upsampler():
    torch.manual_seed(0)
    m_ref = ref.DiffusionUpsampler(net_t=ref.UNetV0, upsample_factor=16, in_channels=2,
                                   diffusion_t=ref.VDiffusion, sampler_t=ref.VSampler,
                                   **TINY_NOATT)
    m_port = port.DiffusionUpsamplerPort(upsample_factor=16, in_channels=2, **TINY_NOATT)
    m_port.load_state_dict(m_ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(4)
    low = torch.randn(BATCH, 2, T_LEN // 16, generator=g)
    s_ref = m_ref.sample(low, num_steps=3, generator=torch.Generator().manual_seed(5))
    s_port = m_port.sample(low, num_steps=3, generator=torch.Generator().manual_seed(5))
    same(s_port, s_ref, "DiffusionUpsampler.sample")
    audio = torch.randn(BATCH, 2, T_LEN, generator=g)
    torch.manual_seed(6)
    l_ref = m_ref(audio)
    torch.manual_seed(6)
    l_port = m_port(audio)
    same(l_port.detach(), l_ref.detach(), "DiffusionUpsampler.forward loss")
    l_port.backward()
    gn = torch.stack([p.grad.norm() for p in m_port.parameters()])
    np.savez_compressed(
        os.path.join(OUT, "tiny_upsampler.npz"), low=low.numpy(), sample3=s_ref.numpy(),
        audio=audio.numpy(), loss=l_ref.detach().numpy(), grad_norms=gn.numpy(),
        reupsampled=m_ref.reupsample(audio).numpy(),
        param_fingerprint=param_fingerprint(m_port))


def case_vocoder():
    kw = dict(mel_n_fft=64, mel_channels=8, mel_sample_rate=48000, mel_normalize_log=True,
              **TINY_NOATT)
    torch.manual_seed(0)
    m_ref = ref.DiffusionVocoder(net_t=ref.UNetV0, diffusion_t=ref.VDiffusion,
                                 sampler_t=ref.VSampler, **kw)
    m_port = port.DiffusionVocoderPort(**kw)
    m_port.load_state_dict(m_ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(7)
    mel = torch.randn(BATCH, 2, 8, T_LEN // 16, generator=g)
    s_ref = m_ref.sample(mel, num_steps=3, generator=torch.Generator().manual_seed(8))
    s_port = m_port.sample(mel, num_steps=3, generator=torch.Generator().manual_seed(8))
    same(s_port, s_ref, "DiffusionVocoder.sample")
    audio = torch.randn(BATCH, 2, T_LEN, generator=g)
    torch.manual_seed(9)
    l_ref = m_ref(audio)
    torch.manual_seed(9)
    l_port = m_port(audio)
    same(l_port.detach(), l_ref.detach(), "DiffusionVocoder.forward loss")
    np.savez_compressed(
        os.path.join(OUT, "tiny_vocoder.npz"), mel=mel.numpy(), sample3=s_ref.numpy(),
        audio=audio.numpy(), loss=l_ref.detach().numpy(),
        param_fingerprint=param_fingerprint(m_port))


def case_sampler_algebra():
    """VSampler with a closed-form toy net: pins the step algebra independent of the U-Net."""
    class Toy(torch.nn.Module):
        def forward(self, x, t, **kw):
            return 0.5 * x * t.view(-1, 1, 1) + torch.sin(x)

    g = torch.Generator().manual_seed(10)
    x = torch.randn(3, 2, 64, generator=g)
    s_ref = ref.VSampler(net=Toy())(x, num_steps=7)
    s_port = port.VSamplerPort(net=Toy())(x, num_steps=7)
    same(s_port, s_ref, "VSampler algebra (toy net)")
    np.savez_compressed(os.path.join(OUT, "sampler_toy.npz"), x=x.numpy(), out7=s_ref.numpy())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for case in (case_sampler_algebra, case_unconditional, case_text_cfg, case_upsampler,
                 case_vocoder):
        print(case.__name__)
        case()
    print("golden vectors written to", OUT)
