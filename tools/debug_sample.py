import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_port as port
import audio_diffusion_pytorch_b200 as adp
cfg = dict(in_channels=2, channels=[8, 32, 64], factors=[1, 4, 4], items=[1, 2, 2],
           attentions=[0, 0, 1], attention_heads=2, attention_features=64)
torch.manual_seed(0)
ref = port.DiffusionModelPort(**cfg)
x = torch.randn(2, 2, 4096, generator=torch.Generator().manual_seed(1))
for graph in (False, True):
    for steps in (1, 2, 3, 5):
        model = adp.DiffusionModel(net_t=adp.UNetV0, **cfg).cuda()
        model.net.load_reference_parameters(ref.net)
        model.net.use_cuda_graph = graph
        s_ref = ref.sample(x, num_steps=steps)
        s = model.sample(x.cuda(), num_steps=steps).cpu()
        s2 = model.sample(x.cuda(), num_steps=steps).cpu()
        e = float((s - s_ref).norm() / s_ref.norm()); e2 = float((s2 - s_ref).norm() / s_ref.norm())
        print(f"graph={graph} steps={steps}: first call err {e:.3e}  second call err {e2:.3e}", flush=True)
