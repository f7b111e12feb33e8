"""Training step of the hot path: VDiffusion loss + backward through the B200 U-Net
(reference diffusion.py:82-95, `loss = model(x); loss.backward()`).

`fused_v_loss` returns the same scalar as the reference (`F.mse_loss(net(x_noisy, sigma),
v_target)`) as a tensor wired into autograd by ONE custom Function: its backward runs the
hand-written backward program (data-gradient GEMMs = adp_conv_gemm with transposed weights,
weight-gradient GEMMs = adp_wgrad, GroupNorm / LayerNorm-FiLM / stem backward kernels) and
hands every parameter its fp32 gradient in PyTorch layout, so optimizers, gradient clipping
and DistributedDataParallel (NCCL all-reduce of `.grad` buckets) work unchanged.

The forward noising (alpha*x + beta*noise) is fused into the first kernel and the MSE +
dL/dv into the last one.  Only the 3 tiny time-embedding linears run as PyTorch ops (their
autograd supplies d(features); [B,1024] matrices, < 0.1 % of the step).

Scope: attention-free U-Nets.  A net with AttentionItems raises --
attention backward is not built yet and nothing falls back silently.
"""
from math import pi
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .unet import B200UNet, LevelParams, _pad_to


class _TrainPlan:
    def __init__(self):
        self.fwd: List = []
        self.bwd: List = []          # appended in forward order, executed reversed
        self.post: List = []         # runs after the reversed list (conditioning backward)
        self.graph_f = self.graph_b = None
        self.runs_f = self.runs_b = 0


def _zeros(shape, dev, dtype=torch.float32):
    return torch.zeros(shape, device=dev, dtype=dtype)


def build_train_plan(net: B200UNet, B: int, T: int) -> _TrainPlan:
    assert not any(net.attentions) and not any(net.cross_attentions), \
        "training through AttentionItems is not built yet (attention backward)"
    dev = net.net.down.weight.device
    P = net.packed()
    plan = _TrainPlan()
    G, Fm = net.groups, net.features
    levels = net.levels()
    bf16 = torch.bfloat16

    def act(*shape):
        return torch.empty(*shape, dtype=bf16, device=dev)

    # ---- static I/O
    plan.x = _zeros((B, net.x_channels, T), dev)
    plan.noise = _zeros((B, net.x_channels, T), dev)
    plan.append = _zeros((B, net.append_channels, T), dev) if net.append_channels else None
    plan.alpha, plan.beta = _zeros((B,), dev), _zeros((B,), dev)
    plan.cond = _zeros((B, Fm), dev)                      # SiLU(features), fp32 master
    plan.cond_bf = torch.zeros(1, B, Fm, dtype=bf16, device=dev)
    plan.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    plan.dv = _zeros((B, net.out_channels, T), dev)
    plan.gscale = torch.ones(1, device=dev)
    plan.dcond = _zeros((B, Fm), dev)

    # ---- statistics + gradient arenas
    n_items = sum(len(lv.items_down) + len(lv.items_up) for lv in levels)
    arena = torch.zeros(4 * n_items + 4 * len(levels) + 8, B, G, 2, dtype=torch.float64, device=dev)
    slot = [0]

    def new_stats():
        s = arena[slot[0]]
        slot[0] += 1
        return s

    plan.fwd.append(lambda: (arena.zero_(), plan.loss_sum.zero_()))
    refreshers: List = []                  # re-pack the dgrad weights in place after a weight update

    def packed_dgrad(make):
        t = make()
        refreshers.append(lambda t=t, make=make: t.copy_(make()))
        return t

    grads: Dict[int, Tensor] = {}          # id(param) -> fp32 gradient (PyTorch layout / packed)
    finals: List = []                      # closures turning packed grads into PyTorch layout
    # one flat fp32 arena for every gradient accumulator: a single memset per backward
    n_param = sum(p.numel() for p in net.parameters())
    flat = torch.zeros(int(1.25 * n_param) + 64 * (4 * n_param // 1000 + 4096), device=dev)
    cursor = [0]

    def gbuf(shape):
        n = 1
        for d in shape:
            n *= d
        start = cursor[0]
        cursor[0] = start + (n + 63) // 64 * 64
        assert cursor[0] <= flat.numel(), "gradient arena too small"
        return flat[start:start + n].view(*shape)

    def grad_for(param, shape=None):
        t = gbuf(tuple(param.shape) if shape is None else shape)
        grads[id(param)] = t
        return t

    n_tot = P["cond_n"]
    ss_all = _zeros((B, ops.round_up(n_tot, 8)), dev)
    dss_all = gbuf(ss_all.shape)
    ss_stride = ss_all.shape[1]
    cond_bias = _pad_to(P["cond_b"], ss_all.shape[1])
    plan.fwd.append(lambda: plan.cond_bf.copy_(plan.cond.view(1, B, Fm)))
    plan.fwd.append(lambda: ops.conv_gemm(plan.cond_bf, P["cond_w"], ss_all.view(1, B, -1), c_in=Fm,
                                          n_valid=ss_all.shape[1], bias=cond_bias))

    def conv3_bwd(dy: Tensor, a_in: Tensor, gw: Tensor, da: Tensor, wd: Tensor, C: int):
        """dy: grad of the conv output; a_in: its input; writes da, accumulates dW (3 taps)."""
        for tap, off in enumerate((-1, 0, 1)):
            ops.wgrad(dy, a_in, gw[tap], n=C, k=C, off=off)
        ops.conv_gemm(dy, wd, da, c_in=C, n_valid=C, taps=(-1, 0, 1))

    # ---- one chain of ResnetItem+ModulationItem
    def run_items(x: Tensor, x_stats: Tensor, items_p: List[Dict], items_m, lv: LevelParams, Tl: int):
        C = lv.ch
        narrow = C == 8
        for ip, im in zip(items_p, items_m):
            r_ = im.resnet
            ss = ss_all[:, ip["ss_off"]:]
            dss = dss_all[:, ip["ss_off"]:]
            h_stats, y_stats = new_stats(), new_stats()
            S1, S2 = new_stats(), new_stats()
            h, rr, y = act(B, Tl, C), act(B, Tl, C), act(B, Tl, C)
            dgn1 = (grad_for(r_.gn1.weight), grad_for(r_.gn1.bias))
            dgn2 = (grad_for(r_.gn2.weight), grad_for(r_.gn2.bias))
            db1, db2 = grad_for(r_.conv1.bias), grad_for(r_.conv2.bias)
            dr, dh, dx, dxh = act(B, Tl, C), act(B, Tl, C), act(B, Tl, C), act(B, Tl, C)
            if narrow:
                dw1, dw2 = grad_for(r_.conv1.weight), grad_for(r_.conv2.weight)
                db_scratch = gbuf((C,))
                plan.fwd.append(lambda x=x, h=h, s=x_stats, hs=h_stats, ip=ip: ops.narrow_conv(
                    x, h, s, ip["gn1"][0], ip["gn1"][1], ip["w1"], ip["b1"], G, stats_out=hs))
                plan.fwd.append(lambda x=x, h=h, rr=rr, hs=h_stats, ip=ip: ops.narrow_conv(
                    h, rr, hs, ip["gn2"][0], ip["gn2"][1], ip["w2"], ip["b2"], G, residual=x))
                plan.fwd.append(lambda rr=rr, y=y, ss=ss, ys=y_stats: ops.ln_film(
                    rr, y, ss, ss_stride, ys, G, net.MOD_LN_EPS))

                def bwd(dy, x=x, h=h, rr=rr, ss=ss, dss=dss, xs=x_stats, hs=h_stats, ip=ip, dr=dr,
                        dh=dh, dx=dx, dxh=dxh, S1=S1, S2=S2, dgn1=dgn1, dgn2=dgn2, dw1=dw1, dw2=dw2,
                        db1=db1, db2=db2, db_scratch=db_scratch):
                    ops.ln_film_bwd(dy, rr, ss, ss_stride, dr, dss=dss, dss_stride=ss_stride,
                                    eps=net.MOD_LN_EPS)
                    ops.narrow_conv_bwd(dr, h, hs, ip["gn2"][0], ip["gn2"][1], ip["w2"], dxh, dgn2[0],
                                        dgn2[1], S2, dw2, db2, G)
                    ops.gn_bwd_apply(dxh, h, hs, S2, dh, G, colsum=db1)   # fp32 sum, pre-rounding
                    ops.narrow_conv_bwd(dh, x, xs, ip["gn1"][0], ip["gn1"][1], ip["w1"], dxh, dgn1[0],
                                        dgn1[1], S1, dw1, db_scratch, G)
                    ops.gn_bwd_apply(dxh, x, xs, S1, dx, G, dres=dr)
                    return dx
            else:
                a1, a2 = act(B, Tl, C), act(B, Tl, C)
                wd1 = packed_dgrad(lambda r_=r_: ops.pack_conv_dgrad(r_.conv1.weight.detach()))
                wd2 = packed_dgrad(lambda r_=r_: ops.pack_conv_dgrad(r_.conv2.weight.detach()))
                plan.fwd.append(lambda x=x, a1=a1, s=x_stats, ip=ip: ops.gn_silu(
                    x, a1, s, ip["gn1"][0], ip["gn1"][1], G, net.GN_EPS))
                plan.fwd.append(lambda a1=a1, h=h, hs=h_stats, ip=ip: ops.conv_gemm(
                    a1, ip["w1"], h, c_in=C, n_valid=C, taps=(-1, 0, 1), bias=ip["b1"], stats=hs, groups=G))
                plan.fwd.append(lambda a2=a2, h=h, hs=h_stats, ip=ip: ops.gn_silu(
                    h, a2, hs, ip["gn2"][0], ip["gn2"][1], G, net.GN_EPS))
                plan.fwd.append(lambda x=x, a2=a2, rr=rr, ip=ip: ops.conv_gemm(
                    a2, ip["w2"], rr, c_in=C, n_valid=C, taps=(-1, 0, 1), bias=ip["b2"], residual=x))
                plan.fwd.append(lambda rr=rr, y=y, ss=ss, ys=y_stats: ops.ln_film(
                    rr, y, ss, ss_stride, ys, G, net.MOD_LN_EPS))
                da = act(B, Tl, C)
                gw = {"w1": gbuf((3, C, C)), "w2": gbuf((3, C, C))}

                def bwd(dy, x=x, h=h, rr=rr, a1=a1, a2=a2, ss=ss, dss=dss, xs=x_stats, hs=h_stats,
                        ip=ip, dr=dr, dh=dh, dx=dx, dxh=dxh, da=da, S1=S1, S2=S2, dgn1=dgn1,
                        dgn2=dgn2, db1=db1, db2=db2, wd1=wd1, wd2=wd2, gw=gw, C=C):
                    ops.ln_film_bwd(dy, rr, ss, ss_stride, dr, dss=dss, dss_stride=ss_stride,
                                    colsum=db2, eps=net.MOD_LN_EPS)
                    conv3_bwd(dr, a2, gw["w2"], da, wd2, C)
                    ops.gn_silu_bwd(da, h, hs, ip["gn2"][0], ip["gn2"][1], dxh, dgn2[0], dgn2[1], S2, G)
                    ops.gn_bwd_apply(dxh, h, hs, S2, dh, G, colsum=db1)
                    conv3_bwd(dh, a1, gw["w1"], da, wd1, C)
                    ops.gn_silu_bwd(da, x, xs, ip["gn1"][0], ip["gn1"][1], dxh, dgn1[0], dgn1[1], S1, G)
                    ops.gn_bwd_apply(dxh, x, xs, S1, dx, G, dres=dr)
                    return dx
                finals.append(lambda gw=gw, r_=r_: (
                    grads.__setitem__(id(r_.conv1.weight), gw["w1"].permute(1, 2, 0)),
                    grads.__setitem__(id(r_.conv2.weight), gw["w2"].permute(1, 2, 0))))
            grads[id(im.modulation.proj.weight)] = ("cond_w", ip["ss_off"], 2 * C)
            grads[id(im.modulation.proj.bias)] = ("cond_b", ip["ss_off"], 2 * C)
            plan.bwd.append(("item", bwd))
            x, x_stats = y, y_stats
        return x, x_stats

    # ---- recursive level walk; returns (output tensor, its stats, backward closure)
    tape: List = []   # executed in reverse: each entry is a callable(dy) -> dx

    def level(i: int, x_in: Optional[Tensor], T_in: int):
        lv, Lp = levels[i], P["levels"][i]
        Tl, C = T_in // lv.factor, lv.ch
        innermost = i == len(levels) - 1
        x0, st0 = act(B, Tl, C), new_stats()
        db_down = grad_for(lv.down.bias)
        if i == 0:
            dw_down = grad_for(lv.down.weight)
            plan.fwd.append(lambda: ops.stem_in(plan.x, Lp["down_w"], Lp["down_b"], x0, lv.factor,
                                                append=plan.append, noise=plan.noise, alpha=plan.alpha,
                                                beta=plan.beta, stats=st0, groups=G))
        else:
            kdim = lv.factor * lv.in_ch
            plan.fwd.append(lambda: ops.conv_gemm(x_in.view(B, Tl, kdim), Lp["down_w"], x0, c_in=kdim,
                                                  n_valid=C, bias=Lp["down_b"], stats=st0, groups=G))
            wd_down = packed_dgrad(lambda: ops.pack_linear(
                lv.down.weight.detach().permute(0, 2, 1).reshape(C, kdim).t().contiguous()))
            gw_down = gbuf((C, kdim))
            finals.append(lambda: grads.__setitem__(
                id(lv.down.weight), gw_down.view(C, lv.factor, lv.in_ch).permute(0, 2, 1)))
        n_before = len(plan.bwd)
        x, st = run_items(x0, st0, Lp["items_down"], lv.items_down, lv, Tl)
        items_down_bwd = [b for _, b in plan.bwd[n_before:]]
        del plan.bwd[n_before:]
        inner = None
        skip = x
        if not innermost:
            x, st, inner = level(i + 1, skip, Tl)
        n_before = len(plan.bwd)
        x, st = run_items(x, st, Lp["items_up"], lv.items_up, lv, Tl)
        items_up_bwd = [b for _, b in plan.bwd[n_before:]]
        del plan.bwd[n_before:]
        gate = ss_all[:, Lp["gate_off"]:]
        dgate = dss_all[:, Lp["gate_off"]:]
        grads[id(lv.merge.weight)] = ("cond_w", Lp["gate_off"], lv.out_ch)
        grads[id(lv.merge.bias)] = ("cond_b", Lp["gate_off"], lv.out_ch)
        db_up = grad_for(lv.up.bias)
        x_last = x
        if i == 0:
            dw_up = grad_for(lv.up.weight)
            dwa = grad_for(lv.adapter.weight, (lv.out_ch, lv.in_ch)) if lv.adapter is not None else None
            dba = grad_for(lv.adapter.bias) if lv.adapter is not None else None
            if lv.adapter is not None:
                finals.append(lambda: grads.__setitem__(id(lv.adapter.weight), dwa.unsqueeze(-1)))
            dh0 = act(B, Tl, C)
            plan.fwd.append(lambda: ops.stem_out(
                x_last, plan.x, Lp["up_w"], Lp["up_b"], gate, lv.factor, append=plan.append,
                w_adapt=Lp.get("adapt_w"), b_adapt=Lp.get("adapt_b"), noise=plan.noise,
                alpha=plan.alpha, beta=plan.beta, loss_sum=plan.loss_sum, dv=plan.dv))

            def backward_level0():
                ops.stem_out_bwd(plan.dv, x_last, plan.x, Lp["up_w"], Lp["up_b"], gate, lv.factor, dh0,
                                 dw_up, db_up, dgate, gscale=plan.gscale, append=plan.append,
                                 noise=plan.noise, alpha=plan.alpha, beta=plan.beta,
                                 w_adapt=Lp.get("adapt_w"), dw_adapt=dwa, db_adapt=dba)
                d = dh0
                for b_ in reversed(items_up_bwd):
                    d = b_(d)
                if inner is not None:
                    d = inner(d)
                for b_ in reversed(items_down_bwd):
                    d = b_(d)
                ops.stem_in_bwd(d, plan.x, dw_down, db_down, lv.factor, append=plan.append,
                                noise=plan.noise, alpha=plan.alpha, beta=plan.beta)
            return None, None, backward_level0

        # levels >= 1: up conv writes y (pre-gate), skip_gate merges with the level's input
        f, Co = lv.factor, lv.out_ch
        y_up, out, ost = act(B, T_in, Co), act(B, T_in, Co), new_stats()
        dys, dx_last, d_xin = act(B, T_in, Co), act(B, Tl, C), act(B, T_in, lv.in_ch)
        if f > 1:
            plan.fwd.append(lambda: ops.conv_gemm(x_last, Lp["up_w"], y_up.view(B, Tl, f * Co), c_in=C,
                                                  n_valid=Co, up_factor=f, bias=Lp["up_b"]))
            def make_wd_up():
                w = lv.up.weight.detach().float()
                w0, w1, w2 = w[:, :, 0], w[:, :, 1], w[:, :, 2]
                wd = torch.zeros(3, C, f * Co, device=dev)      # taps read dys rows q-1, q, q+1
                wd[2, :, 0:Co] = w0.t()                          # forward off -1 (phase 0, slot 0)
                wd[1, :, 0:Co] = (w1 + w2).t()
                wd[1, :, (f - 1) * Co:f * Co] = (w0 + w1).t()
                wd[0, :, (f - 1) * Co:f * Co] = w2.t()          # forward off +1 (last phase, slot 1)
                for p_ in range(1, f - 1):
                    wd[1, :, p_ * Co:(p_ + 1) * Co] = (w0 + w1 + w2).t()
                return ops._pad_rows(wd.permute(1, 0, 2).reshape(C, 3 * f * Co), ops.round_up(C, 16))
            wd_up = packed_dgrad(make_wd_up)
            gwc = gbuf((f, 2, Co, C))

            def up_wgrad():
                for p_ in range(f):
                    slots = [(0, -1), (1, 0)] if p_ == 0 else ([(0, 0), (1, 1)] if p_ == f - 1 else [(0, 0)])
                    for s_, off in slots:
                        ops.wgrad(dys.view(B, Tl, f * Co), x_last, gwc[p_, s_], n=Co, k=C, off=off,
                                  g_col0=p_ * Co)

            def up_final():
                g0 = gwc[0, 0] + gwc[f - 1, 0]
                g1 = gwc[0, 1] + gwc[f - 1, 0]
                g2 = gwc[0, 1] + gwc[f - 1, 1]
                for p_ in range(1, f - 1):
                    g0, g1, g2 = g0 + gwc[p_, 0], g1 + gwc[p_, 0], g2 + gwc[p_, 0]
                grads[id(lv.up.weight)] = torch.stack([g0, g1, g2], dim=-1)
            finals.append(up_final)

            def up_dgrad():
                ops.conv_gemm(dys.view(B, Tl, f * Co), wd_up, dx_last, c_in=f * Co, n_valid=C,
                              taps=(-1, 0, 1))
        else:
            plan.fwd.append(lambda: ops.conv_gemm(x_last, Lp["up_w"], y_up, c_in=C, n_valid=Co,
                                                  taps=(-1, 0, 1), bias=Lp["up_b"]))
            wd_up = packed_dgrad(lambda: ops.pack_conv_dgrad(lv.up.weight.detach()))
            gw3 = gbuf((3, Co, C))
            finals.append(lambda: grads.__setitem__(id(lv.up.weight), gw3.permute(1, 2, 0)))

            def up_wgrad():
                for tap, off in enumerate((-1, 0, 1)):
                    ops.wgrad(dys, x_last, gw3[tap], n=Co, k=C, off=off)

            def up_dgrad():
                ops.conv_gemm(dys, wd_up, dx_last, c_in=Co, n_valid=C, taps=(-1, 0, 1))
        plan.fwd.append(lambda: ops.skip_gate(y_up, x_in, gate, out, ost, G))

        def backward_level(d_out: Tensor) -> Tensor:
            ops.skip_gate_bwd(d_out, y_up, gate, dys, dgate)
            ops.colsum(dys, db_up)
            up_wgrad()
            up_dgrad()
            d = dx_last
            for b_ in reversed(items_up_bwd):
                d = b_(d)
            if inner is not None:
                d = inner(d)
            for b_ in reversed(items_down_bwd):
                d = b_(d)
            ops.colsum(d, db_down)
            kdim = lv.factor * lv.in_ch
            ops.wgrad(d, x_in.view(B, Tl, kdim), gw_down, n=C, k=kdim, off=0)
            # gradient w.r.t. the level input = dgrad(down conv) + the skip path (d_out)
            ops.conv_gemm(d, wd_down, d_xin.view(B, Tl, kdim), c_in=C, n_valid=kdim,
                          residual=d_out.view(B, Tl, kdim))
            return d_xin

        return out, ost, backward_level

    _, _, backward0 = level(0, None, T)
    plan.backward0 = backward0
    plan.flat = flat
    plan.refreshers, plan.version = refreshers, net._version()
    plan.grads, plan.finals = grads, finals
    plan.ss_all, plan.dss_all = ss_all, dss_all
    # conditioning projection backward
    n_pad_rows = P["cond_w"].shape[0]
    plan.dw_all = _zeros((n_tot, Fm), dev)
    plan.dbias_all = _zeros((n_tot,), dev)
    plan.n_tot = n_tot

    def cond_backward():
        plan.dcond.zero_()
        ops.cond_bwd(dss_all, plan.cond_bf.view(B, Fm).float(), P["cond_w"], plan.dw_all,
                     plan.dbias_all, plan.dcond, n_tot)
    plan.cond_backward = cond_backward
    return plan


def _run(plan: _TrainPlan, which: str, use_graph: bool) -> None:
    """Eager on the first call, captured on the second, replayed afterwards."""
    prog = (lambda: [f() for f in plan.fwd]) if which == "f" else \
        (lambda: (plan.flat.zero_(), plan.backward0(), plan.cond_backward()))
    runs = plan.runs_f if which == "f" else plan.runs_b
    graph = plan.graph_f if which == "f" else plan.graph_b
    if not use_graph or runs == 0:
        prog()
    elif graph is None:
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            prog()
        if which == "f":
            plan.graph_f = g
        else:
            plan.graph_b = g
        g.replay()
    else:
        graph.replay()
    if which == "f":
        plan.runs_f += 1
    else:
        plan.runs_b += 1


class _UNetVLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net: B200UNet, x, noise, sigmas, append, cond, *params):
        B, _, T = x.shape
        key = ("train", B, T)
        net.packed()                         # in-place refresh of the forward packs
        plan = net._plans.get(key)
        if plan is None:
            ops.device_check()
            plan = net._plans[key] = build_train_plan(net, B, T)
        elif plan.version != net._version():
            with torch.no_grad():
                for r in plan.refreshers:
                    r()
            plan.version = net._version()
        plan.x.copy_(x)
        plan.noise.copy_(noise)
        if net.append_channels:
            assert append is not None, "append_channels is required (AppendChannelsPlugin)"
            plan.append.copy_(append)
        angle = sigmas.float() * pi / 2
        plan.alpha.copy_(torch.cos(angle))
        plan.beta.copy_(torch.sin(angle))
        plan.cond.copy_(cond)
        _run(plan, "f", net.use_cuda_graph)
        ctx.plan, ctx.net, ctx.n_params = plan, net, len(params)
        ctx.params = params
        return (plan.loss_sum / plan.dv.numel()).float().reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        plan, net = ctx.plan, ctx.net
        plan.gscale.copy_(grad_out.reshape(1))
        _run(plan, "b", net.use_cuda_graph)
        for fin in plan.finals:
            fin()
        out = []
        for p in ctx.params:
            g = plan.grads.get(id(p))
            if isinstance(g, tuple):
                kind, off, n = g
                g = plan.dw_all[off:off + n] if kind == "cond_w" else plan.dbias_all[off:off + n]
            out.append(None if g is None else g.reshape(p.shape).to(p.dtype).contiguous())
        return (None, None, None, None, None, plan.dcond.clone(), *out)


def fused_v_loss(net: B200UNet, x: Tensor, noise: Tensor, sigmas: Tensor, *,
                 append_channels: Optional[Tensor] = None, features: Optional[Tensor] = None,
                 **unsupported) -> Tensor:
    """mse(net(alpha*x + beta*noise, sigma), alpha*noise - beta*x)  (reference diffusion.py:90-95)."""
    assert x.is_cuda, "the B200 path runs on a CUDA device only (no CPU fallback)"
    for k, v in unsupported.items():
        assert v is None or k in ("embedding_scale", "embedding_mask_proba", "channels"), \
            f"training with `{k}` is outside the built hot path"
    # time features in PyTorch (autograd gives their gradients; three [B,1024] linears)
    if net.time is not None:
        t = net.time
        s = sigmas.float().unsqueeze(-1)
        fr = s * t.weights * 2 * pi
        emb = t.to_out(torch.cat([s, fr.sin(), fr.cos()], dim=-1))
        f = F.gelu(t.mlp(F.gelu(t.mlp(F.gelu(emb)))))
        if features is not None:
            f = f + features
    else:
        assert features is not None, "use_time_conditioning=False needs features="
        f = features
    cond = F.silu(f)
    time_ids = {id(p) for p in (net.time.parameters() if net.time is not None else [])}
    params = [p for p in net.parameters() if id(p) not in time_ids]
    return _UNetVLoss.apply(net, x.float(), noise.float(), sigmas, append_channels, cond, *params)
