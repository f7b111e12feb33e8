#!/usr/bin/env python
"""bench.py -- audio-seconds/sec of the 50-step VSampler on the README U-Net (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete `DiffusionModel.sample(noise[8,2,2^18], num_steps=50)` on each
rank (BASELINE configs[1]; weak scaling: 8 clips per GPU, independent sampling, no
collective on the data path).  Prints ONE JSON line (rank 0).

  value     clips*5.4613 s / wall, inputs resident in HBM, device-timed, max over ranks
  e2e       same through the public API from pinned HOST noise to a HOST result
  roofline  the dominant kernel of the step (largest share of device time), timed live with
            CUDA events in an instrumented eager pass of the same plan
  cpu_baseline  the reference's eager-PyTorch CPU path (oracle port) on this box's host cores,
            bounded sample, extrapolated (stated in `sample`)

`--impl reference` times only that CPU path (rank 0), same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 48000
LENGTH = 2 ** 18
CLIP_SECONDS = LENGTH / SAMPLE_RATE          # 5.4613 s, stereo counts once
BATCH = 8
NUM_STEPS = 50
README = dict(in_channels=2, channels=[8, 32, 64, 128, 256, 512, 512, 1024, 1024],
              factors=[1, 4, 4, 4, 2, 2, 2, 2, 2], items=[1, 2, 2, 2, 2, 2, 2, 4, 4],
              attentions=[0, 0, 0, 0, 0, 1, 1, 1, 1], attention_heads=8, attention_features=64)
METRIC = "audio-seconds/sec (2ch, 2^18 len) VSampler 50-step"
CONFIG = {"workload": "configs[1]: unconditional README UNetV0 9-stage, noise randn[8,2,2**18] per "
                      "GPU, VSampler num_steps=50", "batch_per_gpu": BATCH, "length": LENGTH,
          "num_steps": NUM_STEPS, "sample_rate": SAMPLE_RATE,
          "l2": "no explicit flush: one net evaluation streams ~0.9 GB (activations + 435 MB "
                "weights), far beyond the 126 MB L2"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p["hbm_gbs"], p["bf16_tflops"], p["bf16_tflops_sustained"], "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(steps: int, warmup: int, budget_s: float = 150.0, batch: int = 1):
    """The reference's own eager CPU path (oracle port over the a_unet shim), fp32, all host
    threads.  One timed unit = ONE VSampler step (reference diffusion.py:183-188: every step
    costs the same) on a BOUNDED sample of the workload: `batch` clip(s) of the largest
    power-of-two length <= 2**18 for which warmup+steps units fit `budget_s` (probed at 2**14).
    value = audio-seconds/sec of the full 50-step sampler extrapolated from that unit.  Shorter
    clips are slightly CHEAPER per audio second (attention is quadratic in length), so a
    bounded sample can only flatter the CPU arm."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_port as port
    torch.manual_seed(0)
    model = port.DiffusionModelPort(**README)

    def one_step(length):
        x = torch.randn(batch, 2, length)
        t0 = time.perf_counter()
        with torch.no_grad():
            model.sample(x, num_steps=1)
        return time.perf_counter() - t0

    # "all the host threads it can use": the count that actually runs this path fastest.  On a
    # box whose cgroup grants fewer CPUs than os.cpu_count() reports, cpu_count() threads
    # time-slice and every OpenMP barrier stalls (measured: 47 s per unit at 128 threads), so
    # the thread count is probed upwards from 8 on a short clip and the best one kept.
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:                                      # cgroup v2 CPU quota ("max 100000" = unlimited)
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            avail = max(1, min(avail, int(int(quota) / int(period))))
    except Exception:
        pass
    probe_len = 2 ** 14
    cores, t_probe = None, None
    for n in [c for c in (8, 16, 32, 64, 128, 256) if c < avail] + [avail]:
        torch.set_num_threads(n)
        one_step(2 ** 12)                     # thread pool / allocator warm at this width
        t = one_step(probe_len)
        if t_probe is not None and t > 0.9 * t_probe:
            if t < t_probe:
                cores, t_probe = n, t
            break                             # no longer scaling
        cores, t_probe = n, t
    torch.set_num_threads(cores)
    length = probe_len
    units = max(steps + warmup, 1)
    while length < LENGTH and units * t_probe * (2 * length / probe_len) <= budget_s:
        length *= 2
    times = []
    for i in range(warmup + steps):
        dt = one_step(length)
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    value = batch * CLIP_SECONDS * (length / LENGTH) / (NUM_STEPS * per_step)
    return value, per_step, cores, (f"{len(times)} x one VSampler step on [{batch},2,{length}] fp32 "
                                    f"({cores} threads; full clip is 2**18 = {LENGTH}), x{NUM_STEPS} "
                                    f"extrapolated to the 50-step sample")


def run_reference(args, rank: int):
    if rank != 0:
        return
    value, per_step, cores, sample = cpu_reference_run(args.steps, args.warmup)
    # ms_per_step: one 50-step sample of the bench batch (8 clips) at the measured rate
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": BATCH * CLIP_SECONDS / value * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(json.dumps(line))


UPSAMPLER = dict(upsample_factor=16, in_channels=2,
                 channels=[8, 32, 64, 128, 256, 512, 512, 1024, 1024],
                 factors=[1, 4, 4, 4, 2, 2, 2, 2, 2], items=[1, 2, 2, 2, 2, 2, 2, 4, 4])


def train_step_bench(adp, dev, world, dist, steps, warmup, batch=4):
    """BASELINE configs[3]: DiffusionUpsampler(upsample_factor=16) training step -- fused loss,
    hand-written backward, gradient all-reduce over NCCL (DDP) when world > 1, fused AdamW --
    `batch` clips of [2, 2**18] per GPU.  Returns ms per step (device-timed, max over ranks)."""
    model = adp.DiffusionUpsampler(net_t=adp.UNetV0, **UPSAMPLER).to(dev)
    step_model = model
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP
        step_model = DDP(model, device_ids=[dev.index], bucket_cap_mb=100, gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    audio = torch.randn(batch, 2, LENGTH, device=dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = step_model(audio)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(warmup, 3)):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return {"ms_per_step": float(ms.item()), "loss": float(loss), "batch_per_gpu": batch,
            "config": "configs[3]: DiffusionUpsampler upsample_factor=16, [B,2,2**18], fwd+bwd+"
                      "AdamW" + ("+NCCL grad all-reduce (DDP)" if world > 1 else ""),
            "audio_s_per_s": batch * world * CLIP_SECONDS / (float(ms.item()) * 1e-3)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the
# same shapes (tools/profile_r1c.sh -> profiles/r1_ncu_conv_*_v4.txt); None for other kernels
NCU_DRAM_BYTES = {"conv_gemm[k3 M=2048 K=1024 N=1024x1]": 14776064.0,
                  "conv_gemm[k3 M=32768 K=128 N=128x1]": 17011968.0}
NCU_DRAM_SOURCE = "profiles/r1_ncu_conv_L7_v4.txt, profiles/r1_ncu_conv_L3_v4.txt"


# stdout carries exactly ONE JSON line: libraries that print to fd 1 from C (NCCL's version
# banner) are sent to stderr for the duration of the run
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(text: str):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (text + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-kernel table")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import audio_diffusion_pytorch_b200 as adp
    torch.manual_seed(1234 + rank)
    model = adp.DiffusionModel(net_t=adp.UNetV0, **README).to(dev)
    net = model.net
    warmup = max(args.warmup, 3)
    host_noise = torch.randn(BATCH, 2, LENGTH).pin_memory()
    host_out = torch.empty(BATCH, 2, LENGTH).pin_memory()
    noise = host_noise.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, iters):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_device():
        model.sample(noise, num_steps=NUM_STEPS)

    def step_e2e():
        x = host_noise.to(dev, non_blocking=True)
        out = model.sample(x, num_steps=NUM_STEPS)
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(warmup):
        step_device()
    with ClockSampler(local) as clocks:
        total_ms = timed(step_device, args.steps)
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    ms_per_step = total_ms / args.steps
    clips = BATCH * world
    value = clips * CLIP_SECONDS / (ms_per_step * 1e-3)
    e2e_value = clips * CLIP_SECONDS / (e2e_ms / args.steps * 1e-3)

    # ---- roofline of the dominant kernel: instrumented eager pass of the same plan
    plan = next(p for k, p in net._plans.items() if len(k) > 4 and k[4] == "sample")
    cond_plan = next((p for k, p in net._plans.items() if k[0] == "cond"), None)
    table = net.profile_plan(plan, iters=5)
    hbm, tf_burst, tf_sust, which = peaks()
    top = max(table.values(), key=lambda r: r["ms_total"])
    step_ms = sum(r["ms_total"] for r in table.values())
    ai = top["flops"] / max(top["bytes"], 1)
    if ai >= tf_sust * 1e12 / (hbm * 1e9):
        achieved = top["flops"] / (top["ms_avg"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": tf_sust, "unit": "TFLOP/s",
                "frac": achieved / tf_sust}
    else:
        achieved = top["bytes"] / (top["ms_avg"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm}
    roof.update({"kernel": top["name"], "launches_per_net_eval": top["count"],
                 "share_of_step": top["ms_total"] / step_ms, "peak_source": which,
                 "traffic": NCU_DRAM_BYTES.get(top["name"]), "traffic_source": NCU_DRAM_SOURCE
                 if top["name"] in NCU_DRAM_BYTES else None,
                 "flops_per_launch": top["flops"], "bytes_per_launch": top["bytes"]})
    # whole-step roofline: sum over levels of max(flop time, byte time)
    t_bound = sum(max(r["flops"] / (tf_sust * 1e12), r["bytes"] / (hbm * 1e9)) * r["count"]
                  for r in table.values())
    roof["step_bound_ms"] = t_bound * 1e3
    roof["step_frac"] = t_bound * 1e3 / (ms_per_step / NUM_STEPS)
    if args.profile_ops and rank == 0:
        for r in sorted(table.values(), key=lambda r: -r["ms_total"]):
            print(f"# {r['name']:58s} x{r['count']:3d} avg {r['ms_avg'] * 1e3:8.1f} us  "
                  f"total {r['ms_total'] * 1e3:9.1f} us  {r['flops'] / max(r['ms_avg'], 1e-9) / 1e9:8.1f} TF/s "
                  f"{r['bytes'] / max(r['ms_avg'], 1e-9) / 1e6:8.1f} GB/s", file=sys.stderr)

    line = {"metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": dict(CONFIG, parallelism=f"dp{world} (independent sampling)"),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "audio-s/s",
                    "h2d_bytes_per_step": host_noise.numel() * 4,
                    "d2h_bytes_per_step": host_out.numel() * 4},
            # per step the net's kernels; per sample() call the 6 conditioning launches
            "gpu_launches": (plan.n_kernels * NUM_STEPS + (6 if cond_plan is not None else 0)) * args.steps,
            "ms_per_net_eval": ms_per_step / NUM_STEPS,
            "roofline": roof}
    if not args.no_train:
        del model, net, plan
        torch.cuda.empty_cache()
        line["train_step"] = train_step_bench(adp, dev, world, dist, steps=5, warmup=3)
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        v, per_step, cores, sample = cpu_reference_run(steps=1, warmup=1, budget_s=25.0)
        line["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                "sample": sample}
    if rank == 0:
        emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
