"""Host-side logic of the multi-GPU path on CPU: two gloo ranks shard a batch, gather it
back, and average gradients; bench.py's rank handling for the reference arm."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from audio_diffusion_pytorch_b200 import parallel
    full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    local = parallel.shard_batch(full)
    assert local.shape[0] == (4 if rank == 0 else 3)
    back = parallel.gather_batch(local * 2, 7)
    assert torch.equal(back, full * 2)
    # gradient averaging: rank r holds grads filled with (r + 1)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(11))]
    for p in params:
        p.grad = torch.full_like(p, float(rank + 1))
    parallel.allreduce_gradients(params, bucket_mb=1e-5)     # forces several buckets
    for p in params:
        assert torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2))

    class Toy(torch.nn.Module):
        def sample(self, noise, num_steps, scale=None):
            return noise * num_steps + (0 if scale is None else scale.view(-1, 1))
    noise = torch.arange(10, dtype=torch.float32).reshape(5, 2)
    out = parallel.sample_sharded(Toy(), noise, 3, scale=torch.arange(5, dtype=torch.float32))
    assert torch.equal(out, noise * 3 + torch.arange(5, dtype=torch.float32).view(-1, 1))
    # OverlappedDataParallel host logic: the parameters whose gradients come from PyTorch autograd
    # are averaged by finish_gradient_sync; the conditioning-projection gradient is rebuilt from
    # the gathered rank-B factors instead of all-reducing the full matrix
    lin = torch.nn.Linear(3, 2)
    wrapper = parallel.OverlappedDataParallel.__new__(parallel.OverlappedDataParallel)
    torch.nn.Module.__init__(wrapper)
    wrapper.group, wrapper.outside = None, list(lin.parameters())
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(10 * (rank + 1)))
    wrapper.finish_gradient_sync()
    for p in lin.parameters():
        assert torch.allclose(p.grad, torch.full_like(p, 10 * (1 + world) / 2))
    g = torch.Generator().manual_seed(rank)
    dss, cond = torch.randn(3, 16, generator=g), torch.randn(3, 8, generator=g)
    local = dss.t() @ cond                                   # this rank's dW
    want = local.clone()
    dist.all_reduce(want)
    gathered_d = [torch.empty_like(dss) for _ in range(world)]
    gathered_c = [torch.empty_like(cond) for _ in range(world)]
    dist.all_gather(gathered_d, dss)
    dist.all_gather(gathered_c, cond)
    assert torch.allclose(torch.cat(gathered_d).t() @ torch.cat(gathered_c), want, atol=1e-5)
    t = torch.full((6,), float(rank + 1))
    parallel.average_async(t).wait()
    assert torch.allclose(t, torch.full((6,), (1 + world) / 2))
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_allreduce(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from audio_diffusion_pytorch_b200.parallel import shard_bounds
    for n in (1, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_bench_reference_arm_rank_handling():
    """Under torchrun only rank 0 runs/prints the reference arm; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bucket_schedule_covers_the_arena_once():
    """Gradient-arena ranges arrive in backward order (up parts from the end of the arena towards
    the middle, then down parts towards the start): every element is communicated exactly once,
    in maximal contiguous ranges, in buckets of at least the requested size."""
    sys.path.insert(0, ROOT)
    from audio_diffusion_pytorch_b200.parallel import bucket_schedule, merge_intervals
    # arena layout [dss | L0 down | L1 down | L2 (innermost) | L1 up | L0 up]
    marks = [(900, 1000), (700, 900), (400, 700), (250, 400), (100, 250)]
    sched = bucket_schedule(marks, bucket_elems=280)
    sent = [iv for i in sorted(sched) for iv in sched[i]]
    assert merge_intervals(sent) == [(100, 1000)]
    assert sum(b - a for a, b in sent) == 900                      # no overlap
    assert all(sum(b - a for a, b in sched[i]) >= 280 for i in sorted(sched)[:-1])
    assert max(sched) == len(marks) - 1                            # the last mark always flushes
    assert merge_intervals([(5, 7), (1, 3), (3, 5), (9, 9)]) == [(1, 7)]
