"""N>1 host logic on CPU: two gloo ranks exercise bench.py's sharding seeds and the
max-over-ranks reduction (the only cross-rank exchange of the inference path)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import bench
    from otpose_b200.utils import synthetic as syn
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert bench.dist_env() == (rank, rank, world)
    slow = bench.dist_max(10.0 + 5.0 * rank)           # every rank must see the slowest rank's time
    clips = syn.synth_rough_heatmaps(1, 17, 8, 6, seed=bench.shard_seed(1234, rank))
    gathered = [torch.zeros_like(clips) for _ in range(world)]
    dist.all_gather(gathered, clips)
    q.put((rank, slow, bool(torch.equal(gathered[0], gathered[1]))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_max_time():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [15.0, 15.0]
    assert not res[0][2], "ranks must draw different shards of synthetic clips"


def test_algorithmic_work_matches_survey_flops():
    sys.path.insert(0, ROOT)
    import bench
    w = bench.algorithmic_work(1)
    enc = w["block_front"][1] + w["block_apply"][1] + w["block_back"][1]
    # SURVEY 8d counts 2 x 22.551 + 0.348 GFLOP for the three encoders per clip.  The
    # kernels' algorithmic figure leaves out the depthwise convs (0.23 G, CUDA-core
    # work) and att@v (1.75 G: folded into W_eff, never executed) -> 43.46 G.
    assert abs(enc - 43.46e9) / 43.46e9 < 0.002
    assert 0.94 < enc / (2 * 22.551e9 + 0.348e9) < 1.0
    assert w["mdcn_fwd"][0] == "hbm" and abs(w["mdcn_fwd"][1] - 5 * 13.63e6) / 68e6 < 0.01


def _reduce_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from otpose_b200.train import BucketedGradReducer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                    # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(40, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                torch.nn.Linear(64, 3))
    unused = torch.nn.Parameter(torch.ones(7))             # a parameter that never receives a gradient
    params = list(model.parameters()) + [unused]
    red = BucketedGradReducer(params, bucket_bytes=4096)
    x = torch.randn(5, 40, generator=torch.Generator().manual_seed(100 + rank))
    outs = []
    for step in range(2):                                   # two steps: the buckets re-arm
        red.zero_grad()
        red.enabled = False                                 # gradient accumulation: first micro-batch is local
        model(x[:2]).pow(2).sum().backward()
        red.enabled = True
        model(x[2:]).pow(2).sum().backward()
        red.finish()
        norm = red.clip_grad_norm_(1e9)
        outs.append((red.flat.clone(), float(norm)))
    # reference: the average over the ranks of the single-process gradients
    ref = torch.zeros_like(red.flat)
    for r in range(world):
        for p in params:
            p.grad = None
        xr = torch.randn(5, 40, generator=torch.Generator().manual_seed(100 + r))
        model(xr).pow(2).sum().backward()
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in red.params])
        ref += flat / world
    ok = all(torch.allclose(o, ref, rtol=1e-5, atol=1e-6) for o, _ in outs)
    q.put((rank, ok, red.num_buckets, outs[0][1], float(ref.norm())))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_all_reduce_two_ranks():
    """configs[3] host logic on 2 gloo ranks: bucketed, hook-launched all-reduce == rank-average of the
    gradients, with gradient accumulation, an unused parameter, several buckets and re-arming."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, nb, norm, refnorm in res:
        assert ok, rank
        assert nb >= 2
        assert abs(norm - refnorm) < 1e-4 * refnorm
