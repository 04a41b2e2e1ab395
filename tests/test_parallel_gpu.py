"""Data parallelism on real GPUs (needs >= 2; `gpurun --gpus 2`): NCCL all-reduce captured inside the step's CUDA
graph, the three-stream joins and the 1/world gradient scale — replicas must stay BIT-identical over several replayed
steps, and the reduced gradient must be the mean of the per-shard oracle gradients."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CH, Z, BL = [16, 32, 64, 128, 256], 16, 24
SCALE = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, "heading_gr": 1.0}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import TrainStep
    from scrubvae_b200 import parallel
    from oracle import scvae_oracle as orc
    from test_engine_cpu import build_model, _rel
    torch.manual_seed(10 + rank)  # replicas start DIFFERENT: setup() must make them equal
    m, dcfg = build_model(CH, Z, ["heading"], ["heading"], device=dev)
    m.train()
    opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-3, "lr_schedule": None})
    comm = parallel.setup(m, opt)
    sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    data = {k: v.to(dev) for k, v in orc.synth_batch(BL, seed=100 + rank).items()}
    eps = orc.synth_eps(BL, Z, seed=200 + rank)
    m._noise = eps.to(dev)
    step = TrainStep(m, opt, SCALE, BL, use_graph=True, comm=comm, keep_grads=True)
    step.run(data)  # eager
    torch.cuda.synchronize()
    g1 = {n: (g / world).cpu().clone() for n, g in step.named_grads().items()}
    for _ in range(3):  # replayed from the graph (NCCL captured)
        step.run()
    torch.cuda.synchronize()
    flat = m.engine.flat.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    bufs = torch.cat([b.reshape(-1).float() for b in m.buffers()])
    gb = [torch.zeros_like(bufs) for _ in range(world)]
    dist.all_gather(gb, bufs)
    # oracle: every shard on this rank, gradients averaged
    cfg = orc.Cfg(ch=CH, z_dim=Z)
    gsum = None
    for r in range(world):
        _, g, _, _, _ = orc.train_step(sd0, orc.synth_batch(BL, seed=100 + r), cfg, SCALE, orc.synth_eps(BL, Z, seed=200 + r))
        gsum = g if gsum is None else {k: gsum[k] + g[k] for k in g}
    gavg = {k: v / world for k, v in gsum.items()}
    gn = sum(float((v.double() ** 2).sum()) for v in gavg.values()) ** 0.5
    errs = {n: (_rel(g1[n], gavg[n]), (g1[n].double() - gavg[n].double()).norm().item() / gn) for n in gavg}
    finite = bool(torch.isfinite(flat).all())
    q.put((rank, same, errs, finite))
    del step  # parallel.shutdown finds the step through the comm hook's registry and drops its captured graph
    dist.barrier()
    parallel.shutdown(m)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_graph_step():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():  # never leave a worker behind: pytest would wait for it at exit
            p.kill()
    assert codes == [0] * world, f"teardown did not finish cleanly: exit codes {codes}"
    for rank, same, errs, finite in res:
        assert finite
        assert same, "replicas diverged after 4 data-parallel steps"
        for n, (rel, glob) in errs.items():  # fp32 path: reduced gradient = mean of the shard gradients
            assert rel < 1e-3 or glob < 5e-6, (rank, n, rel, glob)


def _worker_sync(rank, world, port, q):
    """dp_bn "sync" on real GPUs: statistics all-reduces captured in the step graph; the replicas together must reproduce
    ONE device on the concatenated batch (oracle on the global batch), and stay bit-identical."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import TrainStep
    from scrubvae_b200 import parallel
    from oracle import scvae_oracle as orc
    from test_engine_cpu import build_model, _rel
    scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0}
    torch.manual_seed(20 + rank)
    m, dcfg = build_model(CH, Z, ["heading"], [], device=dev)
    m.train()
    opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-3, "lr_schedule": None})
    comm = parallel.setup(m, opt, bn_sync=True)
    sd0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    shards = [orc.synth_batch(BL, seed=100 + r) for r in range(world)]
    epss = [orc.synth_eps(BL, Z, seed=200 + r) for r in range(world)]
    m._noise = epss[rank].to(dev)
    step = TrainStep(m, opt, scale, BL, use_graph=True, comm=comm, keep_grads=True)
    step.run({k: v.to(dev) for k, v in shards[rank].items()})  # eager
    torch.cuda.synchronize()
    g1 = {n: (g / world).cpu().clone() for n, g in step.named_grads().items()}
    for _ in range(2):  # replayed from the graph (gradient AND statistics all-reduces captured)
        step.run()
    torch.cuda.synchronize()
    flat = m.engine.flat.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    cfg = orc.Cfg(ch=CH, z_dim=Z, grad_reversal=())
    glob = {k: torch.cat([s[k] for s in shards], 0) for k in shards[0]}
    _, g, _, _, _ = orc.train_step(sd0, glob, cfg, scale, torch.cat(epss, 0))
    gn = sum(float((v.double() ** 2).sum()) for v in g.values()) ** 0.5
    errs = {n: (_rel(g1[n], g[n]), (g1[n].double() - g[n].double()).norm().item() / gn) for n in g}
    q.put((rank, same, errs, bool(torch.isfinite(flat).all())))
    del step
    dist.barrier()
    parallel.shutdown(m)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sync_batchnorm_on_gpu():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sync, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert codes == [0] * world, f"teardown did not finish cleanly: exit codes {codes}"
    for rank, same, errs, finite in res:
        assert finite and same, "replicas diverged"
        for n, (rel, glob) in errs.items():  # fp32 path: = the single-device gradient on the global batch
            assert rel < 1e-3 or glob < 5e-6, (rank, n, rel, glob)
