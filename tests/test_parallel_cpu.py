"""Data-parallel step on 2 ranks (gloo, CPU): the engine's bucketed gradient all-reduce (parallel.py) over the
torch-on-CPU emulation of the C ABI, against "the oracle run on each shard, gradients averaged" (SURVEY.md §8e:
per-replica BatchNorm statistics, local-B loss normalisation), through one AdamW update."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CH, Z, BL = [8, 16, 32, 64, 128], 8, 5
SCALE = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, "heading_gr": 1.0}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import Engine, TrainStep
    from scrubvae_b200 import parallel
    from oracle import scvae_oracle as orc
    from emu_ops import EmuOps
    from test_engine_cpu import build_model, _rel
    torch.manual_seed(10 + rank)  # replicas start DIFFERENT: broadcast_parameters must make them equal
    m, dcfg = build_model(CH, Z, ["heading"], ["heading"])
    m._engine = Engine(m, ops=EmuOps())
    m.train()
    opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-3, "lr_schedule": None})
    comm = parallel.setup(m, opt)  # broadcast + bucketed all-reduce hook + 1/world gradient scale
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    data = orc.synth_batch(BL, seed=100 + rank)
    eps = orc.synth_eps(BL, Z, seed=200 + rank)
    m._noise = eps
    step = TrainStep(m, opt, SCALE, BL, use_graph=False, comm=comm, keep_grads=True)
    step.run(data)
    losses = {k: v.item() for k, v in step.losses().items()}
    # oracle: every shard on this rank (so that each rank can check alone), gradients averaged, one AdamW step
    cfg = orc.Cfg(ch=CH, z_dim=Z)
    gsum, lref = None, None
    for r in range(world):
        l, g, _, _, _ = orc.train_step(sd0, orc.synth_batch(BL, seed=100 + r), cfg, SCALE, orc.synth_eps(BL, Z, seed=200 + r))
        gsum = g if gsum is None else {k: gsum[k] + g[k] for k in g}
        if r == rank:
            lref = {k: v.item() for k, v in l.items()}
    gavg = {k: v / world for k, v in gsum.items()}
    errs = {}
    gn = sum(float((v.double() ** 2).sum()) for v in gavg.values()) ** 0.5
    for n, gv in step.named_grads().items():
        e = ((gv / world).double() - gavg[n].double()).norm().item()
        errs[n] = (_rel(gv / world, gavg[n]), e / gn)
    new = {n: p.detach().clone() for n, p in m.named_parameters()}
    ref_new = {}
    for n in gavg:
        p1, _, _ = orc.adam_update(sd0[n].clone(), gavg[n], torch.zeros_like(gavg[n]), torch.zeros_like(gavg[n]), 1, 1e-3,
                                   kind="adamw")
        ref_new[n] = p1
    perr = {n: _rel(new[n], ref_new[n]) for n in new}
    flat = torch.cat([p.reshape(-1) for p in new.values()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    q.put((rank, losses, lref, errs, perr, same, comm.bytes_per_step))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_matches_sharded_oracle():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, losses, lref, errs, perr, same, nbytes in res:
        for k, v in lref.items():  # local losses = oracle on the local shard
            assert abs(losses[k] - v) <= 2e-5 * abs(v) + 1e-6, (rank, k, losses[k], v)
        for n, (rel, glob) in errs.items():  # reduced gradient = mean of the shard gradients
            assert rel < 3e-4 or glob < 2e-6, (rank, n, rel, glob)
        # one AdamW step with the averaged gradient.  The first Adam step moves every element by lr * sign(g): the few
        # elements whose gradient is rounding noise flip sign (2e-3 of |p| each), hence 5e-3 and not 1e-5
        from test_engine_cpu import ZERO_GRAD_BIAS  # biases in front of a BatchNorm: their true gradient is 0
        for n, e in perr.items():
            assert e < 5e-3 or ZERO_GRAD_BIAS.search(n), (rank, n, e)
        assert same, "replicas diverged after the update"
        assert nbytes > 0


SCALE_NOGR = {"prior": 1e-4, "jpe": 1.0, "root": 1.0}


def _worker_sync(rank, world, port, q):
    """dp_bn "sync": the two replicas together must reproduce ONE device on the concatenated batch."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import Engine, TrainStep
    from scrubvae_b200 import parallel
    from oracle import scvae_oracle as orc
    from emu_ops import EmuOps
    from test_engine_cpu import build_model, _rel
    torch.manual_seed(10 + rank)
    m, dcfg = build_model(CH, Z, ["heading"], [])
    m._engine = Engine(m, ops=EmuOps())
    m.train()
    opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-3, "lr_schedule": None})
    comm = parallel.setup(m, opt, bn_sync=True)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    shards = [orc.synth_batch(BL, seed=100 + r) for r in range(world)]
    epss = [orc.synth_eps(BL, Z, seed=200 + r) for r in range(world)]
    m._noise = epss[rank]
    step = TrainStep(m, opt, SCALE_NOGR, BL, use_graph=False, comm=comm, keep_grads=True)
    step.run(shards[rank])
    losses = {k: v.item() for k, v in step.losses().items()}
    lsum = torch.tensor([losses[k] for k in sorted(losses)], dtype=torch.double)
    dist.all_reduce(lsum)
    lmean = dict(zip(sorted(losses), (lsum / world).tolist()))
    # ONE device on the global batch
    cfg = orc.Cfg(ch=CH, z_dim=Z, grad_reversal=())
    glob = {k: torch.cat([s[k] for s in shards], 0) for k in shards[0]}
    l, g, _, sd1, _ = orc.train_step(sd0, glob, cfg, SCALE_NOGR, torch.cat(epss, 0))
    lref = {k: v.item() for k, v in l.items()}
    gn = sum(float((v.double() ** 2).sum()) for v in g.values()) ** 0.5
    errs = {n: (_rel(gv / world, g[n]), ((gv / world).double() - g[n].double()).norm().item() / gn)
            for n, gv in step.named_grads().items() if n in g}
    bufs = {k: v for k, v in m.state_dict().items() if k.endswith("running_mean") or k.endswith("running_var")}
    berr = {k: _rel(v.float(), sd1[k].float()) for k, v in bufs.items() if k in sd1}
    q.put((rank, lmean, lref, errs, berr))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sync_batchnorm_matches_one_device_on_the_global_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sync, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, lmean, lref, errs, berr in res:
        for k, v in lref.items():  # mean of the local losses = the global-batch loss
            if k in lmean:
                assert abs(lmean[k] - v) <= 3e-5 * abs(v) + 1e-6, (rank, k, lmean[k], v)
        assert len(errs) > 50
        for n, (rel, glob) in errs.items():  # reduced gradient = the single-device gradient on the global batch
            assert rel < 3e-4 or glob < 2e-6, (rank, n, rel, glob)
        for k, e in berr.items():  # running statistics of the global batch on every rank
            assert e < 1e-5, (rank, k, e)
