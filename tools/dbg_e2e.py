import os, sys, time, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, bench, scrubvae_b200 as sv
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
m, dcfg = bench.build_model(dev, "tf32"); m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
host = bench.synth_host_batch(2048, seed=0)
config = {"loss": dict(bench.LOSS_SCALE), "disentangle": dcfg, "train": {}}
loss_host = torch.zeros(1, pin_memory=True)
CB = [None]
def read_loss(i, vec):
    loss_host.copy_(vec[-1:], non_blocking=True)
def run(n):
    with contextlib.redirect_stdout(io.StringIO()):
        sv.train.train_test_epoch(config, m, [host] * n, dev, 1, optimizer=opt, scheduler=None, mode="train", step_callback=CB[0])
run(3); torch.cuda.synchronize()
for n in (5, 20):
    t0 = time.perf_counter(); run(n); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(n, "steps:", (t1 - t0) / n * 1e3, "ms/step")
CB[0] = read_loss
for n in (5, 20):
    t0 = time.perf_counter(); run(n); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(n, "steps with callback:", (t1 - t0) / n * 1e3, "ms/step")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); run(20); e1.record(); torch.cuda.synchronize()
print("events:", e0.elapsed_time(e1) / 20, "ms/step")
from scrubvae_b200.engine import TrainStep
step2 = TrainStep(m, opt, bench.LOSS_SCALE, 2048, use_graph=True)
dd = {k: v.to(dev) for k, v in host.items()}
step2.run(dd); step2.run(); step2.run(); torch.cuda.synchronize()
torch.cuda.synchronize(); e0.record(); run(20); e1.record(); torch.cuda.synchronize()
print("events after a second TrainStep exists:", e0.elapsed_time(e1) / 20, "ms/step")
st = list(m._train_steps.values())[0]
print("graph", st.graph is not None, "n steps objs", len(m._train_steps))
# H2D alone
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
torch.cuda.synchronize(); print("H2D ms", (time.perf_counter() - t0) / 10 * 1e3, "pinned", [v.is_pinned() for v in host.values()])
# step alone
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): st.run()
torch.cuda.synchronize(); print("graph step ms", (time.perf_counter() - t0) / 10 * 1e3)
d = {k: v.to(dev) for k, v in host.items()}
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): st.run(d)
torch.cuda.synchronize(); print("graph step + D2D ms", (time.perf_counter() - t0) / 10 * 1e3)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); run(10); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
