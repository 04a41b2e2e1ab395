"""Launch one eager training step (for ncu on the elementwise kernels): python tools/ew_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, bench, scrubvae_b200 as sv
from scrubvae_b200.engine import TrainStep
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
m, dcfg = bench.build_model(dev, "tf32"); m.train()
opt, _ = sv.train.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})
data = {k: v.to(dev) for k, v in bench.synth_host_batch(2048, seed=0).items()}
step = TrainStep(m, opt, bench.LOSS_SCALE, 2048, use_graph=False, resident=True)
step.run(data); step.run(); torch.cuda.synchronize()
torch.cuda.nvtx.range_push("once"); step.run(); torch.cuda.synchronize(); torch.cuda.nvtx.range_pop(); print("ok")
