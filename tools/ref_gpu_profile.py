"""Where the UNMODIFIED reference's training step spends its time on a B200 (PyTorch eager): torch.profiler over one
step of reference train_test_epoch at the benchmarked batch, plus step times under a few precision / cudnn settings.
   python tools/ref_gpu_profile.py [--batch 2048]"""
import argparse, contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_runner as rr, refimport
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=2048); a = ap.parse_args()
refimport.import_reference()
from scrubvae.train import trainer
dev = torch.device("cuda", 0)
scale = {"prior": 1e-4, "jpe": 1.0, "root": 1.0, "heading_gr": 1.0}
m, dc = rr.build_model(dev)
data = {k: v.to(dev) for k, v in rr.synth_batch(a.batch).items()}
cfg = {"loss": scale, "disentangle": dc}
with contextlib.redirect_stdout(io.StringIO()):
    opt, _ = trainer.get_optimizer_and_lr_scheduler(m, {"optimizer": "adamw", "lr": 1e-4, "lr_schedule": None})

def epoch(n):
    with contextlib.redirect_stdout(io.StringIO()):
        trainer.train_test_epoch(cfg, m, [data] * n, dev, 1, optimizer=opt, scheduler=None, mode="train")

def timeit(tag, n=3):
    epoch(2); torch.cuda.synchronize(); t0 = time.perf_counter(); epoch(n); torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"{tag:60s} {dt*1e3:9.1f} ms/step  {a.batch/dt:10.0f} windows/s", flush=True)

print("anomaly mode:", torch.is_anomaly_enabled())
with rr.precision("tf32"):
    timeit("reference settings (medium matmul, cudnn tf32, cudnn.benchmark)")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        epoch(1); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=60))
torch.backends.cudnn.benchmark = False
timeit("torch defaults (highest matmul, cudnn tf32, benchmark off)")
with rr.precision("fp32"):
    timeit("strict fp32")
