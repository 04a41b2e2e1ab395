"""Throughput of the pose-window preprocessing kernels on B200, with the CPU oracle
(the reference's numpy/torch chain restated) timed on a bounded sample of the same frames.
   python tools/preprocess_bench.py [--frames 400000] [--cpu-frames 6000]  -> one JSON line"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import scrubvae_b200 as sv
from oracle import scvae_oracle as orc
from test_preprocess import _synth_frames

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=400000)
ap.add_argument("--cpu-frames", type=int, default=6000)
a = ap.parse_args()
pose, ids = _synth_frames(a.frames, seed=11)
torch.cuda.set_device(0)
pose_d = torch.as_tensor(pose).cuda()
kw = dict(window=51, stride=2, speed_threshold=2.25, direction_process="midfwd")
out = sv.data.preprocess_windows(pose_d, ids, orc.KINEMATIC_TREE, orc.OFFSET, **kw)   # warm-up
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    out = sv.data.preprocess_windows(pose_d, ids, orc.KINEMATIC_TREE, orc.OFFSET, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n_w, n_k = out["window_inds"].shape[0], out["x6d"].shape[0]
t0 = time.perf_counter()
ref = orc.preprocess(pose[:a.cpu_frames], ids[:a.cpu_frames], **kw)
cpu_s = time.perf_counter() - t0
cpu_w = ref["window_inds"].shape[0]
# algorithmic bytes per window: read 51*18*3 doubles, write x6d + root + offsets + target (float32)
bytes_per_w = 51 * 18 * 3 * 8 + 51 * (18 * 6 + 3 + 18 * 3 + 18 * 3) * 4
print(json.dumps({"metric": "pose windows/sec preprocessed", "value": n_w / (ms * 1e-3), "unit": "windows/s", "frames": a.frames,
                  "windows": n_w, "kept": n_k, "ms": ms, "includes": "window starts on the host, H2D of the start indices, 3 kernels, kept-window compaction",
                  "alg_GBps": n_w * bytes_per_w / (ms * 1e-3) / 1e9,
                  "cpu_baseline": {"value": cpu_w / cpu_s, "unit": "windows/s", "kind": "port", "sample": f"{a.cpu_frames} frames -> {cpu_w} windows, oracle.preprocess",
                                   "cores": torch.get_num_threads()}}))
