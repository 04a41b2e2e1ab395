"""Fused multi-tensor optimizers over the engine's flat parameter / gradient buffers: one
scv_optim_step launch per step, with clip_grad_norm_ folded in (reference train/trainer.py:60-70,
:160-165).  torch.optim.Optimizer subclass so LR schedulers and state_dict() keep working."""
from __future__ import annotations

import torch

KIND = {"adam": 0, "adamw": 1, "sgd": 2}


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, model, lr, kind="adamw", betas=(0.9, 0.999), eps=1e-8, weight_decay=None, momentum=0.2):
        if kind not in KIND:
            raise ValueError("No valid optimizer selected")
        if weight_decay is None:
            weight_decay = 0.01 if kind == "adamw" else 0.0
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, momentum=momentum)
        super().__init__(list(model.parameters()), defaults)
        self.model, self.kind = model, kind
        self._steps = 0
        self._eng = None
        self.grad_scale = 1.0  # 1/world_size under data parallelism (gradients are all-reduced sums)

    def _bind(self):
        eng = self.model.engine
        if self._eng is not eng:
            old = (self.m, self.v) if self._eng is not None else None
            self._eng = eng
            self.m = torch.zeros_like(eng.flat)
            self.v = torch.zeros_like(eng.flat)
            if old is not None and old[0].numel() == self.m.numel():
                self.m.copy_(old[0])
                self.v.copy_(old[1])
            self.hyper_host = torch.zeros(2, dtype=torch.double, pin_memory=eng.device.type == "cuda")
            self.hyper = torch.zeros(2, dtype=torch.double, device=eng.device)
            names = [n for n, _ in self.model.named_parameters()]
            for n, p in zip(names, eng.params):
                o = eng.poff[n]
                st = self.state[p]
                if self.kind == "sgd":
                    st["momentum_buffer"] = self.m[o:o + p.numel()].view(p.shape)
                else:
                    st["step"] = torch.tensor(float(self._steps))
                    st["exp_avg"] = self.m[o:o + p.numel()].view(p.shape)
                    st["exp_avg_sq"] = self.v[o:o + p.numel()].view(p.shape)
        return eng

    @torch.no_grad()
    def step(self, closure=None):
        eng = self._bind()
        grp = self.param_groups[0]
        self._steps += 1
        self.hyper_host[0] = float(grp["lr"])
        self.hyper_host[1] = float(self._steps)
        self.hyper.copy_(self.hyper_host, non_blocking=True)
        clip = getattr(eng, "clip", None)
        b1 = grp["momentum"] if self.kind == "sgd" else grp["betas"][0]
        eng.ops.optim_step(eng.flat, eng.gflat, self.m, self.v, eng.n_flat, clip[0] if clip else None,
                           clip[1] if clip else 0.0, self.grad_scale, float(grp["lr"]), b1, grp["betas"][1],
                           grp["eps"], grp["weight_decay"], self._steps, KIND[self.kind], hyper=self.hyper)
        eng.clip = None
        if self.kind != "sgd":
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(float(self._steps))
        return None

    def load_state_dict(self, state_dict):
        eng = self._bind()
        names = [n for n, _ in self.model.named_parameters()]
        sd_state = state_dict["state"]
        with torch.no_grad():
            for i, (n, p) in enumerate(zip(names, eng.params)):
                st = sd_state.get(i)
                if st is None:
                    continue
                o = eng.poff[n]
                if "exp_avg" in st:
                    self.m[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
                    self.v[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                    self._steps = int(float(st["step"]))
                elif "momentum_buffer" in st and st["momentum_buffer"] is not None:
                    self.m[o:o + p.numel()].copy_(st["momentum_buffer"].reshape(-1))
                    self._steps = max(self._steps, 1)
        for g, sg in zip(self.param_groups, state_dict["param_groups"]):
            for k, v in sg.items():
                if k != "params":
                    g[k] = v
