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
        # the keys torch.optim.Adam / AdamW / SGD read in step(): a checkpoint written by state_dict() below loads into
        # the reference's torch optimizers (train/trainer.py:60-70, :81-87) and steps there
        if kind == "sgd":
            defaults.update(dampening=0, nesterov=True, maximize=False, foreach=None, differentiable=False, fused=None)
        else:
            defaults.update(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False,
                            fused=None, decoupled_weight_decay=(kind == "adamw"))
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
            # per-step [lr, step] reach the kernel through a RING of pinned slots: an asynchronous copy reads pinned
            # memory when the DMA executes, not when it is enqueued, so a single slot would be overwritten by the host
            # running ahead of the GPU (steps k+1.. queued before step k's copy ran).  A slot is reused only after the
            # copy that read it has completed (event per slot).
            cuda = eng.device.type == "cuda"
            self.hyper_ring = torch.zeros(self.RING, 2, dtype=torch.double, pin_memory=cuda)
            self.hyper_events = [None] * self.RING
            self.hyper_pos = 0
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

    RING = 16

    def push_hyper(self):
        """Enqueues the copy of this step's [lr, step count] to the device vector the optimizer kernel reads."""
        i = self.hyper_pos
        self.hyper_pos = (i + 1) % self.RING
        ev = self.hyper_events[i]
        if ev is not None:
            ev.synchronize()  # only ever waits if the host is RING steps ahead of the device
        slot = self.hyper_ring[i]
        slot[0] = float(self.param_groups[0]["lr"])
        slot[1] = float(self._steps)
        self.hyper.copy_(slot, non_blocking=True)
        if self.hyper.is_cuda:
            ev = ev or torch.cuda.Event()
            ev.record()
            self.hyper_events[i] = ev

    def state_dict(self):
        eng = self._bind()
        eng.ensure_flat()  # moments may live in the packed layout (engine.TrainStep(resident=True))
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None):
        eng = self._bind()
        eng.ensure_flat()
        grp = self.param_groups[0]
        self._steps += 1
        self.push_hyper()
        clip = getattr(eng, "clip", None)
        b1 = grp["momentum"] if self.kind == "sgd" else grp["betas"][0]
        eng.ops.optim_step(eng.flat, eng.gflat, self.m, self.v, eng.n_flat, clip[0] if clip else None,
                           clip[1] if clip else 0.0, self.grad_scale, float(grp["lr"]), b1, grp["betas"][1],
                           grp["eps"], grp["weight_decay"], self._steps, KIND[self.kind], hyper=self.hyper)
        eng.clip = None
        eng.resident_valid = False  # `flat` was updated here: a resident TrainStep must re-import
        if self.kind != "sgd":
            for st in self.state.values():
                if "step" in st:
                    st["step"] = torch.tensor(float(self._steps))
        return None

    def load_state_dict(self, state_dict):
        eng = self._bind()
        eng.ensure_flat()
        eng.resident_valid = False
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
                if k != "params" and k in g:  # hyper-parameters this optimizer knows; foreign keys are ignored
                    g[k] = v
