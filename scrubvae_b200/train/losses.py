"""`get_batch_loss` with the reference's signature and result (reference train/losses.py:182-324),
restricted to the loss keys on the built hot path: prior (:138-146), jpe (:148-171), root (:216-219),
<feat>_gr (:267-284, nested normalisation reproduced) and total (:320-322).

All arithmetic runs in libscv.so kernels (scv_recon_loss, scv_kl, scv_gr_loss, scv_loss_finalize);
`batch_loss["total"].backward()` runs the engine's backward launch list and leaves the gradients in
`param.grad` (views of one flat buffer)."""
from __future__ import annotations

import torch

UNSUPPORTED = ("rotation", "total_correlation")


class _StepLoss(torch.autograd.Function):
    """Connects the kernel-computed loss vector to autograd: backward launches the engine backward."""

    @staticmethod
    def forward(ctx, anchor, plan):
        ctx.plan = plan
        return plan.loss_out.clone()

    @staticmethod
    def backward(ctx, g):
        plan = ctx.plan
        n = plan.gscale.numel()
        # d total / d loss_k = g_total * scale_k + g_k   (device-side, no host sync)
        torch.addcmul(g[:n], plan.loss_scale, g[n], out=plan.gscale)
        eng = plan.eng
        plan.backward(getattr(eng, "comm", None))  # eng.comm: data-parallel gradient all-reduce (parallel.py)
        for p, gv in zip(eng.params, eng.gviews):
            if p.grad is None or p.grad is gv:
                p.grad = gv
            else:
                p.grad = p.grad + gv
        eng.clip = None
        return None, None


def get_batch_loss(model, data, data_o, loss_scale, disentangle_config):
    for k in UNSUPPORTED:
        if k in loss_scale.keys():
            raise NotImplementedError(f"scrubvae_b200: loss '{k}' is outside the built hot path (SURVEY.md §8)")
    plan = data_o.get("_plan")
    if plan is None:
        raise RuntimeError("scrubvae_b200.get_batch_loss needs the data_o returned by model(data)")
    methods = disentangle_config["method"]
    for method in methods:
        if method not in ("conditional", "grad_reversal", "moving_avg_lsq", "qda", "moving_avg", "direct_lsq"):
            raise NotImplementedError(f"scrubvae_b200: scrubbing method '{method}' is outside the built hot path")
    gr_keys = list(methods.get("grad_reversal", []))
    if gr_keys != plan.eng.gr_keys:
        raise RuntimeError("disentangle_config['method']['grad_reversal'] does not match the model's heads")
    if "mcmi" in loss_scale.keys():  # reference :221-225: the estimator (or zero while there is none yet)
        plan.enable_mcmi(disentangle_config["bandwidth"], disentangle_config.get("var_mode") or "sphere")
        plan.mi_set(model.mi_estimator)
    plan.loss(data, loss_scale)
    if torch.is_grad_enabled():
        vec = _StepLoss.apply(plan.anchor, plan)
    else:
        vec = plan.loss_out.clone()
    batch_loss = {}
    for i, name in enumerate(plan.loss_names):
        if name in ("prior", "jpe", "root", "mcmi") and name not in loss_scale.keys():
            continue
        batch_loss[name] = vec[i]
    for k in batch_loss:
        loss_scale[k]  # KeyError for a produced key without a scale, as in the reference (:320-322)
    batch_loss["total"] = vec[len(plan.loss_names)]
    return batch_loss
