"""Step loop with the reference's function names, signatures and call order
(reference train/trainer.py:26-212, :306-399): predict_batch -> get_batch_loss -> zero grads ->
total.backward() -> clip_grad_norm_(1e6) -> optimizer.step() -> scheduler.step().

Out of scope: test-time sklearn metrics, wandb logging, plotting, the non-GR scrubbers'
EMA updates and the MI estimator rebuild."""
from __future__ import annotations

import time
import weakref
from pathlib import Path

import torch
import torch.optim as optim

from .losses import get_batch_loss
from .optim import FusedOptimizer


class CyclicalBetaAnnealing:
    """Reference train/trainer.py:26-40."""

    def __init__(self, beta_max=1, len_cycle=100, R=0.5):
        self.beta_max, self.len_cycle, self.R = beta_max, len_cycle, R
        self.len_increasing = int(len_cycle * R)

    def get(self, epoch):
        remainder = (epoch - 1) % self.len_cycle
        if remainder >= self.len_increasing:
            return self.beta_max
        return self.beta_max * remainder / self.len_increasing


def get_beta_schedule(schedule, beta):
    """Reference train/trainer.py:43-51."""
    if schedule == "cyclical":
        print("Initializing cyclical beta annealing")
        return CyclicalBetaAnnealing(beta_max=beta)
    print("No beta annealing selected")
    return None


def get_optimizer_and_lr_scheduler(model, train_config, load_path=None, start_epoch=None):
    """Reference train/trainer.py:54-89; the optimizers are the fused single-launch versions."""
    if train_config["optimizer"] == "adam":
        print("Initializing Adam optimizer ...")
        optimizer = FusedOptimizer(model, lr=train_config["lr"], kind="adam")
    elif train_config["optimizer"] == "adamw":
        print("Initializing AdamW optimizer ...")
        optimizer = FusedOptimizer(model, lr=train_config["lr"], kind="adamw")
    elif train_config["optimizer"] == "sgd":
        print("Initializing SGD optimizer ...")
        optimizer = FusedOptimizer(model, lr=train_config["lr"], kind="sgd", momentum=0.2)
    else:
        raise ValueError("No valid optimizer selected")

    scheduler = None
    if train_config["lr_schedule"] == "cawr":
        print("Initializing cosine annealing w/warm restarts learning rate scheduler")
        scheduler = optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=50)
    elif train_config["lr_schedule"] is None:
        print("No learning rate scheduler selected")

    if load_path is not None:
        ck = Path("{}/checkpoints/epoch_{}.pth".format(load_path, start_epoch))
        if ck.exists():
            checkpoint = torch.load(ck, weights_only=False)
            optimizer.load_state_dict(checkpoint["optimizer"])
            scheduler = checkpoint["lr_scheduler"]
    return optimizer, scheduler


def clip_grad_norm_(parameters_or_model, max_norm):
    """torch.nn.utils.clip_grad_norm_ as used at reference train/trainer.py:164.  The global L2 norm
    is reduced by one kernel over the flat gradient buffer; the clip coefficient
    min(1, max_norm/(norm+1e-6)) is applied inside the fused optimizer launch that follows."""
    model = parameters_or_model
    eng = model.engine
    if not hasattr(eng, "sumsq"):
        eng.sumsq = torch.zeros(1, dtype=torch.double, device=eng.device)
    eng.sumsq.zero_()
    eng.ops.sumsq(eng.gflat, eng.n_flat, eng.sumsq)
    eng.clip = (eng.sumsq, float(max_norm))
    return eng.sumsq


def predict_batch(model, data, disentangle_keys=None):
    """Reference train/trainer.py:92-99."""
    data_i = {k: v for k, v in data.items() if (k in disentangle_keys) or (k in ["x6d", "root", "var"])}
    return model(data_i)


def _fused_train_epoch(config, model, loader, device, epoch, optimizer, scheduler, step_callback=None):
    """train-mode body of train_test_epoch as ONE captured launch sequence per step (engine.TrainStep: forward, losses,
    backward, gradient all-reduce, clip, optimizer replayed from a CUDA graph) with the next batch's host->device copy
    prefetched on a side stream while the current step runs.  Same kernels, order and results as the piecewise
    path below (predict_batch -> get_batch_loss -> backward -> clip_grad_norm_ -> optimizer.step)."""
    from ..engine import TrainStep
    eng = model.engine
    # captured steps, staging slots and the copy stream live ON the engine: a new engine (model.to(...), .float() ...)
    # starts with none of them, so no graph bound to the old engine's buffers can be replayed
    steps = eng.__dict__.setdefault("_train_steps", {})
    main = torch.cuda.current_stream()
    copy_stream = eng.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=eng.device))
    it = iter(loader)
    # two persistent device staging slots (no per-step allocation): slot s is refilled on the copy stream once the
    # step that read it has been enqueued-and-passed on the main stream
    stage = eng.__dict__.setdefault("_stage", [dict(), dict()])
    consumed = [None, None]
    count = [0]

    def fetch():
        try:
            host = next(it)
        except StopIteration:
            return None
        if any(torch.is_tensor(v) and v.is_cuda for v in host.values()):
            # device-resident loader (data.DevicePoseWindows): the batch was produced by kernels on the MAIN stream and
            # is consumed there (plan.load_inputs) — staging it on the copy stream would race with its producer
            return {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}, None, None
        slot = count[0] & 1
        count[0] += 1
        bufs = stage[slot]
        for k, v in host.items():
            if torch.is_tensor(v) and (k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype):
                bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=device)
        if consumed[slot] is not None:
            copy_stream.wait_event(consumed[slot])
        else:
            copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            dev = {}
            for k, v in host.items():
                if torch.is_tensor(v):
                    bufs[k].copy_(v, non_blocking=True)
                    dev[k] = bufs[k]
                else:
                    dev[k] = v
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dev, ev, slot

    mi_reset = set()

    def step_for(B):
        """The captured step for this batch size, optimizer and the hyper-parameters baked in at capture."""
        grp = optimizer.param_groups[0]
        mikey = (config["disentangle"].get("bandwidth"), config["disentangle"].get("var_mode")) \
            if "mcmi" in config["loss"].keys() else None
        key = (B, id(optimizer), optimizer.kind, tuple(grp["betas"]), grp["eps"], grp["weight_decay"], grp["momentum"],
               optimizer.grad_scale, mikey)
        hit = steps.get(key)
        if hit is not None and hit[0]() is optimizer:
            if mikey is not None and B not in mi_reset:
                hit[1].plan.mi["valid"].zero_()  # model.mi_estimator = None at the start of every epoch (reference :124)
                mi_reset.add(B)
            return hit[1]
        mi_reset.add(B)
        mi = None
        if "mcmi" in config["loss"].keys():
            mi = dict(bandwidth=config["disentangle"]["bandwidth"], var_mode=config["disentangle"].get("var_mode") or "sphere")
        st = TrainStep(model, optimizer, config["loss"], B, max_norm=1e6, use_graph=True, comm=eng.comm, resident=True, mi=mi)
        if mi is not None:
            st.plan.mi["valid"].zero_()  # model.mi_estimator = None at the start of every epoch (reference :124)
        steps[key] = (weakref.ref(optimizer), st)
        return st

    # the steps of this epoch keep the weights' master in the packed GEMM layout: import what the parameters hold now
    # (reset_parameters, load_state_dict, a broadcast ... may have changed them since the last epoch), export at the end
    eng.ensure_flat()
    eng.resident_valid = False
    names = None
    acc = None
    n_batches = 0
    nxt = fetch()
    while nxt is not None:
        data, ev, slot = nxt
        nxt = fetch()  # overlaps with this step's kernels
        if ev is not None:
            main.wait_event(ev)
        B = data["x6d"].shape[0]
        st = step_for(B)
        if names is None:
            names = st.plan.loss_names
            for k in config["loss"].keys():  # every configured loss must be produced, and vice versa (reference :125,:181)
                if k not in names:
                    raise KeyError(k)
            for k in names:
                if k.endswith("_gr"):
                    config["loss"][k]
        st.plan.set_loss_scale(config["loss"])  # the KL weight may have been annealed since the last epoch
        vec = st.run(data)  # copies the slot into the plan's static input buffers, then replays the step
        if slot is not None:
            consumed[slot] = torch.cuda.Event()
            consumed[slot].record(main)
        if step_callback is not None:
            step_callback(n_batches, vec)
        if scheduler is not None:
            scheduler.step(epoch + n_batches / len(loader))
        acc = vec.clone() if acc is None else acc + vec
        n_batches += 1
    eng.ensure_flat()
    epoch_metrics = {}
    host = acc.cpu() if acc is not None else None
    for k in ["total"] + list(config["loss"].keys()):
        idx = len(names) if k == "total" else names.index(k)
        epoch_metrics[k] = host[idx].item() / max(1, n_batches)
        print("====> Epoch: {} Average {} loss: {:.4f}".format(epoch, k, epoch_metrics[k]))
    return epoch_metrics


def train_test_epoch(config, model, loader, device, epoch, optimizer=None, scheduler=None, mode="train",
                     step_callback=None):
    """Reference train/trainer.py:101-212.  `step_callback(batch_idx, loss_vector)` is an optional extension (the
    per-step losses as one device vector [jpe, root, prior, <feat>_gr..., total])."""
    if (mode == "train" and isinstance(optimizer, FusedOptimizer) and torch.device(device).type == "cuda"
            and (config.get("train") or {}).get("fused_step", True)):
        model.train()
        model.mi_estimator = None
        return _fused_train_epoch(config, model, loader, device, epoch, optimizer, scheduler, step_callback)
    if mode == "train":
        model.train()
        grad_env = torch.enable_grad
    elif mode == "test":
        model.eval()
        grad_env = torch.no_grad
    else:
        raise ValueError("This mode is not recognized.")

    with grad_env():
        model.mi_estimator = None
        epoch_metrics = {k: 0 for k in ["total"] + list(config["loss"].keys())}
        for batch_idx, data in enumerate(loader):
            data = {k: v.to(device, non_blocking=True) for k, v in data.items()}
            data_o = predict_batch(model, data, model.disentangle_keys)
            batch_loss = get_batch_loss(model, data, data_o, config["loss"], config["disentangle"])
            if mode == "train":
                for param in model.parameters():
                    param.grad = None
                batch_loss["total"].backward()
                clip_grad_norm_(model, max_norm=1e6)
                optimizer.step()
                if scheduler is not None:
                    scheduler.step(epoch + batch_idx / len(loader))
                # reference :169-178: running covariances of the moving-average scrubbers, from this batch's mu
                if "moving_avg_lsq" in model.disentangle.keys():
                    data_o["_plan"].mals_update()
                if "qda" in model.disentangle.keys():
                    data_o["_plan"].qda_update()
                if "moving_avg" in model.disentangle.keys():
                    data_o["_plan"].ma_update()
            epoch_metrics = {k: v + batch_loss[k].detach() for k, v in epoch_metrics.items()}
            if "mcmi" in config["loss"].keys():  # reference :184-199: estimator rebuilt from the updated encoder
                from ..model.disentangle import MutInfoEstimator
                var_y = data_o["var"].clone()
                updated = model.encode(data)
                model.mi_estimator = MutInfoEstimator(
                    x_s=updated["mu"].detach().clone(), y_s=var_y, bandwidth=config["disentangle"]["bandwidth"],
                    var_mode=config["disentangle"]["var_mode"],
                    model_var=updated["L"].detach().clone() if "L" in updated.keys() else None, device=device)
            if step_callback is not None:
                step_callback(batch_idx, batch_loss["total"].detach().reshape(1))

        for k, v in epoch_metrics.items():
            epoch_metrics[k] = v.item() / len(loader)
            print("====> Epoch: {} Average {} loss: {:.4f}".format(epoch, k, epoch_metrics[k]))
    return epoch_metrics


def train_epoch(config, model, loader, device, optimizer, scheduler, epoch):
    """Reference train/trainer.py:306-318."""
    return train_test_epoch(config, model, loader, device, epoch, optimizer, scheduler, mode="train")


def test_epoch(config, model, loader, device="cuda", epoch=0):
    """Reference train/trainer.py:215-303: eval-mode pass over the loader — per-loss epoch averages, the latent means of
    every window, and the generative-restrictiveness R^2 of each conditioned variable except `ids` (decode with a
    resampled conditional, FK, re-extract: eval.generative_restrictiveness).  Returns (epoch_metrics, z) like the
    reference.  The `mcmi` estimator rebuild (:228-252) belongs to the MutInfo scrubber, outside the built path."""
    from ..eval import generative_restrictiveness, r2_score
    print("Running test epoch")
    model.eval()
    with torch.no_grad():
        z = []
        epoch_metrics = {k: 0 for k in ["total"] + list(config["loss"].keys())}
        gen_res = {k1: {k2: [] for k2 in ["pred", "target"]} for k1 in model.disentangle_keys if k1 != "ids"}
        tree = getattr(getattr(loader, "dataset", None), "kinematic_tree", None) or model.kinematic_tree
        for batch_idx, data in enumerate(loader):
            data = {k: v.to(device) for k, v in data.items()}
            data_o = predict_batch(model, data, model.disentangle_keys)
            mu = data_o["mu"].clone().detach()  # (the plan's static buffer is overwritten by the decodes below)
            z += [mu]
            batch_metrics = get_batch_loss(model, data, data_o, config["loss"], config["disentangle"])
            batch_metrics = {k: v.detach().clone() for k, v in batch_metrics.items()}
            for key in gen_res.keys():
                key_pred, key_target = generative_restrictiveness(model, mu, data, key, tree)
                gen_res[key]["pred"] += [key_pred.detach().cpu()]
                gen_res[key]["target"] += [key_target.detach().cpu()]
            epoch_metrics = {k: v + batch_metrics[k] for k, v in epoch_metrics.items()}
    for k, v in epoch_metrics.items():
        epoch_metrics[k] = v.item() / len(loader)
        print("====> Epoch: {} Average {} loss: {:.4f}".format(epoch, k, epoch_metrics[k]))
    for key in gen_res.keys():
        epoch_metrics["r2_gen_restrict_{}".format(key)] = r2_score(torch.cat(gen_res[key]["target"], dim=0),
                                                                   torch.cat(gen_res[key]["pred"], dim=0))
    return epoch_metrics, torch.cat(z, dim=0).cpu()


def train(config, model, loader_dict, run=None, device="cuda"):
    """Epoch loop of reference train/trainer.py:321-399: beta annealing, train_epoch, per-epoch
    re-initialisation of the GR scrubbers, weight / optimizer checkpoints."""
    optimizer, scheduler = get_optimizer_and_lr_scheduler(
        model, config["train"], config["model"]["load_model"], config["model"]["start_epoch"])
    if "prior" in config["loss"].keys():
        beta_scheduler = get_beta_schedule(config["loss"]["prior"], config["train"]["beta_anneal"])
    else:
        beta_scheduler = None
    start = config["model"]["start_epoch"] or 0
    metrics = {}
    for epoch in range(start + 1, config["train"]["num_epochs"] + 1):
        if beta_scheduler is not None:
            config["loss"]["prior"] = beta_scheduler.get(epoch)
            print("Beta schedule: {:.3f}".format(config["loss"]["prior"]))
        t0 = time.time()
        train_metrics = train_epoch(config, model, loader_dict["train"], device, optimizer, scheduler, epoch)
        metrics = {"{}_train".format(k): v for k, v in train_metrics.items()}  # reference :351-361
        if "grad_reversal" in model.disentangle.keys():
            for key in model.disentangle["grad_reversal"].keys():
                model.disentangle["grad_reversal"][key].reset_parameters()
            from ..parallel import resync
            resync(model)  # data parallelism: every rank drew its own re-initialisation; rank 0's everywhere
        # automatically tuned forgetting factors of the moving-average scrubbers (reference :372-384)
        if "moving_avg_lsq" in model.disentangle.keys():
            for key in model.disentangle["moving_avg_lsq"].keys():
                metrics["lambda_mals_{}".format(key)] = model.disentangle["moving_avg_lsq"][key].lam1.detach().cpu().numpy()
        if "qda" in model.disentangle.keys():
            for key in model.disentangle["qda"].keys():
                metrics["lambda_qda_{}".format(key)] = model.disentangle["qda"][key].lama.detach().cpu().numpy()
        metrics["time"] = time.time() - t0
        if epoch % 5 == 0:
            Path("{}/weights".format(config["out_path"])).mkdir(parents=True, exist_ok=True)
            torch.save({k: v.cpu() for k, v in model.state_dict().items()},
                       "{}/weights/epoch_{}.pth".format(config["out_path"], epoch))
            # validation metrics every 5 epochs from epoch 50 on (reference :400-414; its sklearn decodability scores and
            # wandb logging that follow are outside the built path)
            if epoch >= 50 and loader_dict.get("val") is not None:
                test_metrics, _ = test_epoch(config, model, loader_dict["val"], device, epoch)
                metrics.update({"{}_test".format(k): v for k, v in test_metrics.items()})
        if epoch % 20 == 0:
            Path("{}/checkpoints".format(config["out_path"])).mkdir(parents=True, exist_ok=True)
            torch.save({"optimizer": optimizer.state_dict(), "lr_scheduler": scheduler},
                       "{}/checkpoints/epoch_{}.pth".format(config["out_path"], epoch))
    return model, metrics
