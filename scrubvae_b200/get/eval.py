"""`scrubvae_b200.get.latents` — the reference's latent export (get/eval.py:8-70): embed a whole loader with the
eval-mode encoder (z = mu), cache the result as `<out_path>/latents/<split>_<epoch>.npy`, report the latent dimensions
whose standard deviation over the dataset exceeds 0.1.  NOT for training (no gradients, model.eval())."""
from pathlib import Path

import numpy as np
import torch


def latents(config, model=None, epoch=None, loader=None, device="cuda", train_val_test="test", overwrite=False):
    if model is not None:
        model.eval()
    latent_path = "{}/latents/{}_{}.npy".format(config["out_path"], train_val_test, epoch)

    if not Path(latent_path).exists() or overwrite:
        print("Latent projections not found - Embedding dataset ...")
        chunks = []
        with torch.no_grad():
            for data in loader:
                data = {k: v.to(device) for k, v in data.items() if k in ["x6d", "root"]}
                # the plan's mu buffer is overwritten by the next batch: copy before moving on
                chunks.append(model.encode(data)["mu"].detach().to("cpu", copy=True))
        lat = torch.cat(chunks, axis=0)
        Path(latent_path).parent.mkdir(parents=True, exist_ok=True)
        np.save(latent_path, np.array(lat))
    else:
        print("Found existing latent projections - Loading ...")
        lat = np.load(latent_path)
        if loader is not None and hasattr(loader, "dataset"):
            assert lat.shape[0] == len(loader.dataset)
        lat = torch.tensor(lat)

    nonzero_std_z = torch.where(lat.std(dim=0) > 0.1)[0]
    print("Latent dimensions with variance over the dataset > 0.1 : {}".format(len(nonzero_std_z)))
    print(lat.std(dim=0))
    return lat
