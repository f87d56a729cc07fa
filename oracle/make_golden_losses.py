"""TEST INFRASTRUCTURE — fixture for the loss / NLL terms of ConditionalDDPM.forward
(conditional_model.py:198-320), minted by running the UNMODIFIED reference (oracle/ref_shims.py) in the build
container with injected timesteps and noise, in training and in evaluation mode, fp32 and fp64.

    python -m oracle.make_golden_losses          ->  tests/golden/losses_ca_small.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cmd_gen_b200.config import DynamicsConfig          # noqa: E402
from cmd_gen_b200.weights import init_weights           # noqa: E402
from cmd_gen_b200.synthetic import make_pocket_batch, draw_noise   # noqa: E402
from oracle import ref_shims                            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TERMS = ["delta_log_px", "error_t_phar", "error_t_pocket", "SNR_weight", "loss_0_x_phar", "loss_0_x_pocket", "loss_0_h",
         "neg_log_constants", "kl_prior", "log_pN", "t_int", "xh_phar_hat"]


def main():
    cfg = DynamicsConfig()
    sizes, counts = [20, 35, 28], [4, 6, 5]
    state = init_weights(cfg, seed=0)
    hist = np.ones((16, 64))
    pocket0 = make_pocket_batch(sizes, cfg.residue_nf, seed=21)
    B, n_p = len(sizes), sum(counts)
    counts_t = torch.tensor(counts)
    mask_p = torch.repeat_interleave(torch.arange(B), counts_t)
    gen = torch.Generator().manual_seed(5)
    com = torch.stack([pocket0["x"][pocket0["mask"] == b].mean(0) for b in range(B)])
    phar_x = com[mask_p] + 3.0 * torch.randn(n_p, 3, generator=gen)
    phar_types = torch.randint(0, cfg.phar_nf, (n_p,), generator=gen)
    noise = draw_noise(2, n_p, 3 + cfg.phar_nf, seed=9)
    t_train = torch.tensor([[0.0], [250.0], [500.0]])
    t_eval = torch.tensor([[1.0], [250.0], [500.0]])
    out = dict(sizes=np.array(sizes), counts=np.array(counts), phar_x=phar_x.numpy(), phar_types=phar_types.numpy(),
               pocket_x=pocket0["x"].numpy(), pocket_one_hot=pocket0["one_hot"].numpy(), pocket_mask=pocket0["mask"].numpy(),
               noise=noise.numpy(), t_train=t_train.numpy(), t_eval=t_eval.numpy(), wseed=0)
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        ddpm = ref_shims.build_reference_model(cfg, state, T=500, dtype=dt, size_histogram=hist)
        for mode, t_fix in (("train", t_train), ("eval", t_eval)):
            ddpm.train(mode == "train")
            phar = {"x": phar_x.clone().to(dt), "one_hot": torch.nn.functional.one_hot(phar_types, cfg.phar_nf).to(dt),
                    "size": counts_t.clone(), "mask": mask_p.clone()}
            pocket = {"x": pocket0["x"].clone().to(dt), "one_hot": pocket0["one_hot"].clone().to(dt),
                      "size": pocket0["size"].clone(), "mask": pocket0["mask"].clone()}
            real_randint = torch.randint
            torch.randint = lambda *a, **k: t_fix.clone().long()
            try:
                with torch.no_grad(), ref_shims.InjectedNoise(ddpm, noise.to(dt)):
                    res = ddpm(phar, pocket, return_info=True)
            finally:
                torch.randint = real_randint
            for name, v in zip(TERMS, res[:-1]):
                out[f"{mode}_{tag}_{name}"] = np.asarray(v.detach().to(torch.float64).numpy() if torch.is_tensor(v) else v, dtype=np.float64)
            for k, v in res[-1].items():
                out[f"{mode}_{tag}_info_{k}"] = float(v)
            print(mode, tag, "error_t", out[f"{mode}_{tag}_error_t_phar"], "loss_0_h", out[f"{mode}_{tag}_loss_0_h"], "kl", out[f"{mode}_{tag}_kl_prior"])
    np.savez_compressed(os.path.join(OUT, "losses_ca_small.npz"), **out)


if __name__ == "__main__":
    main()
