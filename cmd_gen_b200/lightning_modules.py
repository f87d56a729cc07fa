"""Drop-in for the sampling half of DiffPhar/lightning_modules.py::PharPocketDDPM — without
pytorch_lightning or BioPython (neither is installed here).

Kept: the constructor's hyper-parameter names, the ``ddpm.dynamics.*`` state-dict keys,
``load_from_checkpoint`` (reads ``hyper_parameters`` + ``state_dict`` from a Lightning ``.ckpt`` with plain
``torch.load``), and ``generate_phars`` with its exact output structure — including the quirk that
``Molecule_k`` is the k-th point slot aggregated over all samples (lightning_modules.py:512-541).
Training / validation hooks are out of scope (they are not on the sampling path).
"""
from __future__ import annotations

from argparse import Namespace
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import pdb as pdbio
from .constants import FLOAT_TYPE, INT_TYPE, dataset_params
from .equivariant_diffusion.conditional_model import ConditionalDDPM
from .equivariant_diffusion.dynamics import EGNNDynamics
from .utils import batch_to_list, scatter_mean


def _ns(obj):
    """hyper-parameters arrive as argparse.Namespace (Lightning pickles them) or plain dicts."""
    if isinstance(obj, Namespace):
        return obj
    if isinstance(obj, dict):
        return Namespace(**obj)
    raise TypeError(f"cannot interpret hyper-parameter block of type {type(obj)}")


class PharPocketDDPM(torch.nn.Module):
    def __init__(self, outdir=None, dataset="crossdock", datadir=None, batch_size=1, lr=0.0, egnn_params=None,
                 diffusion_params=None, num_workers=0, augment_noise=0, augment_rotation=False, clip_grad=False,
                 eval_epochs=0, eval_params=None, mode="pocket_conditioning", node_histogram=None,
                 pocket_representation="CA", precision="f16fast"):
        super().__init__()
        if mode != "pocket_conditioning":
            raise NotImplementedError(f"mode '{mode}': only 'pocket_conditioning' (ConditionalDDPM) is on the "
                                      "accelerated path; every reference config uses it")
        assert pocket_representation in {"CA", "full-atom"}
        egnn_params, diffusion_params = _ns(egnn_params), _ns(diffusion_params)
        self.hparams = dict(outdir=outdir, dataset=dataset, datadir=datadir, batch_size=batch_size, lr=lr,
                            egnn_params=egnn_params, diffusion_params=diffusion_params, num_workers=num_workers,
                            augment_noise=augment_noise, augment_rotation=augment_rotation, clip_grad=clip_grad,
                            eval_epochs=eval_epochs, eval_params=eval_params, mode=mode,
                            node_histogram=node_histogram, pocket_representation=pocket_representation)
        self.mode = mode
        self.pocket_representation = pocket_representation
        self.dataset_name = dataset
        self.dataset_info = dataset_params[dataset]
        self.T = diffusion_params.diffusion_steps
        self.phar_type_encoder = self.dataset_info["phar_encoder"]
        self.phar_type_decoder = self.dataset_info["phar_decoder"]
        ca = pocket_representation == "CA"
        self.pocket_type_encoder = self.dataset_info["aa_encoder" if ca else "atom_encoder"]
        self.pocket_type_decoder = self.dataset_info["aa_decoder" if ca else "atom_decoder"]
        self.phar_nf = len(self.phar_type_decoder)
        self.aa_nf = len(self.pocket_type_decoder)
        self.x_dims = 3
        ep = vars(egnn_params)
        dynamics = EGNNDynamics(
            phar_nf=self.phar_nf, residue_nf=self.aa_nf, n_dims=self.x_dims, joint_nf=ep["joint_nf"],
            device=ep.get("device", "cuda") if torch.cuda.is_available() else "cpu", hidden_nf=ep["hidden_nf"],
            act_fn=torch.nn.SiLU(), n_layers=ep["n_layers"], attention=ep["attention"], tanh=ep["tanh"],
            norm_constant=ep["norm_constant"], inv_sublayers=ep["inv_sublayers"],
            sin_embedding=ep.get("sin_embedding", False), normalization_factor=ep["normalization_factor"],
            aggregation_method=ep["aggregation_method"], edge_cutoff=ep.get("edge_cutoff"),
            update_pocket_coords=False, precision=precision)
        if node_histogram is None:
            node_histogram = np.ones((2, 2))
        self.ddpm = ConditionalDDPM(
            dynamics=dynamics, phar_nf=self.phar_nf, residue_nf=self.aa_nf, n_dims=self.x_dims,
            timesteps=diffusion_params.diffusion_steps, noise_schedule=diffusion_params.diffusion_noise_schedule,
            noise_precision=diffusion_params.diffusion_noise_precision,
            loss_type=diffusion_params.diffusion_loss_type, norm_values=diffusion_params.normalize_factors,
            size_histogram=node_histogram)

    @property
    def device(self):
        return next(self.parameters()).device

    # ------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **overrides):
        """Reads a pytorch-lightning checkpoint: ``hyper_parameters`` -> constructor, ``state_dict`` -> weights
        (generate_phars.py:32-34).  Extra keyword arguments override hyper-parameters (e.g. precision)."""
        ckpt = torch.load(str(checkpoint_path), map_location="cpu", weights_only=False)
        hp = dict(ckpt["hyper_parameters"])
        hp.update(overrides)
        model = cls(**hp)
        missing, unexpected = model.load_state_dict(ckpt["state_dict"], strict=False)
        missing = [k for k in missing if not k.endswith("buffer")]
        if missing or unexpected:
            raise KeyError(f"checkpoint / model mismatch: missing {missing[:5]}, unexpected {unexpected[:5]}")
        if map_location is not None:
            model = model.to(map_location)
        return model

    # ------------------------------------------------------------------
    def _pocket_from_pdb(self, pdb_file, pocket_ids, ref_ligand):
        residues = pdbio.read_pdb(pdb_file)
        if pocket_ids is not None:
            chosen = pdbio.select_residues(residues, pocket_ids)
        else:
            chosen = pdbio.pocket_from_ligand(residues, ref_ligand)
        if not chosen:
            raise ValueError("the pocket definition selected no residues")
        xyz, types = pdbio.pocket_tensors(chosen, self.pocket_representation, self.pocket_type_encoder)
        return (torch.tensor(xyz, device=self.device, dtype=FLOAT_TYPE),
                torch.tensor(types, device=self.device, dtype=INT_TYPE))

    @torch.no_grad()
    def generate_phars(self, pdb_file, n_samples, pocket_ids=None, ref_ligand=None, num_nodes_phar=None,
                       sanitize=False, largest_frag=False, relax_iter=0, timesteps=None, **kwargs):
        """lightning_modules.py:385-541.  ``sanitize``/``largest_frag``/``relax_iter`` and the inpainting
        kwargs are accepted and ignored, as on the reference's conditional path."""
        assert (pocket_ids is None) ^ (ref_ligand is None)
        pocket_coord, pocket_types = self._pocket_from_pdb(pdb_file, pocket_ids, ref_ligand)
        pocket_one_hot = F.one_hot(pocket_types, num_classes=len(self.pocket_type_encoder))
        n_res = len(pocket_coord)
        pocket = {
            "x": pocket_coord.repeat(n_samples, 1),
            "one_hot": pocket_one_hot.repeat(n_samples, 1),
            "size": torch.tensor([n_res] * n_samples, device=self.device, dtype=INT_TYPE),
            "mask": torch.repeat_interleave(torch.arange(n_samples, device=self.device, dtype=INT_TYPE), n_res),
        }
        pocket_com_before = scatter_mean(pocket["x"], pocket["mask"])
        if num_nodes_phar is None:
            num_nodes_phar = self.ddpm.size_distribution.sample_conditional(n1=None, n2=pocket["size"])
        xh_phar, xh_pocket, phar_mask, pocket_mask = self.ddpm.sample_given_pocket(
            pocket, num_nodes_phar, timesteps=timesteps)

        # back to the original pocket frame (lightning_modules.py:495-504)
        pocket_com_after = scatter_mean(xh_pocket[:, :self.x_dims], pocket_mask)
        shift = pocket_com_before - pocket_com_after
        xh_pocket[:, :self.x_dims] += shift[pocket_mask]
        xh_phar[:, :self.x_dims] += shift[phar_mask]

        phar_mask = phar_mask.cpu()
        x = xh_phar[:, :self.x_dims].detach().cpu()
        phar_type = xh_phar[:, self.x_dims:].argmax(1).detach().cpu()
        phar_to_coords = {}
        for coords_batch, types in zip(batch_to_list(x, phar_mask), batch_to_list(phar_type, phar_mask)):
            # the slot counter restarts for every sample: "Molecule_k" collects the k-th point of ALL samples
            for slot, (t, coords) in enumerate(zip(types.tolist(), coords_batch), start=1):
                bucket = phar_to_coords.setdefault(f"Molecule_{slot}", {})
                bucket.setdefault(self.phar_type_decoder[t], []).append(coords)
        return phar_to_coords


def make_checkpoint(path, egnn_params: Optional[dict] = None, diffusion_params: Optional[dict] = None,
                    dataset="crossdock", pocket_representation="CA", seed=0, node_histogram=None):
    """Writes a Lightning-format checkpoint with RANDOM-INIT weights of the configured architecture (the
    reference ships no checkpoint and there is no network): used by the CLI smoke test and benchmarks."""
    from .config import DynamicsConfig
    from .weights import init_weights
    egnn = dict(device="cuda", edge_cutoff=6.0, joint_nf=32, hidden_nf=256, n_layers=5, attention=True, tanh=True,
                norm_constant=1, inv_sublayers=1, sin_embedding=False, aggregation_method="sum",
                normalization_factor=100)
    egnn.update(egnn_params or {})
    diff = dict(diffusion_steps=500, diffusion_noise_schedule="polynomial_2", diffusion_noise_precision=1.0e-5,
                diffusion_loss_type="l2", normalize_factors=[1, 4])
    diff.update(diffusion_params or {})
    info = dataset_params[dataset]
    res_nf = len(info["aa_decoder"] if pocket_representation == "CA" else info["atom_decoder"])
    cfg = DynamicsConfig(phar_nf=len(info["phar_decoder"]), residue_nf=res_nf, joint_nf=egnn["joint_nf"],
                         hidden_nf=egnn["hidden_nf"], n_layers=egnn["n_layers"], inv_sublayers=egnn["inv_sublayers"],
                         attention=egnn["attention"], tanh=egnn["tanh"], norm_constant=float(egnn["norm_constant"]),
                         normalization_factor=float(egnn["normalization_factor"]),
                         aggregation_method=egnn["aggregation_method"], edge_cutoff=egnn.get("edge_cutoff"))
    from .schedule import gamma_table
    state = {f"ddpm.dynamics.{k}": v for k, v in init_weights(cfg, seed).items()}
    state["ddpm.gamma.gamma"] = gamma_table(diff["diffusion_noise_schedule"], diff["diffusion_steps"],
                                            diff["diffusion_noise_precision"])
    state["ddpm.buffer"] = torch.zeros(1)
    hp = dict(outdir=None, dataset=dataset, datadir=None, batch_size=1, lr=1e-4, egnn_params=Namespace(**egnn),
              diffusion_params=Namespace(**diff), num_workers=0, augment_noise=0, augment_rotation=False,
              clip_grad=True, eval_epochs=0, eval_params=Namespace(), mode="pocket_conditioning",
              node_histogram=node_histogram if node_histogram is not None else np.ones((16, 512)),
              pocket_representation=pocket_representation)
    torch.save({"state_dict": state, "hyper_parameters": hp, "pytorch-lightning_version": "1.8.5"}, str(path))
    return cfg
