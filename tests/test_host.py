"""CPU: host-side logic, the weight ABI, and that the C-ABI library loads and exports
every symbol include/diffphar_b200.h declares (no compute without a GPU)."""
import json
import os
import re

import pytest
import torch

from cmd_gen_b200 import _lib
from cmd_gen_b200.config import DynamicsConfig, weight_count, weight_spec
from cmd_gen_b200.weights import init_weights, pack_blob
from tests.helpers import CASE_CFG, GOLDEN, ROOT, case_config


@pytest.fixture(scope="module")
def lib():
    from cmd_gen_b200.build import build
    build()
    return _lib.load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "diffphar_b200.h")).read()
    declared = set(re.findall(r"\b(dp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dp_abi_version() == 3


def test_no_gpu_means_loud_failure(lib):
    if lib.dp_device_count() > 0:
        pytest.skip("a B200 is visible")
    with pytest.raises(_lib.DiffPharError):
        _lib.Handle(DynamicsConfig(), "cuda:0")
    with pytest.raises(_lib.DiffPharError):
        _lib.Handle(DynamicsConfig(), "cpu")


@pytest.mark.parametrize("name", list(CASE_CFG))
def test_weight_abi_matches_reference_state_dict(name):
    ref = json.load(open(os.path.join(GOLDEN, "state_keys.json")))[name]
    cfg = case_config(name)
    spec = {k: list(s) for k, s in weight_spec(cfg)}
    assert spec == ref["dynamics"]
    from cmd_gen_b200.equivariant_diffusion.dynamics import EGNNDynamics
    dyn = EGNNDynamics(cfg.phar_nf, cfg.residue_nf, 3, joint_nf=cfg.joint_nf, hidden_nf=256, n_layers=cfg.n_layers,
                       attention=cfg.attention, tanh=cfg.tanh, norm_constant=cfg.norm_constant,
                       inv_sublayers=cfg.inv_sublayers, aggregation_method=cfg.aggregation_method,
                       condition_time=cfg.condition_time, update_pocket_coords=False, edge_cutoff=cfg.edge_cutoff)
    ours = {k: list(v.shape) for k, v in dyn.state_dict().items()}
    assert ours == ref["dynamics"]
    state = init_weights(cfg, 0)
    dyn.load_state_dict(state, strict=True)
    blob = pack_blob(cfg, dyn.state_dict())
    assert blob.numel() == weight_count(cfg)
    if name == "ca_small":
        assert weight_count(cfg) + 501 == 2987815    # SURVEY.md §8b also counts gamma[501]
    from cmd_gen_b200.equivariant_diffusion.conditional_model import ConditionalDDPM
    ddpm = ConditionalDDPM(dyn, cfg.phar_nf, cfg.residue_nf, 3, [[1.0, 1.0], [1.0, 1.0]], timesteps=500,
                           noise_schedule="polynomial_2", noise_precision=1e-5, loss_type="l2", norm_values=(1.0, 4.0))
    extra = {k: list(v.shape) for k, v in ddpm.state_dict().items() if not k.startswith("dynamics.")}
    assert extra == ref["ddpm_extra"]


def test_pack_blob_rejects_bad_state():
    cfg = DynamicsConfig(n_layers=1)
    st = init_weights(cfg, 0)
    bad = dict(st); bad.pop("egnn.embedding.bias")
    with pytest.raises(KeyError):
        pack_blob(cfg, bad)
    bad = dict(st); bad["egnn.embedding.bias"] = torch.zeros(3)
    with pytest.raises(ValueError):
        pack_blob(cfg, bad)


def test_unsupported_configurations_raise():
    from cmd_gen_b200.equivariant_diffusion.dynamics import EGNNDynamics
    with pytest.raises(NotImplementedError):
        EGNNDynamics(8, 20, 3, hidden_nf=64)
    with pytest.raises(NotImplementedError):
        EGNNDynamics(8, 20, 3, hidden_nf=256, mode="gnn_dynamics")
    with pytest.raises(NotImplementedError):
        EGNNDynamics(8, 20, 3, hidden_nf=256, sin_embedding=True)


def test_conditional_ddpm_error_conventions():
    from cmd_gen_b200.equivariant_diffusion.dynamics import EGNNDynamics
    from cmd_gen_b200.equivariant_diffusion.conditional_model import ConditionalDDPM
    dyn = EGNNDynamics(8, 20, 3, joint_nf=32, hidden_nf=256, n_layers=1, update_pocket_coords=False)
    ddpm = ConditionalDDPM(dyn, 8, 20, 3, [[1.0]], timesteps=50, noise_schedule="polynomial_2",
                           noise_precision=1e-5, loss_type="l2", norm_values=(1.0, 4.0))
    with pytest.raises(NotImplementedError):
        ddpm.sample()
    with pytest.raises(NotImplementedError):
        ddpm.sample_normal()
    pocket = {"x": torch.zeros(4, 3), "one_hot": torch.zeros(4, 20), "size": torch.tensor([4]),
              "mask": torch.zeros(4, dtype=torch.int64)}
    with pytest.raises(AssertionError):
        ddpm.sample_given_pocket(pocket, torch.tensor([2]), return_frames=0)
    with pytest.raises(AssertionError):
        ddpm.sample_given_pocket(pocket, torch.tensor([2]), return_frames=3, timesteps=50)
    bad = EGNNDynamics(8, 20, 3, joint_nf=32, hidden_nf=256, n_layers=1, update_pocket_coords=True)
    with pytest.raises(AssertionError):
        ConditionalDDPM(bad, 8, 20, 3, [[1.0]], timesteps=50, noise_schedule="polynomial_2",
                        noise_precision=1e-5, loss_type="l2", norm_values=(1.0, 4.0))
    with pytest.raises(ValueError):          # en_diffusion.py:64-77 norm-value sanity check
        ConditionalDDPM(dyn, 8, 20, 3, [[1.0]], timesteps=50, noise_schedule="cosine",
                        noise_precision=1e-4, loss_type="l2", norm_values=(1.0, 4.0))


def test_precision_modes_match_the_header_enum():
    """config.PRECISION_MODES (the names the Python mirror and the CLI accept) against dp_precision in the header."""
    from cmd_gen_b200.config import PRECISION_MODES
    header = open(os.path.join(ROOT, "include", "diffphar_b200.h")).read()
    enum = dict((n, int(v)) for n, v in re.findall(r"\b(DP_(?:FP32|TF32|BF16|F16|F16_FAST|F16_FAST32))\s*=\s*(\d+)", header))
    want = {"fp32": "DP_FP32", "tf32": "DP_TF32", "bf16": "DP_BF16", "f16": "DP_F16", "f16fast": "DP_F16_FAST",
            "f16fast32": "DP_F16_FAST32"}
    assert set(PRECISION_MODES) == set(want)
    for name, sym in want.items():
        assert PRECISION_MODES[name] == enum[sym], name


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """The product path has no fallback: a missing build raises (also through the DIFFPHAR_LIB A/B override)."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("DIFFPHAR_LIB", str(tmp_path / "no_such_build.so"))
    with pytest.raises(_lib.DiffPharError):
        _lib.load_library()
    monkeypatch.delenv("DIFFPHAR_LIB")
    monkeypatch.setattr(_lib, "_lib", None)
    assert _lib.load_library() is not None


def test_analysis_has_no_cpu_fallback(lib):
    from cmd_gen_b200.analysis import phar_statistics
    with pytest.raises(_lib.DiffPharError):
        phar_statistics({"Molecule_0": {"Donor": [[0.0, 0.0, 0.0]]}}, [0.0, 0.0, 0.0], device="cpu")
