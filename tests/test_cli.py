"""Host-side drop-in surface of the reference's CLI layer: PDB pocket
ingestion without BioPython, Lightning-checkpoint loading without pytorch_lightning, generate_phars's
flags and its JSON structure.  CPU tests use a stubbed sampler; the GPU tests run the real thing."""
import json
import os

import numpy as np
import pytest
import torch

from cmd_gen_b200 import generate_phars as cli
from cmd_gen_b200 import pdb as pdbio
from cmd_gen_b200.constants import dataset_params
from cmd_gen_b200.lightning_modules import PharPocketDDPM, make_checkpoint
from cmd_gen_b200.synthetic import write_synthetic_pdb

PDB_TEXT = """\
ATOM      1  N   ALA A  10      10.000  10.000  10.000  1.00 20.00           N
ATOM      2  CA AALA A  10      11.000  10.000  10.000  0.40 20.00           C
ATOM      3  CA BALA A  10      11.500  10.000  10.000  0.60 20.00           C
ATOM      4  C   ALA A  10      12.000  10.500  10.000  1.00 20.00           C
ATOM      5  N   GLY A  11      14.000  10.000  10.000  1.00 20.00           N
ATOM      6  CA  GLY A  11      15.000  10.000  10.000  1.00 20.00           C
ATOM      7  H   GLY A  11      15.500  10.500  10.000  1.00 20.00           H
ATOM      8  CA  TRP A  12      30.000  10.000  10.000  1.00 20.00           C
ATOM      9  CA  MSE A  13      13.000  11.000  10.000  1.00 20.00           C
HETATM   10  C1  LIG A 501      13.000  10.000  10.000  1.00 20.00           C
HETATM   11  O   HOH A 601      13.000  10.000  11.000  1.00 20.00           O
ENDMDL
ATOM     12  CA  LYS A  99      13.000  10.000  10.000  1.00 20.00           C
"""


def write(tmp_path, text=PDB_TEXT):
    p = tmp_path / "t.pdb"
    p.write_text(text)
    return str(p)


def test_pdb_reader_altloc_models_and_hetero(tmp_path):
    res = pdbio.read_pdb(write(tmp_path))
    assert [(r.resname, r.resseq, r.hetero) for r in res] == [
        ("ALA", 10, False), ("GLY", 11, False), ("TRP", 12, False), ("MSE", 13, False), ("LIG", 501, True), ("HOH", 601, True)]
    ala = res[0]
    assert list(ala.atoms) == ["N", "CA", "C"]
    assert np.allclose(ala.atoms["CA"].coord, [11.5, 10, 10])          # highest-occupancy altloc wins
    # second MODEL is ignored
    assert all(r.resseq != 99 for r in res)


def test_pocket_from_ligand_matches_reference_rule(tmp_path):
    res = pdbio.read_pdb(write(tmp_path))
    got = pdbio.pocket_from_ligand(res, "A:501")
    # standard amino acids with any atom < 8 A from any ligand atom; MSE (non-standard), water and the ligand are out
    assert [r.resseq for r in got] == [10, 11]
    assert [r.resseq for r in pdbio.pocket_from_ligand(res, "A:501", dist_cutoff=18.0)] == [10, 11, 12]
    with pytest.raises(AssertionError):
        pdbio.pocket_from_ligand(res, "A:777")


def test_pocket_from_ligand_keeps_standard_amino_acids_written_as_hetatm(tmp_path):
    """utils.py:109-117 filters on is_aa(resname, standard=True) only — the record type plays no role — and its ligand
    skip (`residue.id[1] == resi`, int vs str) never fires: a peptide ligand made of a standard residue is kept."""
    text = PDB_TEXT.replace("ENDMDL", "HETATM   12  CA  SER A 502      12.000  11.000  10.000  1.00 20.00           C\nENDMDL")
    res = pdbio.read_pdb(write(tmp_path, text))
    assert [r.resseq for r in pdbio.pocket_from_ligand(res, "A:501")] == [10, 11, 502]
    assert [r.resseq for r in pdbio.pocket_from_ligand(res, "A:502")] == [10, 11, 502]      # the ligand itself is a standard residue


def test_resi_list_and_pocket_tensors(tmp_path):
    res = pdbio.read_pdb(write(tmp_path))
    sel = pdbio.select_residues(res, ["A:11", "A:10"])
    assert [r.resseq for r in sel] == [11, 10]
    with pytest.raises(KeyError):
        pdbio.select_residues(res, ["A:501"])                           # HETATM: not addressable with a blank hetero flag
    xyz, types = pdbio.pocket_tensors(sel, "CA", dataset_params["crossdock"]["aa_encoder"])
    assert xyz.dtype == np.float32 and np.allclose(xyz, [[15, 10, 10], [11.5, 10, 10]])
    assert types.tolist() == [dataset_params["crossdock"]["aa_encoder"]["G"], dataset_params["crossdock"]["aa_encoder"]["A"]]
    xyz, types = pdbio.pocket_tensors(sel, "full-atom", dataset_params["crossdock_full"]["atom_encoder"])
    assert xyz.shape == (5, 3)                                          # hydrogens dropped (lightning_modules.py:428-429)
    assert types.tolist() == [1, 0, 1, 0, 0]


def test_cli_flags_match_reference():
    a = cli.build_parser().parse_args(["m.ckpt", "--pdbfile", "x.pdb", "--ref_ligand", "A:1"])
    assert (a.n_samples, a.num_nodes_phar, a.resamplings, a.jump_length, a.timesteps) == (20, 3, 10, 1, None)
    assert a.resi_list is None and not (a.all_frags or a.sanitize or a.relax)
    b = cli.build_parser().parse_args(["m.ckpt", "--pdbfile", "x.pdb", "--resi_list", "A:1", "A:2", "--outdir", "o",
                                       "--n_samples", "4", "--num_nodes_phar", "7", "--timesteps", "50", "--relax"])
    assert b.resi_list == ["A:1", "A:2"] and b.n_samples == 4 and b.num_nodes_phar == 7 and b.timesteps == 50 and b.relax
    assert cli.OUTPUT_NAME == "phar_to_coords_no_tensor_PI3K_dul.json"


def test_checkpoint_round_trip_and_state_keys(tmp_path):
    path = tmp_path / "m.ckpt"
    cfg = make_checkpoint(path, egnn_params=dict(n_layers=2))
    ck = torch.load(str(path), weights_only=False)
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_keys.json")))
    # the reference's own state-dict keys (recorded from the unmodified modules) for the 5-block CA model
    ref = golden["ca_small"]
    want = {"ddpm.dynamics." + k for k in ref["dynamics"]
            if ".e_block_" not in k or int(k.split(".e_block_")[1].split(".")[0]) < 2}
    want |= {"ddpm." + k for k in ref["ddpm_extra"]}
    model = PharPocketDDPM.load_from_checkpoint(path, map_location="cpu", precision="fp32")
    assert set(model.state_dict()) == set(ck["state_dict"])
    assert want <= set(ck["state_dict"])
    assert model.ddpm.dynamics.cfg.n_layers == 2 and model.T == 500
    assert model.ddpm.dynamics.precision == "fp32"
    with pytest.raises(KeyError):
        bad = dict(ck); bad["state_dict"] = {k: v for k, v in ck["state_dict"].items() if "att_mlp" not in k}
        torch.save(bad, str(tmp_path / "bad.ckpt"))
        PharPocketDDPM.load_from_checkpoint(tmp_path / "bad.ckpt", map_location="cpu")


def test_generate_phars_structure_with_stub_sampler(tmp_path):
    """Molecule_k = k-th point slot over ALL samples; coordinates are shifted back to the PDB frame."""
    path = tmp_path / "m.ckpt"
    make_checkpoint(path, egnn_params=dict(n_layers=1))
    model = PharPocketDDPM.load_from_checkpoint(path, map_location="cpu")
    pdb = str(tmp_path / "p.pdb")
    write_synthetic_pdb(pdb, n_res=40)
    n_samples, n_pts = 3, 4
    seen = {}

    def stub(pocket, num_nodes_phar, timesteps=None):
        n_res = int(pocket["size"][0])
        seen["n_res"], seen["timesteps"] = n_res, timesteps
        mask_p = torch.repeat_interleave(torch.arange(n_samples), n_pts)
        x = torch.arange(n_samples * n_pts * 3, dtype=torch.float32).view(-1, 3)
        h = torch.nn.functional.one_hot(torch.arange(n_samples * n_pts) % 8, 8).float()
        com = torch.stack([pocket["x"][pocket["mask"] == b].mean(0) for b in range(n_samples)])
        shifted = pocket["x"] - com[pocket["mask"]] * 0.5                # sampler frame != PDB frame
        xh_pocket = torch.cat([shifted, pocket["one_hot"].float()], 1)
        seen["shift"] = com * 0.5
        return torch.cat([x, h], 1), xh_pocket, mask_p, pocket["mask"]

    model.ddpm.sample_given_pocket = stub
    out = model.generate_phars(pdb, n_samples, None, "A:901", torch.ones(n_samples, dtype=int) * n_pts, timesteps=7)
    assert seen["timesteps"] == 7 and 0 < seen["n_res"] <= 40
    assert sorted(out) == [f"Molecule_{k}" for k in range(1, n_pts + 1)]
    names = dataset_params["crossdock"]["phar_decoder"]
    total = 0
    for k in range(n_pts):
        for feat, coords in out[f"Molecule_{k + 1}"].items():
            assert feat in names
            total += len(coords)
    assert total == n_samples * n_pts
    # slot 1 of sample 0 is point 0 (type 0), moved back by the pocket-COM difference
    first = out["Molecule_1"][names[0]][0]
    assert torch.allclose(first, torch.tensor([0.0, 1.0, 2.0]) + seen["shift"][0], atol=1e-4)
    plain = cli.to_plain(out)
    json.dumps(plain)
    assert isinstance(plain["Molecule_1"][names[0]][0], list)
    with pytest.raises(AssertionError):
        model.generate_phars(pdb, 2, ["A:1"], "A:901")


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cli_end_to_end_writes_reference_json(tmp_path, monkeypatch, capsys):
    ckpt, pdb = tmp_path / "m.ckpt", str(tmp_path / "p.pdb")
    make_checkpoint(ckpt)
    write_synthetic_pdb(pdb, n_res=120)
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    out = cli.main([str(ckpt), "--pdbfile", pdb, "--ref_ligand", "A:901", "--n_samples", "6",
                    "--num_nodes_phar", "5", "--timesteps", "25"])
    data = json.load(open(tmp_path / cli.OUTPUT_NAME))
    assert sorted(data) == [f"Molecule_{k}" for k in range(1, 6)]
    pts = [c for m in data.values() for cs in m.values() for c in cs]
    assert len(pts) == 30 and all(len(c) == 3 and all(np.isfinite(c)) for c in pts)
    assert "Molecule_1" in capsys.readouterr().out                      # the reference prints the dict
    assert set(out) == set(data)


@pytest.mark.gpu
def test_generate_phars_matches_oracle_with_injected_noise(tmp_path):
    from cmd_gen_b200.schedule import gamma_table, step_table
    from cmd_gen_b200.synthetic import draw_noise
    from cmd_gen_b200.weights import init_weights
    from oracle import diffphar_oracle as orc
    ckpt, pdb = tmp_path / "m.ckpt", str(tmp_path / "p.pdb")
    cfg = make_checkpoint(ckpt, egnn_params=dict(n_layers=2))
    write_synthetic_pdb(pdb, n_res=80)
    model = PharPocketDDPM.load_from_checkpoint(ckpt, map_location="cuda", precision="fp32")
    n_samples, n_pts, steps = 3, 4, 12
    noise = draw_noise(steps + 2, n_samples * n_pts, 11, seed=3)
    it = iter(noise)
    model.ddpm.sample_gaussian = lambda size, device: next(it).to(device)
    out = model.generate_phars(pdb, n_samples, None, "A:901", torch.ones(n_samples, dtype=int) * n_pts, timesteps=steps)
    # oracle: same pocket through the CPU restatement + the glue of lightning_modules.py:458-504
    res = pdbio.pocket_from_ligand(pdbio.read_pdb(pdb), "A:901")
    xyz, types = pdbio.pocket_tensors(res, "CA", dataset_params["crossdock"]["aa_encoder"])
    x = torch.tensor(xyz).repeat(n_samples, 1)
    oh = torch.nn.functional.one_hot(torch.tensor(types), 20).repeat(n_samples, 1)
    mask = torch.repeat_interleave(torch.arange(n_samples), len(xyz))
    tab = step_table(gamma_table("polynomial_2", 500, 1e-5), 500, steps)
    ref_phar, ref_pocket, mp, _ = orc.sample_given_pocket(init_weights(cfg, 0), cfg, tab, x, oh, mask,
                                                          torch.full((n_samples,), n_pts), noise)
    com_b = torch.stack([x[mask == b].mean(0) for b in range(n_samples)])
    com_a = torch.stack([ref_pocket[mask == b, :3].mean(0) for b in range(n_samples)])
    ref_x = ref_phar[:, :3] + (com_b - com_a)[mp]
    ref_t = ref_phar[:, 3:].argmax(1)
    names = dataset_params["crossdock"]["phar_decoder"]
    scale = float(ref_x.abs().max())
    seen = {k: 0 for k in out}
    for b in range(n_samples):
        for k in range(n_pts):
            i = b * n_pts + k
            coords = out[f"Molecule_{k + 1}"][names[int(ref_t[i])]]
            d = min(float((c - ref_x[i]).abs().max()) for c in coords)
            assert d <= 2e-4 * max(scale, 1.0)
