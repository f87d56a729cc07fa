"""Evaluation statistics of generated point clouds (reference DiffPhar/test.py:157-197): the oracle restatement against
hand-computed values (CPU), and the GPU reduction behind cmd_gen_b200.analysis.phar_statistics against the oracle."""
import numpy as np
import pytest
import torch

from oracle import diffphar_oracle as orc

gpu = pytest.mark.gpu


def _random_result(seed, n_mol, max_types=4, max_pts=40):
    rng = np.random.default_rng(seed)
    names = ['Aromatic', 'Hydrophobe', 'PosIonizable', 'NegIonizable', 'Acceptor', 'Donor', 'LumpedHydrophobe', 'others', 'ZnBinder']
    out = {}
    for m in range(n_mol):
        mol = {}
        for name in rng.choice(names, size=rng.integers(1, max_types + 1), replace=False):
            mol[str(name)] = (rng.normal(size=(rng.integers(1, max_pts + 1), 3)) * 6.0 + [12.0, -7.5, 30.25]).tolist()
        out[f"Molecule_{m}"] = mol
    return out


def test_oracle_statistics_known_answers():
    res = {"Molecule_0": {"Aromatic": [[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]], "Donor": [[0.0, 0.0, 12.0]]},
           "Molecule_1": {"Hydrophobe": [[1.0, 1.0, 1.0]]},
           "Molecule_2": {"Unknown": [[2.0, 0.0, 0.0], [0.0, 2.0, 0.0]]}}
    st = orc.phar_statistics(res, [1.0, 1.0, 1.0])
    assert st["num_gen"].tolist() == [3, 1, 2]
    assert np.allclose(st["max_phar_distances"], [13.0, 0.0, np.sqrt(8.0)])
    assert np.allclose(st["com_distances"], [np.linalg.norm([0.0, 1.0 / 3.0, 3.0]), 0.0, 1.0])
    # one histogram entry per type NAME of a molecule; the reference's one-hot sets column value - 1, so index 0
    # ('Aromatic') is read back as the last class and every other type as its predecessor (test.py:43-52, 34-41)
    assert st["gen_phar_hist"] == {'Aromatic': 1, 'Hydrophobe': 0, 'PosIonizable': 0, 'NegIonizable': 0, 'Acceptor': 1,
                                   'Donor': 0, 'LumpedHydrophobe': 1, 'others': 1}


def test_oracle_statistics_empty():
    st = orc.phar_statistics({}, [0.0, 0.0, 0.0])
    assert st["num_gen"].size == 0 and sum(st["gen_phar_hist"].values()) == 0


@gpu
@pytest.mark.parametrize("seed,n_mol,max_pts", [(0, 12, 40), (1, 3, 700), (2, 200, 9)])
def test_gpu_statistics_match_oracle(seed, n_mol, max_pts):
    from cmd_gen_b200.analysis import phar_statistics
    res = _random_result(seed, n_mol, max_pts=max_pts)
    ref_c = [11.0, -6.0, 29.0]
    want = orc.phar_statistics(res, ref_c)
    got = phar_statistics(res, ref_c)
    assert np.array_equal(got["num_gen"], want["num_gen"])
    assert got["gen_phar_hist"] == want["gen_phar_hist"]
    # float64 on both sides; only the order of the centroid's additions differs (numpy sums pairwise)
    assert np.allclose(got["com_distances"], want["com_distances"], rtol=1e-12, atol=1e-12)
    assert np.allclose(got["max_phar_distances"], want["max_phar_distances"], rtol=1e-13, atol=0.0)


@gpu
def test_gpu_statistics_empty_and_single_point():
    from cmd_gen_b200.analysis import phar_statistics
    assert phar_statistics({}, [0.0, 0.0, 0.0])["num_gen"].size == 0
    st = phar_statistics({"Molecule_0": {"Donor": [[1.0, 2.0, 3.0]]}}, [1.0, 2.0, 7.0])
    assert st["num_gen"].tolist() == [1] and st["max_phar_distances"].tolist() == [0.0] and st["com_distances"].tolist() == [4.0]


def test_oracle_type_histogram_matches_reference_helpers():
    """tests/golden/analysis_hist.npz holds outputs of the reference's own convert_pharmacophore_to_one_hot /
    get_type_histograms (oracle/make_golden_analysis.py); one molecule per index so that the oracle's per-molecule
    type list reproduces the index list."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "analysis_hist.npz"))
    names = ['Aromatic', 'Hydrophobe', 'PosIonizable', 'NegIonizable', 'Acceptor', 'Donor', 'LumpedHydrophobe', 'others']
    from cmd_gen_b200.analysis import _one_hot_columns
    for k in range(int(g["n_cases"])):
        idx = g[f"idx_{k}"]
        res = {f"Molecule_{i}": {names[v]: [[0.0, 0.0, float(i)]]} for i, v in enumerate(idx)}
        st = orc.phar_statistics(res, [0.0, 0.0, 0.0])
        assert [st["gen_phar_hist"][n] for n in names] == g[f"hist_{k}"].tolist()
        assert _one_hot_columns(idx.tolist()) == g[f"onehot_{k}"].argmax(1).tolist()     # the product's host logic
