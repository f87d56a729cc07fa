"""Evaluation statistics of generated pharmacophore point clouds — the host-side mirror of the loop in the
reference's ``DiffPhar/test.py:157-197``: for every "molecule" of ``generate_phars``' output dictionary the number of
points, the distance of their centroid to the reference ligand's centroid and the largest pairwise distance, plus
the histogram of the pharmacophore types present.  The three per-molecule reductions run on the GPU through the
C-ABI call ``dp_pointcloud_stats`` (one CTA per molecule, float64 like the reference's numpy arrays); there is no
CPU fallback — the checker for it lives in ``oracle/diffphar_oracle.py::phar_statistics``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib
from .constants import dataset_params

NUM_PHAR_CLASSES = 8                                    # test.py:29


def _type_rows(molecule: Dict[str, list], phar_dict: Dict[str, int]) -> List[int]:
    """One index per distinct type NAME of the molecule (test.py:171-173; unknown names map to 7 and, like the
    reference's ``setdefault``, stay in the dictionary for later molecules)."""
    return [phar_dict.setdefault(name, 7) for name in molecule.keys()]


def _one_hot_columns(indices: Sequence[int]) -> List[int]:
    """``convert_pharmacophore_to_one_hot`` (test.py:43-52) sets column ``value - 1`` — index 0 lands in the LAST
    column — and ``get_type_histograms`` (test.py:34-41) reads the argmax back: the column it will see."""
    cols = []
    for v in indices:
        if not 0 <= v <= NUM_PHAR_CLASSES:
            raise ValueError("Pharmacophore value is out of range.")
        cols.append((v - 1) % NUM_PHAR_CLASSES)
    return cols


def phar_statistics(phar_to_coords: Dict[str, Dict[str, list]], ref_centroid, device="cuda:0",
                    dataset: str = "crossdock_full") -> Dict[str, object]:
    """Statistics of one ``generate_phars`` result against the reference ligand's centroid.

    Returns ``{"com_distances", "max_phar_distances", "num_gen"}`` (numpy arrays, one entry per molecule in dictionary
    order, as test.py:194-196 appends them) and ``"gen_phar_hist"`` (type name -> count, test.py:202)."""
    phar_dict = dict(dataset_params[dataset]["phar_encoder"])
    names = list(dataset_params[dataset]["phar_encoder"].keys())
    coords, offsets, columns = [], [0], []
    for molecule in phar_to_coords.values():
        columns.extend(_one_hot_columns(_type_rows(molecule, phar_dict)))
        for pts in molecule.values():                   # all_phar_coords.extend(phar_coords), test.py:176-177
            coords.extend(pts)
        offsets.append(len(coords))
    n_groups = len(offsets) - 1
    hist = {k: 0 for k in names}
    for c in columns:
        hist[names[c]] += 1
    if n_groups == 0:
        return {"com_distances": np.zeros(0), "max_phar_distances": np.zeros(0), "num_gen": np.zeros(0, dtype=np.int64),
                "gen_phar_hist": hist}
    lib = _lib.load_library()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.DiffPharError("phar_statistics runs on a CUDA device (no CPU fallback)")
    xyz = torch.tensor(np.asarray(coords, dtype=np.float64).reshape(-1, 3), device=dev)
    off = torch.tensor(offsets, dtype=torch.int32, device=dev)
    out = torch.empty((n_groups, 3), dtype=torch.float64, device=dev)
    ref = (C.c_double * 3)(*[float(v) for v in np.asarray(ref_centroid, dtype=np.float64).reshape(3)])
    with torch.cuda.device(dev):
        _lib._check(lib.dp_pointcloud_stats(xyz.data_ptr(), off.data_ptr(), n_groups, ref, out.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    res = out.cpu().numpy()
    return {"com_distances": res[:, 1].copy(), "max_phar_distances": res[:, 2].copy(),
            "num_gen": res[:, 0].astype(np.int64), "gen_phar_hist": hist}
