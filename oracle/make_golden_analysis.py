"""Mints tests/golden/analysis_hist.npz from the UNMODIFIED helper functions of the reference's DiffPhar/test.py
(get_type_histograms :34-41, convert_pharmacophore_to_one_hot :43-52).  test.py is a script (its evaluation loop sits
under ``__main__`` and its imports need rdkit / pytorch_lightning), so the two functions are pulled out of its source
with ``ast`` and executed as they are — nothing is copied into this repo.  Run in the build container only:

    python oracle/make_golden_analysis.py
"""
import ast
import os

import numpy as np

REF = "/root/reference/DiffPhar/test.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "analysis_hist.npz")

src = open(REF).read()
tree = ast.parse(src)
wanted = {"get_type_histograms", "convert_pharmacophore_to_one_hot"}
ns = {"np": np, "num_phar_classes": 8}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in wanted:
        exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
phar_encoder = {'Aromatic': 0, 'Hydrophobe': 1, 'PosIonizable': 2, 'NegIonizable': 3, 'Acceptor': 4, 'Donor': 5,
                'LumpedHydrophobe': 6, 'others': 7}      # constants.py:99
rng = np.random.default_rng(7)
cases = {}
for k in range(6):
    idx = rng.integers(0, 8, size=rng.integers(1, 30))
    one_hot = ns["convert_pharmacophore_to_one_hot"](list(idx))
    hist = ns["get_type_histograms"](one_hot, phar_encoder)
    cases[f"idx_{k}"] = idx.astype(np.int64)
    cases[f"onehot_{k}"] = one_hot
    cases[f"hist_{k}"] = np.array([hist[n] for n in phar_encoder], dtype=np.int64)
np.savez(OUT, n_cases=6, **cases)
print("wrote", OUT)
