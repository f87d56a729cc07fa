"""Type vocabularies of the DiffPhar datasets (DiffPhar/constants.py:95-121) — the integer <-> name
tables the sampler's inputs (pocket one-hot) and outputs (pharmacophore type names) are defined by.
Only what the sampling path reads is restated; histograms used for training-time statistics are not.
"""
import torch

FLOAT_TYPE = torch.float32          # constants.py:8
INT_TYPE = torch.int64              # constants.py:9

PHAR_TYPES = ["Aromatic", "Hydrophobe", "PosIonizable", "NegIonizable", "Acceptor", "Donor",
              "LumpedHydrophobe", "others"]
AMINO_ACIDS = list("ACDEFGHIKLMNPQRSTVWY")
ELEMENTS_FULL = ["C", "N", "O", "S", "B", "Br", "Cl", "P", "I", "F", "others"]
ELEMENTS = ELEMENTS_FULL[:-1]


def _enc(names):
    return {n: i for i, n in enumerate(names)}


dataset_params = {
    # full-atom pockets: 11 element classes incl. 'others' (constants.py:97-98)
    "crossdock_full": {
        "atom_encoder": _enc(ELEMENTS_FULL), "atom_decoder": list(ELEMENTS_FULL),
        "phar_encoder": _enc(PHAR_TYPES), "phar_decoder": list(PHAR_TYPES),
        "aa_encoder": _enc(ELEMENTS_FULL), "aa_decoder": list(ELEMENTS_FULL),
    },
    # C-alpha pockets: 20 amino-acid classes (constants.py:115-116)
    "crossdock": {
        "atom_encoder": _enc(ELEMENTS), "atom_decoder": list(ELEMENTS),
        "phar_encoder": _enc(PHAR_TYPES), "phar_decoder": list(PHAR_TYPES),
        "aa_encoder": _enc(AMINO_ACIDS), "aa_decoder": list(AMINO_ACIDS),
    },
}

THREE_TO_ONE = {
    "ALA": "A", "CYS": "C", "ASP": "D", "GLU": "E", "PHE": "F", "GLY": "G", "HIS": "H", "ILE": "I", "LYS": "K",
    "LEU": "L", "MET": "M", "ASN": "N", "PRO": "P", "GLN": "Q", "ARG": "R", "SER": "S", "THR": "T", "VAL": "V",
    "TRP": "W", "TYR": "Y",
}
