"""Minimal PDB reader for pocket ingestion — what ``PharPocketDDPM.generate_phars`` needs from
Bio.PDB (lightning_modules.py:410-441, utils.py:96-119), without BioPython (absent in this image).

Behaviour mirrored from the reference's use of Bio.PDB:
  * only the first MODEL is read (``get_structure(...)[0]``);
  * a residue is addressed as ``<chain>:<resi>`` with a blank hetero flag and insertion code
    (``pdb_struct[chain][(' ', resi, ' ')]``), i.e. ATOM records only for ``--resi_list``;
  * alternate locations: one atom per name, the highest-occupancy altloc (Bio's DisorderedAtom default);
  * ``get_pocket_from_ligand``: every standard amino-acid residue with any atom closer than
    ``dist_cutoff`` (8 A, strict ``<``) to any atom of the reference residue ``<chain>:<resi>`` (which may be
    a HETATM ligand).  The reference's own "skip the ligand" test compares an int with a str and never
    fires; the ligand is excluded because it is not a standard amino acid — same outcome here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from .constants import THREE_TO_ONE


@dataclass
class Atom:
    name: str
    element: str
    coord: np.ndarray
    occupancy: float
    altloc: str


@dataclass
class Residue:
    chain: str
    resseq: int
    icode: str
    hetero: bool
    resname: str
    atoms: Dict[str, Atom] = field(default_factory=dict)      # insertion-ordered, one per atom name

    def coords(self) -> np.ndarray:
        return np.stack([a.coord for a in self.atoms.values()]).astype(np.float32)


def _element(line: str, name: str) -> str:
    el = line[76:78].strip() if len(line) >= 78 else ""
    if not el:                                   # fall back to the atom-name columns
        el = "".join(ch for ch in name if ch.isalpha())[:1]
    return el.upper()


def read_pdb(path: str) -> List[Residue]:
    """Residues of the first model in file order."""
    residues: Dict[Tuple[str, int, str, bool], Residue] = {}
    order: List[Residue] = []
    with open(path) as fh:
        for line in fh:
            rec = line[:6]
            if rec.startswith("ENDMDL"):
                break
            if rec not in ("ATOM  ", "HETATM"):
                continue
            name = line[12:16].strip()
            altloc = line[16]
            resname = line[17:20].strip()
            chain = line[21]
            resseq = int(line[22:26])
            icode = line[26]
            xyz = np.array([float(line[30:38]), float(line[38:46]), float(line[46:54])], dtype=np.float32)
            try:
                occ = float(line[54:60])
            except ValueError:
                occ = 1.0
            hetero = rec == "HETATM" and resname not in ("HOH", "WAT")
            water = rec == "HETATM" and resname in ("HOH", "WAT")
            key = (chain, resseq, icode, hetero or water)
            res = residues.get(key)
            if res is None:
                res = Residue(chain, resseq, icode, hetero or water, resname)
                residues[key] = res
                order.append(res)
            prev = res.atoms.get(name)
            if prev is None or (altloc != prev.altloc and occ > prev.occupancy):
                res.atoms[name] = Atom(name, _element(line, name), xyz, occ, altloc)   # dict keeps first-seen order
    return order


def select_residues(residues: List[Residue], pocket_ids: List[str]) -> List[Residue]:
    """``--resi_list A:123 A:124 ...`` (lightning_modules.py:413-417)."""
    index = {(r.chain, r.resseq): r for r in residues if not r.hetero and r.icode == " "}
    out = []
    for item in pocket_ids:
        chain, resi = item.split(":")
        key = (chain, int(resi))
        if key not in index:
            raise KeyError(f"residue {item} not found (blank hetero flag / insertion code expected)")
        out.append(index[key])
    return out


def pocket_from_ligand(residues: List[Residue], ligand_id: str, dist_cutoff: float = 8.0) -> List[Residue]:
    """utils.py:102-119."""
    chain, resi = ligand_id.split(":")
    lig = [r for r in residues if r.chain == chain and r.resseq == int(resi)]
    if len(lig) != 1:
        raise AssertionError(f"expected exactly one residue with id {ligand_id}, found {len(lig)}")
    lig_xyz = lig[0].coords().astype(np.float64)
    out = []
    for r in residues:
        # is_aa(resname, standard=True) is the reference's only filter (utils.py:115): a standard amino acid written as
        # HETATM (peptide ligands, some modified-chain exports) counts; its `residue.id[1] == resi` ligand skip compares
        # an int with a str and never fires, so a ligand that is itself a standard amino acid is kept too
        if r.resname not in THREE_TO_ONE:
            continue
        d = r.coords().astype(np.float64)[:, None, :] - lig_xyz[None, :, :]
        if np.sqrt((d * d).sum(-1)).min() < dist_cutoff:
            out.append(r)
    return out


def pocket_tensors(residues: List[Residue], representation: str, type_encoder: Dict[str, int]):
    """(coords [n,3] float32, type indices [n] int64) — lightning_modules.py:422-437."""
    if representation == "CA":
        xyz = np.stack([r.atoms["CA"].coord for r in residues]).astype(np.float32)
        types = np.array([type_encoder[THREE_TO_ONE[r.resname]] for r in residues], dtype=np.int64)
        return xyz, types
    atoms = [a for r in residues for a in r.atoms.values()
             if (a.element.capitalize() in type_encoder or a.element != "H")]
    xyz = np.stack([a.coord for a in atoms]).astype(np.float32)
    types = np.array([type_encoder[a.element.capitalize()] for a in atoms], dtype=np.int64)
    return xyz, types
