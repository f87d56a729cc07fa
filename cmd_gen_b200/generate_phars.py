"""Drop-in for DiffPhar/generate_phars.py: same positional argument and flags (generate_phars.py:11-25),
same output file — ``phar_to_coords_no_tensor_PI3K_dul.json`` in the CURRENT directory, name hard-coded by
the reference (``--outdir`` is parsed and unused there too) — and the same final ``print`` of the dict.

    python -m cmd_gen_b200.generate_phars <checkpoint> --pdbfile P.pdb --ref_ligand A:1101 \\
        --n_samples 10 --num_nodes_phar 10 [--timesteps 100] [--precision f16fast|f16|bf16|fp32]

``--precision`` is the only addition (arithmetic mode of the CUDA kernels; default f16fast: f16 tensor-core tiles, packed-f16 first layer).
"""
import argparse
import json
from pathlib import Path

import numpy as np
import torch

from .lightning_modules import PharPocketDDPM

OUTPUT_NAME = "phar_to_coords_no_tensor_PI3K_dul.json"


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('checkpoint', type=Path)
    parser.add_argument('--pdbfile', type=str)
    parser.add_argument('--resi_list', type=str, nargs='+', default=None)
    parser.add_argument('--ref_ligand', type=str, default=None)
    parser.add_argument('--outdir', type=Path)
    parser.add_argument('--n_samples', type=int, default=20)
    parser.add_argument('--num_nodes_phar', type=int, default=3)
    parser.add_argument('--all_frags', action='store_true')
    parser.add_argument('--sanitize', action='store_true')
    parser.add_argument('--relax', action='store_true')
    parser.add_argument('--resamplings', type=int, default=10)
    parser.add_argument('--jump_length', type=int, default=1)
    parser.add_argument('--timesteps', type=int, default=None)
    parser.add_argument('--precision', type=str, default='f16fast', choices=['fp32', 'bf16', 'f16', 'f16fast'])
    return parser


def to_plain(phar_to_coords):
    return {mol: {feat: [c.tolist() for c in coords] for feat, coords in feats.items()}
            for mol, feats in phar_to_coords.items()}


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("generate_phars: no CUDA device — this implementation has no CPU fallback")
    device = 'cuda'
    model = PharPocketDDPM.load_from_checkpoint(args.checkpoint, map_location=device, precision=args.precision)
    model = model.to(device)
    if args.num_nodes_phar is not None:
        num_nodes_phar = torch.ones(args.n_samples, dtype=int) * args.num_nodes_phar
    else:
        num_nodes_phar = None
    phar_to_coords = model.generate_phars(
        args.pdbfile, args.n_samples, args.resi_list, args.ref_ligand, num_nodes_phar, args.sanitize,
        largest_frag=not args.all_frags, relax_iter=(200 if args.relax else 0),
        resamplings=args.resamplings, jump_length=args.jump_length, timesteps=args.timesteps)
    with open(Path(OUTPUT_NAME), 'w') as f:
        json.dump(to_plain(phar_to_coords), f,
                  default=lambda x: x.tolist() if isinstance(x, np.ndarray) else x)
    print(phar_to_coords)
    return phar_to_coords


if __name__ == "__main__":
    main()
