"""cmd_gen_b200 — B200-native drop-in for DiffPhar's pocket-conditioned
reverse-diffusion sampler (the hot path BASELINE.json names).  See DESIGN.md."""
from .config import DynamicsConfig, weight_spec, weight_count, PRECISION_MODES  # noqa: F401

__all__ = ["DynamicsConfig", "weight_spec", "weight_count", "PRECISION_MODES"]
