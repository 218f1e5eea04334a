"""marl-mass_b200 — B200-native (sm_100a) batched merge environment + HSS/MASS CBF shields.

Hot path of hkbharath/MARL-MASS only: `MergeEnvLCMARL.reset()/step()` with the `safety_layer` shields inside,
as hand-written CUDA behind a C ABI (include/marl_mass_b200.h).  See DESIGN.md and INTEGRATION.md.

The directory is named `marl-mass_b200`; import it as `marl_mass_b200` (symlink at the repo root).
"""
from ._lib import MAXV, NA, NS, MMError, lib  # noqa: F401
from .env import (DEFAULT_CONFIG, MergeEnvBatched, MergeEnvLCHDV, MergeEnvLCMARL, MergeEnvMARL, make, make_mm_config,  # noqa: F401
                  set_step_variant, shield_qp)
from . import spawn  # noqa: F401
from . import evaluation, rollout  # noqa: F401

__all__ = ["MergeEnvBatched", "MergeEnvLCMARL", "MergeEnvLCHDV", "MergeEnvMARL", "make", "shield_qp", "set_step_variant", "spawn", "make_mm_config", "DEFAULT_CONFIG",
           "MAXV", "NA", "NS", "MMError", "lib", "evaluation", "rollout"]
