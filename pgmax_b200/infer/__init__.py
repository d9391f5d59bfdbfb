"""Inference (host mirror of the pgmax.infer sub-package).

``build_inferer(bp_state, backend)`` dispatches on a backend string exactly like
the reference (pgmax/infer/__init__.py:34-40): "bp" is the B200 hot path, "sdlp"
the smooth dual LP-MAP solver on the same plan and kernels.
"""

from pgmax_b200.infer import grad
from pgmax_b200.infer.bp import BeliefPropagation
from pgmax_b200.infer.bp import BP
from pgmax_b200.infer.bp import get_marginals
from pgmax_b200.infer.bp_state import BPArrays
from pgmax_b200.infer.bp_state import BPState
from pgmax_b200.infer.bp_state import Evidence
from pgmax_b200.infer.bp_state import FToVMessages
from pgmax_b200.infer.bp_state import LogPotentials
from pgmax_b200.infer.dual_lp import SDLP
from pgmax_b200.infer.dual_lp import SmoothDualLP
from pgmax_b200.infer.energy import compute_energy
from pgmax_b200.infer.inferer import decode_map_states
from pgmax_b200.infer.inferer import Inferer
from pgmax_b200.infer.inferer import InfererContext


def build_inferer(bp_state: BPState, backend: str) -> Inferer:
  """Inferer for ``bp_state``: backend "bp" -> BP(bp_state), "sdlp" -> SDLP(bp_state)."""
  if backend == "bp":
    return BP(bp_state)
  if backend == "sdlp":
    return SDLP(bp_state)
  raise NotImplementedError(f"Inferer {backend} is not supported.")
