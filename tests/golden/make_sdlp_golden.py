#!/usr/bin/env python
"""LP-MAP optima of the reference's dual-LP test models (tests/lp/test_dual_lp.py), solved with an
independent LP solver (SciPy HiGHS on the program of pgmax/utils/primal_lp.py:31-178, restated in
oracle/primal_lp_oracle.py) -> tests/golden/sdlp_lp.npz.

The reference's tests assert `primal_upper_bound == cvxpy_lp_objval` (rtol 5e-3) on these models;
cvxpy is not in the image, so the LP optimum is generated here once and committed.  Run from the
repo root:  python tests/golden/make_sdlp_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
  if p not in sys.path:
    sys.path.insert(0, p)

import models  # noqa: E402
from oracle import primal_lp_oracle  # noqa: E402


def cases():
  """(name, fg, evidence_updates) of every model / seed the SDLP tests use."""
  for seed in (0, 1):
    fg, variables = models.sdlp_ising_model(seed=seed)
    yield f"ising_{seed}", fg, {variables: np.random.RandomState(seed).gumbel(size=(4, 4, 3))}
  for seed in range(3):
    fg, top, bottom, evidence = models.sdlp_line_model(seed=seed)
    yield f"line_{seed}", fg, evidence
  for seed in (0, 1):
    fg, matrix, all_ones, evidence, truth = models.sdlp_and_model(seed=seed)
    yield f"and_{seed}", fg, evidence
  for seed in (0, 1):
    fg, variables = models.sdlp_pool_model()
    updates = np.random.RandomState(seed).gumbel(size=(variables.shape[0], 2))
    updates[0, 1] = 1_000
    yield f"pool_{seed}", fg, {variables: updates}


def main():
  out = {}
  for name, fg, evidence in cases():
    solution, objval = primal_lp_oracle.primal_lp_solver(fg, evidence)
    out[f"{name}_objval"] = np.float64(objval)
    out[f"{name}_solution"] = solution.astype(np.float64)
    print(f"{name:10s} LP optimum {objval:.9g}  integral: {bool(np.all(np.minimum(solution, 1 - solution) < 1e-6))}")
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sdlp_lp.npz"), **out)


if __name__ == "__main__":
  main()
