"""Differential property tests for the oracle's OR / AND / Pool updates (no GPU).

The reference pins its closed-form logical updates by building every model twice
— with OR/AND/Pool factors and with the equivalent EnumFactors — and comparing
beliefs after 5 iterations (tests/factor/test_or.py:30-290, test_and.py,
test_pool.py:29-281).  No stored JAX outputs exist for these paths, so the same
property is what pins the oracle's restatement of pgmax/factor/logical.py:561-779
and pgmax/factor/pool.py:328-474 (its Enum path is pinned by the golden vectors).
"""

import numpy as np
import pytest

import models
from oracle import bp_oracle
from pgmax_b200 import infer

# (temperature, atol) per seed % 4, as tests/factor/test_or.py:67-87.
TEMPS = [(0.0, 1e-5), (0.001, 5e-3), (0.3, 5e-3), (0.8, 1e-5)]


@pytest.mark.parametrize("kind", ["or", "and", "pool"])
@pytest.mark.parametrize("seed", range(8))
def test_logical_equals_equivalent_enum(kind, seed):
  temperature, atol = TEMPS[seed % 4]
  if kind == "pool":
    atol = max(atol, 1e-5)
  data = models.logical_pair(kind, seed)
  beliefs = []
  for entry in data["graphs"]:
    bp = infer.BP(entry[0].bp_state, temperature=temperature)
    arrays = models.init_logical(bp, entry, data)
    graph = bp_oracle.graph_from_context(bp.context)
    msgs, _ = bp_oracle.run_bp(graph, arrays.log_potentials, arrays.ftov_msgs, arrays.evidence,
                               num_iters=5, damping=0.5, temperature=temperature)
    beliefs.append(bp_oracle.flat_beliefs(graph, msgs, arrays.evidence))
  np.testing.assert_allclose(beliefs[0], beliefs[1], atol=atol, rtol=0)
