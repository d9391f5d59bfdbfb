"""Numerics policy constants shared by the host mirror, the oracle and the kernels.

The three values are the reference's (pgmax/utils/__init__.py:26-37) and are
also compiled into csrc/pgx_kernels.cuh; tests/test_abi.py checks both agree.
"""

import functools

# Lower clip of a normalised message (reference pgmax/utils/__init__.py:26).
MSG_NEG_INF = -1e32
# |log potential| clip applied inside run() only (pgmax/utils/__init__.py:32).
LOG_POTENTIAL_MAX_ABS = 1e6
# "no valid configuration" marker (pgmax/utils/__init__.py:37).
NEG_INF = -float("inf")


def cached_property(func):
  """Memoised read-only property (same contract as pgmax/utils/__init__.py:40)."""
  return property(functools.lru_cache(None)(func))
