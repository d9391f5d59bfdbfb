"""The C-ABI library loads on a CPU-only box, exports every symbol include/pgx.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

import models
from pgmax_b200 import _native, infer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
  if not os.path.exists(_native.LIB_PATH):
    _native.build()
  return _native.load()


def test_header_symbols_are_exported(lib):
  header = open(os.path.join(ROOT, "include", "pgx.h")).read()
  declared = set(re.findall(r"\b(pgx_[a-z_]+)\s*\(", header))
  assert declared == set(_native.EXPORTED_SYMBOLS)
  raw = ctypes.CDLL(_native.LIB_PATH)
  for name in declared:
    assert hasattr(raw, name), name


def test_build_info_and_error_string(lib):
  info = lib.pgx_build_info().decode()
  assert "sm_100a" in info
  assert isinstance(lib.pgx_last_error(), bytes)


def test_struct_layouts_match_header(tmp_path):
  """ctypes mirrors vs the sizes / offsets gcc gives the structs of include/pgx.h."""
  import subprocess
  src = tmp_path / "sizes.c"
  src.write_text(
      '#include <stdio.h>\n#include <stddef.h>\n#include "pgx.h"\n'
      'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pgx_enum_block),'
      'sizeof(pgx_logical_desc), sizeof(pgx_graph_desc), sizeof(pgx_plan_info),'
      'offsetof(pgx_graph_desc, or_factors), offsetof(pgx_graph_desc, pool_factors),'
      'offsetof(pgx_enum_block, first_edge));return 0;}\n')
  exe = tmp_path / "sizes"
  subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
  got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
  want = [ctypes.sizeof(_native.EnumBlockC), ctypes.sizeof(_native.LogicalDescC),
          ctypes.sizeof(_native.GraphDescC), ctypes.sizeof(_native.PlanInfoC),
          _native.GraphDescC.or_factors.offset, _native.GraphDescC.pool_factors.offset,
          _native.EnumBlockC.first_edge.offset]
  assert got == want


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the CPU-only behaviour")
def test_no_cpu_fallback(lib):
  fg, variables, evidence = models.ising_model(n=4)
  bp = infer.BP(fg.bp_state, temperature=0.0)
  arrays = bp.init(evidence_updates={variables: evidence})  # host-side state handling works
  assert arrays.evidence.shape == (32,)
  with pytest.raises(_native.PgxError) as err:
    bp.run(arrays, num_iters=1)
  assert err.value.code == _native.PGX_ERR_NO_DEVICE
  with pytest.raises(_native.PgxError):
    bp.get_beliefs(arrays)
  # null handles are rejected, not dereferenced
  assert lib.pgx_bp_run(None, None, 1, None, 0, None, 0, None, 0, None, None, 1, 0.5, 0.0) == _native.PGX_ERR_INVALID
  assert b"null plan" in lib.pgx_last_error()


def test_flag_constants_match_the_header():
  """Every PGX_PATH_* / PGX_RUN_* / PGX_STRIP_* bit of include/pgx.h has the same value in the
  Python mirror (pgmax_b200/_native.py: Plan.PATH_* / Plan.RUN_* / Strip flags), and no two path
  bits collide."""
  import re
  from pgmax_b200 import _native
  text = open(os.path.join(ROOT, "include", "pgx.h")).read()
  defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(PGX_(?:PATH|RUN)_[A-Z0-9_]+)\s+(\d+)u", text)}
  paths = {k: v for k, v in defines.items() if k.startswith("PGX_PATH_")}
  assert len(paths) >= 20 and len(set(paths.values())) == len(paths)
  assert all(v & (v - 1) == 0 for v in paths.values())          # single bits
  for name, value in defines.items():
    attr = name[len("PGX_"):]
    assert getattr(_native.Plan, attr) == value, name
