"""Generates the golden fixtures under tests/golden/ from the reference tree.

Run in the build container (the reference is mounted read-only there):

    python tests/golden/make_golden.py [/root/reference]

JAX is not installable in this image, so nothing is *executed* from the
reference; the script lifts the known answers the reference's own tests and
benchmark results hold for the BP path:

  e2e_sanity.npz   the 84 golden messages (``true_final_msgs_output``) and the
                   12 golden MAP states of tests/test_pgmax.py:63-152,252-265
                   (100 iterations, T=0, damping 0.5 on the 3x3 "cut" model).
  rbm24.npz        benchmark/precomputed_results/n_units_24_rbm_idx_{0..49}:
                   weights (W, bh, bv), and the reference's decoded hidden /
                   visible states + energies after 20 and 200 iterations of
                   max-product BP, CPU and GPU back-ends, batch size 1
                   (harness benchmark/rbm_lib.py:135-214).  All 24-unit files are
                   identical between the reference's CPU and GPU back-ends.
  rbm_large.npz    the same for the first 4 RBMs of n_units 40, 100 and 200 after 20 iterations
                   (the short horizon on which the reference's CPU and GPU back-ends agree,
                   SURVEY.md §8c), batch size 1.
"""

import ast
import os
import pickle
import sys

import joblib
import numpy as np
from joblib import numpy_pickle

HERE = os.path.dirname(os.path.abspath(__file__))


def lift_e2e_sanity(ref_root: str) -> None:
  path = os.path.join(ref_root, "tests", "test_pgmax.py")
  tree = ast.parse(open(path).read())
  msgs, map_states = None, None
  for node in ast.walk(tree):
    if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
      name = node.targets[0].id
      if name == "true_final_msgs_output":
        # jax.device_put(jnp.array([...]))
        call = node.value
        while isinstance(call, ast.Call):
          call = call.args[0]
        msgs = np.array(ast.literal_eval(call), dtype=np.float64)
      elif name == "true_map_state_output":
        # {(grid_vars, (0, 0, 0)): 2, ...}: keep (group name, index, state)
        map_states = [
            (k.elts[0].id, ast.literal_eval(k.elts[1]), ast.literal_eval(v))
            for k, v in zip(node.value.keys, node.value.values)
        ]
  assert msgs is not None and msgs.shape == (84,), "golden messages not found"
  assert map_states is not None and len(map_states) == 12
  np.savez(
      os.path.join(HERE, "e2e_sanity.npz"),
      true_final_msgs_output=msgs,
      map_groups=np.array([m[0] for m in map_states]),
      map_indices=np.array([m[1] for m in map_states], dtype=np.int64),
      map_states=np.array([m[2] for m in map_states], dtype=np.int64),
  )
  print("e2e_sanity.npz:", msgs.shape, len(map_states), "MAP states")


class _NoJaxUnpickler(numpy_pickle.NumpyUnpickler):
  """Reads joblib files holding pickled jax.Arrays without importing jax."""

  def find_class(self, module, name):
    if module.startswith("jax") and name == "_reconstruct_array":

      def rebuild(fun, args, arr_state, aval_state):
        del aval_state
        arr = fun(*args)
        arr.__setstate__(arr_state)
        return arr

      return rebuild
    return super().find_class(module, name)


def _load(path):
  with open(path, "rb") as f:
    return _NoJaxUnpickler(path, f, ensure_native_byte_order=False).load()


def lift_rbm24(ref_root: str) -> None:
  folder = os.path.join(ref_root, "benchmark", "precomputed_results")
  out = {}
  W, bh, bv = [], [], []
  for idx in range(50):
    w = _load(os.path.join(folder, f"n_units_24_rbm_idx_{idx}_weights.joblib"))
    W.append(np.asarray(w[0] if not isinstance(w, dict) else w["W"]))
    bh.append(np.asarray(w[1] if not isinstance(w, dict) else w["bh"]))
    bv.append(np.asarray(w[2] if not isinstance(w, dict) else w["bv"]))
  out["W"], out["bh"], out["bv"] = np.stack(W), np.stack(bh), np.stack(bv)
  for backend in ("cpu", "gpu"):
    for iters in (20, 200):
      hid, vis, en = [], [], []
      for idx in range(50):
        r = _load(os.path.join(
            folder,
            f"n_units_24_rbm_idx_{idx}_pgmax_{backend}_num_iters_{iters}_batch_size_1.joblib"))
        hid.append(np.asarray(r["hidden"]).astype(np.int64))
        vis.append(np.asarray(r["visible"]).astype(np.int64))
        en.append(float(np.asarray(r["energy"]).reshape(-1)[0]))
      out[f"hidden_{backend}_{iters}"] = np.stack(hid)
      out[f"visible_{backend}_{iters}"] = np.stack(vis)
      out[f"energy_{backend}_{iters}"] = np.array(en)
  np.savez_compressed(os.path.join(HERE, "rbm24.npz"), **out)
  print("rbm24.npz:", {k: v.shape for k, v in out.items()})


def lift_rbm_large(ref_root: str, sizes=(40, 100, 200), count: int = 4) -> None:
  folder = os.path.join(ref_root, "benchmark", "precomputed_results")
  out = {}
  for n in sizes:
    for idx in range(count):
      w = _load(os.path.join(folder, f"n_units_{n}_rbm_idx_{idx}_weights.joblib"))
      for key, arr in zip(("W", "bh", "bv"), (w[0], w[1], w[2]) if not isinstance(w, dict) else (w["W"], w["bh"], w["bv"])):
        out[f"{key}_{n}_{idx}"] = np.asarray(arr)
      for backend in ("cpu", "gpu"):
        r = _load(os.path.join(folder, f"n_units_{n}_rbm_idx_{idx}_pgmax_{backend}_num_iters_20_batch_size_1.joblib"))
        out[f"hidden_{backend}_{n}_{idx}"] = np.asarray(r["hidden"]).astype(np.int64).reshape(-1)
        out[f"visible_{backend}_{n}_{idx}"] = np.asarray(r["visible"]).astype(np.int64).reshape(-1)
        out[f"energy_{backend}_{n}_{idx}"] = np.array(float(np.asarray(r["energy"]).reshape(-1)[0]))
  np.savez_compressed(os.path.join(HERE, "rbm_large.npz"), **out)
  print("rbm_large.npz:", len(out), "arrays")


if __name__ == "__main__":
  root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
  lift_e2e_sanity(root)
  lift_rbm24(root)
  lift_rbm_large(root)
