"""Model builders shared by the tests (graphs follow the reference's tests / notebooks)."""

import itertools

import numpy as np
from scipy.ndimage import gaussian_filter

from pgmax_b200 import factor, fgraph, fgroup, infer, vgroup


# 29 valid configurations of the 4-variable "cut" factors and the suppression
# configurations of the reference's e2e test (tests/test_pgmax.py:176-213, 37-61).
CUT_CONFIGS = np.array([
    [0, 0, 0, 0], [1, 0, 1, 0], [2, 0, 2, 0], [0, 0, 1, 1], [0, 0, 2, 2], [2, 0, 0, 1],
    [1, 0, 0, 2], [1, 0, 1, 1], [2, 0, 2, 1], [1, 0, 1, 2], [2, 0, 2, 2], [0, 1, 0, 1],
    [1, 1, 0, 0], [0, 1, 2, 0], [1, 1, 0, 1], [2, 1, 0, 1], [0, 1, 1, 1], [0, 1, 2, 1],
    [1, 1, 1, 0], [2, 1, 2, 0], [0, 2, 0, 2], [2, 2, 0, 0], [0, 2, 1, 0], [2, 2, 0, 2],
    [1, 2, 0, 2], [0, 2, 2, 2], [0, 2, 1, 2], [2, 2, 2, 0], [1, 2, 1, 0],
])


def suppression_configs(diameter: int) -> np.ndarray:
  rows = [[0] * diameter]
  for idx in range(diameter):
    for val in (1, 2):
      row = [0] * diameter
      row[idx] = val
      rows.append(row)
  return np.array(rows)


def cut_model(im_size: int = 3, seed: int = 23):
  """The 3x3 depth-scene cut model of tests/test_pgmax.py:35-421 (evidence from
  default_rng(23) logistic noise, drawn in the same order).  Returns
  (fg, bp_state, grid_vars, additional_vars)."""
  rng = np.random.default_rng(seed)
  depth = 5.0 * np.ones((im_size, im_size))
  depth[np.tril_indices(im_size, 0)] = 1.0
  depth = gaussian_filter(depth, sigma=0.5)
  M = N = im_size
  dh = depth[:-1] - depth[1:]
  dv = depth[:, :-1] - depth[:, 1:]
  cuts = np.zeros((2, M, N), dtype=np.int32)
  cuts[0, :-1] = np.where(dh < 0, 1, np.where(dh > 0, 2, 0))
  cuts[1, :, :-1] = np.where(dv < 0, 1, np.where(dv > 0, 2, 0))

  grid_vars = vgroup.NDVarArray(shape=(2, M - 1, N - 1), num_states=3)
  extra_names = tuple(
      [(0, row, N - 1) for row in range(M - 1)] + [(1, M - 1, col) for col in range(N - 1)]
  )
  additional_vars = vgroup.VarDict(variable_names=extra_names, num_states=3)

  grid_ev = np.zeros((2, M - 1, N - 1, 3))
  extra_ev = {}
  for i, row, col in itertools.product(range(2), range(M), range(N)):
    ev = np.zeros(3)
    ev[cuts[i, row, col]] = 2.0
    ev = ev - ev[0]
    ev[1:] += 0.1 * rng.logistic(size=2)  # noise is drawn for EVERY (i, row, col)
    if row < M - 1 and col < N - 1:
      grid_ev[i, row, col] = ev
    elif (i, row, col) in extra_names:
      extra_ev[i, row, col] = ev

  def var(i, row, col):
    if row < M - 1 and col < N - 1:
      return grid_vars[i, row, col]
    return additional_vars[i, row, col]

  fg = fgraph.FactorGraph(variable_groups=[grid_vars, additional_vars])
  for row, col in itertools.product(range(M - 1), range(N - 1)):
    fg.add_factors(
        factor.EnumFactor(
            variables=[var(0, row, col), var(1, row, col), var(0, row, col + 1), var(1, row + 1, col)],
            factor_configs=CUT_CONFIGS,
            log_potentials=np.zeros(CUT_CONFIGS.shape[0]),
        )
    )
  diameter = 2
  supp = suppression_configs(diameter)
  vert = [
      [var(0, r, col) for r in range(start, start + diameter)]
      for col in range(N)
      for start in range(M - diameter)
  ]
  horz = [
      [var(1, row, c) for c in range(start, start + diameter)]
      for row in range(M)
      for start in range(N - diameter)
  ]
  fg.add_factors(fgroup.EnumFactorGroup(variables_for_factors=vert, factor_configs=supp))
  fg.add_factors(
      fgroup.EnumFactorGroup(
          variables_for_factors=horz, factor_configs=supp, log_potentials=np.zeros(supp.shape[0])
      )
  )
  bp_state = fg.bp_state
  bp_state.evidence[grid_vars] = grid_ev
  bp_state.evidence[additional_vars] = extra_ev
  return fg, bp_state, grid_vars, additional_vars


def ising_model(n: int = 50, coupling: float = 0.8, seed: int = 0, batch=None):
  """n x n binary torus, pairwise factors [(i,j),(i+1,j)] and [(i,j),(i,j+1)]
  (examples/ising_model.ipynb cells 8-12; tests/test_examples.py:26-41).
  Returns (fg, variables, evidence_array)."""
  variables = vgroup.NDVarArray(num_states=2, shape=(n, n))
  fg = fgraph.FactorGraph(variable_groups=variables)
  pairs = []
  for ii in range(n):
    for jj in range(n):
      kk, ll = (ii + 1) % n, (jj + 1) % n
      pairs.append([variables[ii, jj], variables[kk, jj]])
      pairs.append([variables[ii, jj], variables[ii, ll]])
  fg.add_factors(
      fgroup.PairwiseFactorGroup(
          variables_for_factors=pairs,
          log_potential_matrix=coupling * np.array([[1.0, -1.0], [-1.0, 1.0]]),
      )
  )
  rng = np.random.default_rng(seed)
  shape = (n, n, 2) if batch is None else (batch, n, n, 2)
  return fg, variables, rng.gumbel(size=shape)


def rbm_model(W, bh, bv):
  """RBM as in benchmark/rbm_lib.py:138-169: hidden and visible unary EnumFactors,
  then one pairwise factor per (hidden, visible) pair with log-potential W at (1, 1).
  Returns (fg, hidden_vars, visible_vars)."""
  nh, nv = bh.shape[0], bv.shape[0]
  hidden = vgroup.NDVarArray(num_states=2, shape=(nh,))
  visible = vgroup.NDVarArray(num_states=2, shape=(nv,))
  fg = fgraph.FactorGraph(variable_groups=[hidden, visible])
  unary_cfg = np.arange(2)[:, None]
  hidden_unaries = fgroup.EnumFactorGroup(
      variables_for_factors=[[hidden[i]] for i in range(nh)],
      factor_configs=unary_cfg,
      log_potentials=np.stack([np.zeros_like(bh), bh], axis=1),
  )
  visible_unaries = fgroup.EnumFactorGroup(
      variables_for_factors=[[visible[j]] for j in range(nv)],
      factor_configs=unary_cfg,
      log_potentials=np.stack([np.zeros_like(bv), bv], axis=1),
  )
  lpm = np.zeros((nh * nv, 2, 2))
  lpm[:, 1, 1] = W.ravel()
  pairwise = fgroup.PairwiseFactorGroup(
      variables_for_factors=[[hidden[i], visible[j]] for i in range(nh) for j in range(nv)],
      log_potential_matrix=lpm,
  )
  fg.add_factors([hidden_unaries, visible_unaries, pairwise])
  return fg, hidden, visible


def heretic_model(seed: int = 0):
  """The reference's e2e "heretic" model (tests/test_pgmax.py:424-475): 28 x 28 hidden variables
  of 17 states, 30 x 30 pixel variables of 3 states, 9 PairwiseFactorGroups (one per offset of a
  3 x 3 window) with a shared 17 x 3 potential matrix each: 7 056 factors of 20 edge-states.
  Returns (fg, pixel_vars, hidden_vars)."""
  im_size = (30, 30)
  pixel_vars = vgroup.NDVarArray(shape=im_size, num_states=3)
  hidden_vars = vgroup.NDVarArray(shape=(im_size[0] - 2, im_size[1] - 2), num_states=17)
  fg = fgraph.FactorGraph([pixel_vars, hidden_vars])
  w_pot = np.random.RandomState(seed).normal(size=(17, 3, 3, 3))
  for k_row in range(3):
    for k_col in range(3):
      fg.add_factors(fgroup.PairwiseFactorGroup(
          variables_for_factors=[[hidden_vars[r, c], pixel_vars[r + k_row, c + k_col]]
                                 for r in range(28) for c in range(28)],
          log_potential_matrix=w_pot[:, :, k_row, k_col]))
  return fg, pixel_vars, hidden_vars


def rbm_energy(hidden, visible, W, bh, bv):
  """Energy of an RBM configuration (benchmark/rbm_lib.py calc_energies)."""
  return -(hidden @ bh) - (visible @ bv) - hidden @ W @ visible


def _logical_enum_configs(kind: str, num_parents: int) -> np.ndarray:
  """All valid configurations of an OR / AND / Pool factor with `num_parents`
  parents (choices) and the child (indicator) last — the equivalent EnumFactor of
  tests/factor/test_or.py:120-140, test_and.py, test_pool.py."""
  configs = np.array(list(itertools.product([0, 1], repeat=num_parents + 1)))
  parents, child = configs[:, :-1], configs[:, -1]
  if kind == "or":
    ok = (parents.sum(axis=1) >= 1) == (child == 1)
  elif kind == "and":
    ok = (parents.sum(axis=1) == num_parents) == (child == 1)
  elif kind == "pool":
    ok = parents.sum(axis=1) == child
  else:
    raise ValueError(kind)
  return configs[ok]


def logical_pair(kind: str, seed: int, parents_range=(1, 10)):
  """Two equivalent graphs in the spirit of the reference's differential tests
  (tests/factor/test_or.py:30-290): graph A holds the first half of the factors
  as `kind` factors and the second half as their equivalent EnumFactors; graph B
  the other way round (seed 0: all-logical vs all-enum).  Returns a dict with the
  two graphs, their variable groups, and shared random evidence / initial messages."""
  rng = np.random.RandomState(seed)
  num_factors = rng.randint(10, 20)
  num_parents = rng.randint(parents_range[0], parents_range[1], num_factors)
  cum = np.insert(np.cumsum(num_parents), 0, 0)
  group_cls = {"or": fgroup.ORFactorGroup, "and": fgroup.ANDFactorGroup,
               "pool": fgroup.PoolFactorGroup}[kind]
  split = num_factors if seed == 0 else num_factors // 2

  graphs = []
  for which in range(2):
    parents = vgroup.NDVarArray(num_states=2, shape=(int(num_parents.sum()),))
    children = vgroup.NDVarArray(num_states=2, shape=(num_factors,))
    fg = fgraph.FactorGraph(variable_groups=[parents, children])
    vff = [[parents[i] for i in range(cum[f], cum[f + 1])] + [children[f]]
           for f in range(num_factors)]
    as_logical = list(range(split)) if which == 0 else list(range(split, num_factors))
    as_enum = [f for f in range(num_factors) if f not in as_logical]
    for f in as_enum:
      cfg = _logical_enum_configs(kind, int(num_parents[f]))
      fg.add_factors(factor.EnumFactor(variables=vff[f], factor_configs=cfg,
                                       log_potentials=np.zeros(cfg.shape[0])))
    if as_logical:
      fg.add_factors(group_cls(variables_for_factors=[vff[f] for f in as_logical]))
    graphs.append((fg, parents, children))
  ev_parents = rng.gumbel(size=(int(num_parents.sum()), 2))
  ev_children = rng.gumbel(size=(num_factors, 2))
  msg_parents = rng.normal(size=(int(num_parents.sum()), 2))
  msg_children = rng.normal(size=(num_factors, 2))
  return dict(graphs=graphs, ev_parents=ev_parents, ev_children=ev_children,
              msg_parents=msg_parents, msg_children=msg_children,
              num_factors=num_factors, num_parents=num_parents)


def init_logical(bp, entry, data):
  """BPArrays for one graph of logical_pair: evidence + per-variable initial messages
  (update_ftov_msgs by variable spreads data / num_edges, bp_state.py:208-228)."""
  _, parents, children = entry
  msgs = {parents[i]: data["msg_parents"][i] for i in range(data["msg_parents"].shape[0])}
  msgs.update({children[i]: data["msg_children"][i] for i in range(data["msg_children"].shape[0])})
  return bp.init(
      evidence_updates={parents: data["ev_parents"], children: data["ev_children"]},
      ftov_msgs_updates=msgs,
  )


def rcn_valid_configs(r: int, hps: int, vps: int) -> np.ndarray:
  """Valid (state0, state1) pairs of an RCN lateral factor with perturb radius r
  (examples/rcn.ipynb cell 24): pool positions within a Chebyshev box of radius r."""
  rows, cols = 2 * hps + 1, 2 * vps + 1
  r1, c1 = np.divmod(np.arange(rows * cols), cols)
  configs = []
  for i in range(rows * cols):
    rr = np.arange(max(r1[i] - r, 0), min(r1[i] + r, 2 * hps) + 1)
    cc = np.arange(max(c1[i] - r, 0), min(c1[i] + r, 2 * vps) + 1)
    j = (rr[:, None] * cols + cc[None, :]).ravel()
    configs.append(np.stack([np.full(j.shape, i), j], axis=1))
  return np.concatenate(configs)


def rcn_model(num_models=20, num_vars=80, hps=12, vps=12, radii=(2, 3, 5, 8), extra_edges=80, seed=0):
  """Synthetic RCN-shaped graph (examples/rcn.ipynb cells 21-26; rcn.npz is not shipped):
  per model `num_vars` variables with (2 hps + 1)(2 vps + 1) states, lateral pairwise
  EnumFactors along a random spanning tree plus `extra_edges` random edges, each factor its
  OWN EnumFactorGroup with the valid configs of a random perturb radius and zero potentials.
  Returns (fg, variable groups, {VarGroup: evidence array})."""
  rng = np.random.default_rng(seed)
  M = (2 * hps + 1) * (2 * vps + 1)
  tables = {r: rcn_valid_configs(r, hps, vps) for r in radii}
  groups = [vgroup.NDVarArray(num_states=M, shape=(num_vars,)) for _ in range(num_models)]
  fg = fgraph.FactorGraph(groups)
  for vg in groups:
    pairs = set()
    order = rng.permutation(num_vars)
    for k in range(1, num_vars):  # random spanning tree
      pairs.add((int(order[rng.integers(k)]), int(order[k])))
    if num_vars - 1 + extra_edges > num_vars * (num_vars - 1) // 2:
      raise ValueError("more edges requested than distinct variable pairs exist")
    while len(pairs) < num_vars - 1 + extra_edges:
      i, j = rng.integers(num_vars, size=2)
      if i != j and (int(i), int(j)) not in pairs and (int(j), int(i)) not in pairs:
        pairs.add((int(i), int(j)))
    for i, j in sorted(pairs):
      r = int(rng.choice(radii))
      fg.add_factors(fgroup.EnumFactorGroup(
          variables_for_factors=[[vg[i], vg[j]]], factor_configs=tables[r]))
  # bottom-up evidence in {+1, -1}: sparse edge map gathered per pool window (cell 34)
  evidence = {vg: np.where(rng.random((num_vars, M)) < 0.05, 1.0, -1.0) for vg in groups}
  return fg, groups, evidence


def deconv_model(im_height=28, im_width=28, n_feat=5, feat_height=6, feat_width=6, n_chan=1):
  """Binary deconvolution graph for ONE image (examples/pmp_binary_deconvolution.ipynb
  cells 12-14): AND factors (S, W -> SW) and OR factors (SW... -> X).
  Returns (fg, dict of the variable groups S, W, SW, X)."""
  s_height, s_width = im_height - feat_height + 1, im_width - feat_width + 1
  W = vgroup.NDVarArray(num_states=2, shape=(n_chan, n_feat, feat_height, feat_width))
  S = vgroup.NDVarArray(num_states=2, shape=(1, n_feat, s_height, s_width))
  SW = vgroup.NDVarArray(num_states=2, shape=(1, n_chan, im_height, im_width, n_feat, feat_height, feat_width))
  X = vgroup.NDVarArray(num_states=2, shape=(1, n_chan, im_height, im_width))
  fg = fgraph.FactorGraph(variable_groups=[S, W, SW, X])
  and_vars, or_vars = [], {}
  for c in range(n_chan):
    for sh in range(s_height):
      for sw in range(s_width):
        for f in range(n_feat):
          for fh in range(feat_height):
            for fw in range(feat_width):
              ih, iw = fh + sh, fw + sw
              sw_var = SW[0, c, ih, iw, f, fh, fw]
              and_vars.append([S[0, f, sh, sw], W[c, f, fh, fw], sw_var])
              or_vars.setdefault((c, ih, iw), []).append(sw_var)
  fg.add_factors(fgroup.ANDFactorGroup(and_vars))
  fg.add_factors(fgroup.ORFactorGroup(
      [parents + [X[0, c, ih, iw]] for (c, ih, iw), parents in or_vars.items()]))
  return fg, dict(S=S, W=W, SW=SW, X=X)


def deconv_evidence(groups, batch, seed=0, pW=0.25, pS=1e-75, pX=1e-100):
  """Evidence of the deconvolution notebook (cells 19-21) for `batch` synthetic images:
  random 5x5 features OR-convolved at Bernoulli locations; Gumbel noise on S and W."""
  from scipy.special import logit
  rng = np.random.default_rng(seed)
  S, W, SW, X = groups["S"], groups["W"], groups["SW"], groups["X"]
  _, n_chan, ih, iw = X.shape
  imgs = np.zeros((batch, 1, n_chan, ih, iw))
  feats = rng.random((4, 5, 5)) < 0.4
  for b in range(batch):
    for _ in range(6):
      f, r, c = rng.integers(4), rng.integers(ih - 5), rng.integers(iw - 5)
      imgs[b, 0, :, r : r + 5, c : c + 5] = np.maximum(imgs[b, 0, :, r : r + 5, c : c + 5], feats[f])
  uW = np.zeros((batch,) + W.shape + (2,)); uW[..., 1] = logit(pW)
  uS = np.zeros((batch,) + S.shape + (2,)); uS[..., 1] = logit(pS)
  uX = np.zeros((batch,) + X.shape + (2,)); uX[..., 0] = (2 * imgs - 1) * logit(pX)
  return {S: uS + rng.gumbel(size=uS.shape), W: uW + rng.gumbel(size=uW.shape),
          SW: np.zeros((batch,) + SW.shape + (2,)), X: uX}


def sdlp_ising_model(grid_size: int = 4, num_states: int = 3, seed: int = 0, scale: float = 0.01):
  """The "fully connected" categorical Ising model of tests/lp/test_dual_lp.py:44-66: pairwise
  factors between (i, j) and every (k, l) with k > i and l > j, potentials scale * N(0, 1)."""
  rng = np.random.RandomState(1000 + seed)
  variables = vgroup.NDVarArray(num_states=num_states, shape=(grid_size, grid_size))
  fg = fgraph.FactorGraph(variable_groups=variables)
  pairs = [[variables[i, j], variables[k, l]]
           for i in range(grid_size) for j in range(grid_size)
           for k in range(i + 1, grid_size) for l in range(j + 1, grid_size)]
  fg.add_factors(fgroup.PairwiseFactorGroup(
      variables_for_factors=pairs,
      log_potential_matrix=scale * rng.normal(size=(len(pairs), num_states, num_states))))
  return fg, variables


def sdlp_line_model(line_length: int = 20, seed: int = 0):
  """Line sparsification with ORFactors (tests/lp/test_dual_lp.py:152-195): bottom variables are
  on, each is explained by any of its (up to) 3 closest top variables; top variables prefer off."""
  rng = np.random.RandomState(seed)
  top = vgroup.NDVarArray(num_states=2, shape=(line_length,))
  bottom = vgroup.NDVarArray(num_states=2, shape=(line_length,))
  fg = fgraph.FactorGraph(variable_groups=[top, bottom])
  vff = []
  for f in range(line_length):
    parents = [top[f]]
    if f >= 1:
      parents.append(top[f - 1])
    if f <= line_length - 2:
      parents.append(top[f + 1])
    vff.append(parents + [bottom[f]])
  fg.add_factors(fgroup.ORFactorGroup(vff))
  ev_bottom = np.zeros((line_length, 2))
  ev_bottom[..., 0] = -10_000
  ev_top = np.zeros((line_length, 2))
  ev_top[..., 1] = -100
  evidence = {top: ev_top + rng.gumbel(size=ev_top.shape), bottom: ev_bottom}
  return fg, top, bottom, evidence


def sdlp_pool_model(n_layers: int = 4, choices_per_pool: int = 2):
  """Hierarchy of PoolFactors (tests/lp/test_bp_for_lp.py:288-315): layer n holds 2^n pool
  variables; every variable of layer n is the indicator of a pool over 2 choices of layer n + 1."""
  per_layer = [choices_per_pool**i for i in range(n_layers)]
  cum = np.insert(np.cumsum(per_layer), 0, 0)
  variables = vgroup.NDVarArray(num_states=2, shape=(int(cum[-1]),))
  fg = fgraph.FactorGraph(variable_groups=[variables])
  vff = []
  for layer in range(n_layers - 1):
    start = cum[layer + 1]
    for indicator in range(cum[layer], cum[layer + 1]):
      vff.append([variables[start + c] for c in range(choices_per_pool)] + [variables[indicator]])
      start += choices_per_pool
  fg.add_factors(fgroup.PoolFactorGroup(vff))
  return fg, variables


def lp_bp_cases():
  """(name, graph, evidence updates, factor type) of the reference's three
  convergence / consistency tests (tests/lp/test_bp_for_lp.py:28-391), one seed each."""
  cases = []
  fg, variables = sdlp_ising_model(seed=7, scale=1.0)
  rng = np.random.RandomState(0)
  cases.append(("enum", fg, {variables: rng.gumbel(size=(4, 4, 3))}, factor.EnumFactor))
  fg, top, bottom, evidence = sdlp_line_model(seed=1)
  cases.append(("or", fg, evidence, factor.ORFactor))
  fg, variables = sdlp_pool_model()
  updates = np.random.RandomState(2).gumbel(size=(variables.shape[0], 2))
  updates[0, 1] = 10
  cases.append(("pool", fg, {variables: updates}, factor.PoolFactor))
  return cases


def check_lp_bp_properties(get_bp_updates, context, temperature, atol=1e-5):
  """The two properties of tests/lp/test_bp_for_lp.py: (1) the BP updates at a very low temperature
  are within `temperature` of the max-product updates; (2) the max / logsumexp of a factor's
  outgoing messages over its configurations is the same at every edge of the factor."""
  updates_t0, maxes_t0 = get_bp_updates(0.0)
  updates_t, lse_t = get_bp_updates(temperature)
  np.testing.assert_allclose(updates_t0, updates_t, atol=temperature, rtol=0)
  edge_of_es = np.asarray(context.edge_indices_for_edge_states)
  factor_of_es = np.asarray(context.factor_indices_for_edge_states)
  factor_of_edge = np.zeros((context.num_edges,), dtype=np.int64)
  factor_of_edge[edge_of_es] = factor_of_es
  for vals in (np.asarray(maxes_t0), np.asarray(lse_t)):
    lo = np.full((context.num_factors,), np.inf)
    hi = np.full((context.num_factors,), -np.inf)
    np.minimum.at(lo, factor_of_edge, vals)
    np.maximum.at(hi, factor_of_edge, vals)
    np.testing.assert_allclose(lo, hi, atol=atol, rtol=1e-6)


def sdlp_and_model(num_rows: int = 10, num_cols: int = 5, p_on: float = 0.8, seed: int = 0):
  """ANDFactors finding the rows of a binary matrix that are all ones
  (tests/lp/test_dual_lp.py:242-281).  Returns (fg, matrix, all_ones, evidence updates, ground truth)."""
  rng = np.random.RandomState(seed)
  matrix = vgroup.NDVarArray(num_states=2, shape=(num_rows, num_cols))
  all_ones = vgroup.NDVarArray(num_states=2, shape=(num_rows,))
  fg = fgraph.FactorGraph(variable_groups=[matrix, all_ones])
  fg.add_factors(fgroup.ANDFactorGroup(
      [[matrix[r, c] for c in range(num_cols)] + [all_ones[r]] for r in range(num_rows)]))
  obs = rng.binomial(1, p_on, (num_rows, num_cols))
  evidence = np.zeros((num_rows, num_cols, 2))
  evidence[..., 1] = 1_000 * (2 * obs - 1)
  return fg, matrix, all_ones, {matrix: evidence}, np.all(obs, axis=1).astype(int)
