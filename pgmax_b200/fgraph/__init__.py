"""Factor graph (host mirror of the pgmax.fgraph sub-package)."""

from pgmax_b200.fgraph.fgraph import FactorGraph
from pgmax_b200.fgraph.fgraph import FactorGraphState
