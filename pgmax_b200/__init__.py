"""pgmax_b200 — B200-native loopy belief propagation behind PGMax's inference API.

Sub-packages mirror the reference's: vgroup, factor, fgroup, fgraph, infer.
The message-passing loop itself lives in csrc/ (hand-written sm_100a kernels
behind the C ABI of include/pgx.h) and is reached through infer.BP.
"""

__version__ = "0.1.0"

from pgmax_b200 import factor
from pgmax_b200 import fgraph
from pgmax_b200 import fgroup
from pgmax_b200 import infer
from pgmax_b200 import utils
from pgmax_b200 import vgroup
