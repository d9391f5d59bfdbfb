"""Factor types and the fixed factor-type order of the message vector.

``FACTOR_TYPES`` fixes the order Enum -> OR -> AND -> Pool in which factor types
occupy the flat message / potential vectors; it plays the role of the keys of
the reference's FAC_TO_VAR_UPDATES registry (pgmax/factor/__init__.py:37-43).
The values of that registry (the jnp update functions) have no host
counterpart here: the updates are the CUDA kernels behind include/pgx.h.
"""

import collections

from pgmax_b200.factor.enum import EnumBlock
from pgmax_b200.factor.enum import EnumFactor
from pgmax_b200.factor.enum import EnumWiring
from pgmax_b200.factor.factor import concatenate_var_states_for_edges
from pgmax_b200.factor.factor import Factor
from pgmax_b200.factor.factor import Wiring
from pgmax_b200.factor.logical import ANDFactor
from pgmax_b200.factor.logical import LogicalFactor
from pgmax_b200.factor.logical import LogicalWiring
from pgmax_b200.factor.logical import ORFactor
from pgmax_b200.factor.pool import PoolFactor
from pgmax_b200.factor.pool import PoolWiring

FACTOR_TYPES = (EnumFactor, ORFactor, ANDFactor, PoolFactor)

# Same keys and order as the reference registry; the value names the device
# operator (kernel family in csrc/) that updates that slice of the messages.
FAC_TO_VAR_UPDATES = collections.OrderedDict([
    (EnumFactor, "pgx_enum_f2v"),
    (ORFactor, "pgx_logical_f2v"),
    (ANDFactor, "pgx_logical_f2v"),
    (PoolFactor, "pgx_pool_f2v"),
])
