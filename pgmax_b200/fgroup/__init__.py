"""Factor groups (host mirror of the pgmax.fgroup sub-package)."""

from pgmax_b200.fgroup.enum import EnumFactorGroup
from pgmax_b200.fgroup.enum import PairwiseFactorGroup
from pgmax_b200.fgroup.fgroup import FactorGroup
from pgmax_b200.fgroup.fgroup import SingleFactorGroup
from pgmax_b200.fgroup.logical import ANDFactorGroup
from pgmax_b200.fgroup.logical import LogicalFactorGroup
from pgmax_b200.fgroup.logical import ORFactorGroup
from pgmax_b200.fgroup.pool import PoolFactorGroup
