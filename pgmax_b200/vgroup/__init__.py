"""Variable groups (host mirror of the pgmax.vgroup sub-package)."""

from pgmax_b200.vgroup.varray import NDVarArray
from pgmax_b200.vgroup.vdict import VarDict
from pgmax_b200.vgroup.vgroup import VarGroup
