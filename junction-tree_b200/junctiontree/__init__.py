"""B200-native sum-product propagation behind the junction-tree Python API.

Drop-in for the propagation path of jluttine/junction-tree: same import name and the same
public names as the reference's ``junctiontree/__init__.py`` (``from .junctiontree import *``).
See DESIGN.md for the architecture and INTEGRATION.md for the C ABI.
"""

from . import semirings  # noqa: F401
from .junctiontree import *  # noqa: F401,F403
from .junctiontree import (CliqueGraph, FactorGraph, JunctionTree, create_junction_tree,  # noqa: F401
                           einsum)

__version__ = "0.2.0+b200.1"
