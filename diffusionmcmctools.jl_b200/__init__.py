"""B200-native guided-proposal path updates behind the DiffusionMCMCTools.jl API.

The directory name carries a dot, so import it through the shim at the repo root:  `import dmt_b200`.
"""
from . import _lib, configs  # noqa: F401
from ._lib import Ctx, DmtError  # noqa: F401
