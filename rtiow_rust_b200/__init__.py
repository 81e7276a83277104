"""Importable alias for the package directory `rtiow-rust_b200/` (a hyphen is not a valid Python
identifier).  All code lives in ../rtiow-rust_b200/; this module only points its search path there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rtiow-rust_b200"))

from .api import *  # noqa: F401,F403,E402
from .api import __all__  # noqa: F401,E402
