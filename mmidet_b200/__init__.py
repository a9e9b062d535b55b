"""mmidet_b200 -- importable name of the package whose sources live in ./mmi-det_b200/ (the hyphenated directory
name the project layout prescribes cannot be imported directly, so this shim extends __path__ to it)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mmi-det_b200")
__path__.append(_real)

from . import _lib  # noqa: E402,F401
from .ops import selective_scan  # noqa: E402,F401
