"""Import alias: `import aframe_gaussian_splatting_b200 as gs` loads the hyphenated package directory."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("aframe-gaussian-splatting_b200")
sys.modules[__name__] = _pkg
