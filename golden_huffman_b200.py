"""Import shim: the package directory is named `golden-huffman_b200` (not a Python identifier), so
`import golden_huffman_b200` loads it from there."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "golden-huffman_b200")
_spec = importlib.util.spec_from_file_location(
    "golden_huffman_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["golden_huffman_b200"] = _mod
_spec.loader.exec_module(_mod)
