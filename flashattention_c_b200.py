"""Import shim: the package directory is literally named `flashattention.c_b200/` (the project's layout
contract), which Python's import statement cannot spell.  `import flashattention_c_b200` loads that
directory as a regular package under this importable name."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "flashattention.c_b200"
_spec = importlib.util.spec_from_file_location(
    "flashattention_c_b200", str(_pkg_dir / "__init__.py"), submodule_search_locations=[str(_pkg_dir)]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["flashattention_c_b200"] = _mod
_spec.loader.exec_module(_mod)
