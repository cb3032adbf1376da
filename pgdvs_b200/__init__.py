"""Importable alias of the `ml-pgdvs_b200/` package directory (a hyphen cannot appear in a
Python module name).  `import pgdvs_b200` executes ml-pgdvs_b200/__init__.py in this
namespace and resolves sub-modules (pgdvs_b200.ops, ...) from that directory."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "ml-pgdvs_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
