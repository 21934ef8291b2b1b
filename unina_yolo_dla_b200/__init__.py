"""Importable alias of the ``unina-yolo-dla_b200/`` package directory (hyphens are not valid in
Python module names): this package's search path simply points there."""
from pathlib import Path as _Path

_real = _Path(__file__).resolve().parent.parent / "unina-yolo-dla_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
