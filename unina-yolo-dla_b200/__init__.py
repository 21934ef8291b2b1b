"""B200-native UNINA-YOLO-DLA inference hot path (host side).

The importable name is ``unina_yolo_dla_b200`` (a shim package next to this directory maps
onto it, since ``unina-yolo-dla_b200`` is not a valid Python identifier).
"""
from ._lib import LIB_PATH, UydError, lib  # noqa: F401
from .plan import Plan, Slice, fold_bn  # noqa: F401
from .custom import UninaCustomB200  # noqa: F401
from .yolo import DEFAULT_YAML, UninaYoloB200  # noqa: F401

__all__ = ["UninaYoloB200", "UninaCustomB200", "Plan", "Slice", "fold_bn", "UydError", "lib", "LIB_PATH", "DEFAULT_YAML"]
