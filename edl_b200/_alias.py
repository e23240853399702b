"""Import-alias machinery: makes ``paddle_edl.x.y`` / ``edl.x.y`` resolve to ``edl_b200.x.y``
(same module objects, so singletons and isinstance checks agree across the three names)."""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys

_TARGET = "edl_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def __init__(self, alias):
        self.alias = alias

    def find_spec(self, fullname, path=None, target=None):
        if fullname != self.alias and not fullname.startswith(self.alias + "."):
            return None
        real = _TARGET + fullname[len(self.alias):]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        spec = importlib.machinery.ModuleSpec(fullname, _AliasLoader(real),
                                              is_package=real_spec.submodule_search_locations is not None)
        return spec


def install_alias(alias: str) -> None:
    real = importlib.import_module(_TARGET)
    if not any(isinstance(f, _AliasFinder) and f.alias == alias for f in sys.meta_path):
        sys.meta_path.insert(0, _AliasFinder(alias))
    pkg = sys.modules.get(alias)
    if pkg is not None:
        # expose the real package's public names on the alias package object
        for k, v in vars(real).items():
            if not k.startswith("__"):
                setattr(pkg, k, v)
        pkg.__path__ = list(getattr(real, "__path__", []))
