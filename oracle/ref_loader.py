"""Import the REAL reference (dherrera1911/sqfa) without its matplotlib-dependent `plot` package.

TEST INFRASTRUCTURE. Works only where the reference sources exist (the build container:
/root/reference/src). On the GPU box they do not, `available()` is False and every caller must
skip. Nothing in `-m gpu` tests, `smoke()` or `bench.py` may depend on this module at run time.
"""

import importlib.util
import os
import sys
import types

_CANDIDATES = ("/root/reference/src/sqfa",)
_PKG = "_sqfa_reference"
_MODULES = ("linalg", "statistics", "distances", "constraints", "_optim", "model")


def _root():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "model.py")):
            return c
    return None


def available():
    return _root() is not None


def load():
    """Return a namespace with the reference's submodules (statistics, linalg, distances,
    constraints, _optim, model) loaded under a private package name, skipping sqfa/__init__.py
    (which imports matplotlib through `plot`)."""
    root = _root()
    if root is None:
        raise RuntimeError("reference sources not found (expected /root/reference/src/sqfa)")
    if _PKG in sys.modules:
        return sys.modules[_PKG]
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [root]
    sys.modules[_PKG] = pkg
    for name in _MODULES:
        spec = importlib.util.spec_from_file_location(f"{_PKG}.{name}", os.path.join(root, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{_PKG}.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return pkg
