"""Helpers shared by the PYTHONPATH drop-in (``dataset.py``, ``Models/``, ``sitecustomize.py``):
find the reference checkout's own files on ``sys.path`` and import the B200 package that lives
one directory up (its directory name has hyphens, so it is loaded by file location)."""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(HERE)
PKG_NAME = os.path.basename(PKG_DIR)  # "neural-vocoders-as-speech-enhancers_b200"


def load_package():
    """The B200 package, imported once under its directory name (and the alias ``nvse_b200``)."""
    mod = sys.modules.get(PKG_NAME)
    if mod is None:
        spec = importlib.util.spec_from_file_location(PKG_NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                      submodule_search_locations=[PKG_DIR])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[PKG_NAME] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            sys.modules.pop(PKG_NAME, None)
            raise
    sys.modules.setdefault("nvse_b200", mod)
    return mod


def find_reference(relpath):
    """First ``<entry>/<relpath>`` over ``sys.path`` (and $NVSE_REFERENCE_ROOT) that is not inside this
    drop-in directory: the reference checkout's own copy of a file this directory shadows."""
    here = os.path.realpath(HERE)
    roots = [os.environ.get("NVSE_REFERENCE_ROOT")] + [p or os.getcwd() for p in sys.path]
    for root in roots:
        if not root:
            continue
        cand = os.path.realpath(os.path.join(root, relpath))
        if os.path.isfile(cand) and not cand.startswith(here + os.sep):
            return cand
    return None


def load_reference_module(name, relpath):
    """Import the reference's own ``relpath`` under the private module name ``name``."""
    mod = sys.modules.get(name)
    if mod is not None:
        return mod
    path = find_reference(relpath)
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    return mod
