"""Imported by the interpreter at start-up when this directory is on PYTHONPATH.

``python train_time_wi_inv.py`` puts the script's own directory -- the reference checkout, which holds the
reference's ``dataset.py`` and ``Models/`` -- at ``sys.path[0]``, AHEAD of every PYTHONPATH entry.  A finder at the
front of ``sys.meta_path`` therefore resolves the two shadowed top-level names to this directory whatever the
order of ``sys.path``; everything else (``utils``, ``env``, ``Models.models`` ...) is found as usual.  The
inference scripts live in ``infers/`` and append the checkout to ``sys.path`` (infers/inference_hifigan.py:7-8),
so for them PYTHONPATH alone would do; the finder makes both cases the same.

If another ``sitecustomize`` sits further down ``sys.path`` (a site-wide hook) it is executed afterwards, so this
file does not hide it."""
import importlib.abc
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


class _NvseDropinFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == "dataset":
            return importlib.util.spec_from_file_location("dataset", os.path.join(_HERE, "dataset.py"))
        if fullname == "Models":
            pkg = os.path.join(_HERE, "Models")
            return importlib.util.spec_from_file_location("Models", os.path.join(pkg, "__init__.py"),
                                                          submodule_search_locations=[pkg])
        return None


if not any(type(f).__name__ == "_NvseDropinFinder" for f in sys.meta_path):
    sys.meta_path.insert(0, _NvseDropinFinder())


def _chain():
    here = os.path.realpath(_HERE)
    for entry in sys.path:
        cand = os.path.join(entry or os.getcwd(), "sitecustomize.py")
        if os.path.isfile(cand) and os.path.realpath(os.path.dirname(cand)) != here:
            spec = importlib.util.spec_from_file_location("_nvse_chained_sitecustomize", cand)
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except Exception as e:  # a broken site hook must not take the interpreter down
                sys.stderr.write(f"nvse dropin: chained sitecustomize {cand} failed: {e!r}\n")
            return


_chain()
