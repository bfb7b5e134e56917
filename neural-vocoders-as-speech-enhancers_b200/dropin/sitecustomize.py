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


def _neutralise_anomaly_mode():
    """Opt-in (NVSE_B200_NO_ANOMALY=1): the reference's trainer switches autograd's anomaly detection on globally at import
    (train_time_wi_inv.py:5) -- a debugging aid that captures a Python stack trace for every autograd node of every forward and
    checks every backward output for NaNs with a device synchronisation; on a B200 it is most of a training step
    (tools/train_script_bench.py).  With the variable set the call becomes a no-op; the script itself stays untouched."""
    if os.environ.get("NVSE_B200_NO_ANOMALY", "0") in ("", "0"):
        return

    class _LazyPatch(importlib.abc.MetaPathFinder):  # patch torch.autograd once torch has been imported by the script
        def find_spec(self, fullname, path=None, target=None):
            return None

    import builtins
    real_import = builtins.__import__

    def patched_import(name, globals=None, locals=None, fromlist=(), level=0):
        mod = real_import(name, globals, locals, fromlist, level)
        if name == "torch" or name.startswith("torch."):
            t = sys.modules.get("torch")
            ag = getattr(t, "autograd", None) if t is not None else None
            if ag is not None and hasattr(ag, "set_detect_anomaly") and not getattr(ag.set_detect_anomaly, "_nvse_noop", False):
                class _NoAnomaly:
                    _nvse_noop = True

                    def __init__(self, mode=True, check_nan=True):
                        pass

                    def __enter__(self):
                        return self

                    def __exit__(self, *a):
                        return False
                ag.set_detect_anomaly = _NoAnomaly
                builtins.__import__ = real_import  # done: restore the plain import
        return mod

    builtins.__import__ = patched_import


_neutralise_anomaly_mode()


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
