"""Load the UNMODIFIED reference (maksimt/rri_nmf, Python 2 sources) under Python 3.12.

TEST INFRASTRUCTURE ONLY.  Used in this container (where /root/reference exists) to pin the
numpy restatement in `oracle/rri_oracle.py` and to generate the golden vectors under
`tests/golden/` (see `oracle/make_golden.py`).  It never runs on the GPU box (the reference does
not travel) and nothing in the product package `rri_nmf_b200/` imports it.

The shim does not edit reference files; it only patches the interpreter environment
(SURVEY.md App. B):
  * py2 implicit-relative imports  -> put `<ref>/src/rri_nmf` itself on sys.path and import the
    modules top-level (`nmf`, `optimization`, `matrixops`, `initialization`, `sklearn_interface`)
  * `numexpr` absent               -> fake module whose `evaluate` evals in the caller's frame
  * `time.clock` removed in 3.8    -> alias to `time.process_time`
  * `scipy.maximum/minimum/argmax/sum`, `np.alltrue`, `np.int` removed -> alias to NumPy
  * logger trap (nmf.py:46-47,366-367): the module logger level is NOTSET (=0) so
    `logger.level <= logging.DEBUG` is true and the reference recomputes the full objective
    around every update (and crashes on py2 `func_name`).  We set it to WARNING.
"""
import importlib
import logging
import os
import sys
import time
import types
import warnings

REF_ROOT = os.environ.get("RRI_REFERENCE_ROOT", "/root/reference")
_cached = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "src", "rri_nmf", "nmf.py"))


def load():
    """Return a namespace with the reference modules: .nmf (module), .optimization, .matrixops,
    .initialization, .sklearn_interface."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    import numpy as np
    import scipy

    if "numexpr" not in sys.modules:
        fake = types.ModuleType("numexpr")

        def evaluate(expr, local_dict=None, global_dict=None, **kw):
            fr = sys._getframe(1)
            return eval(expr, dict(fr.f_globals), dict(fr.f_locals))

        fake.evaluate = evaluate
        fake.set_num_threads = lambda n: None
        fake.detect_number_of_cores = lambda: 1
        sys.modules["numexpr"] = fake
    if not hasattr(time, "clock"):
        time.clock = time.process_time
    for name in ("maximum", "minimum", "argmax", "sum"):
        if not hasattr(scipy, name):
            setattr(scipy, name, getattr(np, name))
    if not hasattr(np, "alltrue"):
        np.alltrue = np.all
    if not hasattr(np, "int"):
        np.int = int

    src = os.path.join(REF_ROOT, "src", "rri_nmf")
    if src not in sys.path:
        sys.path.insert(0, src)
    ns = types.SimpleNamespace()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # `nmf` would collide with nothing of ours: our package is `rri_nmf_b200.nmf`
        for m in ("matrixops", "optimization", "initialization", "nmf", "sklearn_interface"):
            setattr(ns, m, importlib.import_module(m))
    ns.nmf.logger.setLevel(logging.WARNING)
    ns.optimization.logger.setLevel(logging.WARNING)
    _cached = ns
    return ns
