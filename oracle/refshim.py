"""TEST INFRASTRUCTURE ONLY -- load the *unmodified* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` and by the ``not gpu`` tests that pin the numpy
restatement against the live reference.  Nothing is written to the reference
tree; the shim rebinds one name in the imported module's namespace:

* ``mbpls.mbpls.check_array`` -> wrapper translating the keyword
  ``force_all_finite=`` (removed from scikit-learn >= 1.8) into
  ``ensure_all_finite=`` (call sites ``mbpls/mbpls.py:293,310,318,336,342,1094,...``).

``traced_fit`` additionally records the NIPALS trip count per component with
``sys.settrace`` (line 839 opens a component, line 914 closes a trip), because
the reference keeps ``run`` in a local (``mbpls/mbpls.py:839,914``).
"""
from __future__ import annotations

import functools
import os
import sys
import warnings

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "mbpls", "mbpls.py"))


@functools.lru_cache(maxsize=1)
def load():
    """Return the reference ``MBPLS`` class (shimmed)."""
    if not available():
        raise RuntimeError("reference tree not mounted at " + REFERENCE_ROOT)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mbpls.mbpls as ref  # noqa: E402

    orig = ref.check_array
    if not getattr(orig, "_graft_shim", False):
        @functools.wraps(orig)
        def check_array(*a, force_all_finite=True, **k):
            return orig(*a, ensure_all_finite=force_all_finite, **k)
        check_array._graft_shim = True
        ref.check_array = check_array
    return ref.MBPLS


def traced_fit(model, X, Y):
    """``model.fit(X, Y)`` on a reference estimator; returns NIPALS trips per component."""
    code = type(model).fit.__code__
    trips = []

    def tracer(frame, event, arg):
        if frame.f_code is not code:
            return None

        def line_tracer(frame, event, arg):
            if event == "line":
                if frame.f_lineno == 839:
                    trips.append(0)
                elif frame.f_lineno == 914:
                    trips[-1] += 1
            return line_tracer
        return line_tracer

    old = sys.gettrace()
    sys.settrace(tracer)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(X, Y)
    finally:
        sys.settrace(old)
    return trips
