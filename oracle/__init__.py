"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MB-PLS fitting hot path.

This package is a numpy restatement of the algorithms in the reference's
``mbpls/mbpls.py`` (and of scikit-learn's ``StandardScaler`` numerics, which
the reference calls).  It exists so the CUDA path can be checked on a GPU box
where ``/root/reference`` is absent.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``mbpls_b200/`` imports it, and the product path fails loudly when the CUDA
library is missing instead of falling back to this code.

Parity pin: ``oracle/make_golden.py`` (run in the build container, where the
read-only reference is mounted) checks this restatement against (i) the 69
known-answer CSVs of ``mbpls/tests/test_data`` consumed by
``mbpls/tests/test_mbpls.py:34-421`` and (ii) live runs of the shimmed
reference for everything the CSVs do not cover (NaN mode, PLS1,
standardize=False, calc_all=False, norms, trip counts).  The resulting vectors
are committed under ``tests/golden/`` and re-checked by ``pytest -m "not gpu"``.
"""
from .mbpls_oracle import (  # noqa: F401
    OracleMBPLS,
    OracleScaler,
    nan_census,
    legacy_ortho_group_rvs,
)
