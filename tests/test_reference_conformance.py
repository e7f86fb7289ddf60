"""The reference's OWN operator tests, run against this package.

oracle/make_ref.py keeps a transliterated copy of the reference's test files next to the
runnable reference (oracle/_ref, git-ignored, present wherever the oracle was built).  Here
`pykrylov/linop/tests/test_linop.py` (400 lines: T/H inference, algebra and error types,
dtype promotion over allowed_types^2, Identity / Diagonal / Zero / Reduced /
SymmetricallyReduced / linop_from_ndarray) is re-targeted from the reference package to
`pykrylov_b200` by renaming the import and executed unchanged otherwise.  Everything in it is
host-side operator glue (closures run on the host), so it needs no GPU -- except the two
`CoordLinearOperator` cases, whose operator is a CSR in HBM in this package: their host half
(COO -> CSR in the reference's accumulation order) is pinned in
tests/test_host.py::test_coo_to_csr_keeps_reference_accumulation_order, the device half is the
same `csr_operator` every GPU parity test goes through.
"""
import os
import types
import unittest

import pytest

from conftest import ROOT

REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "refpykrylov", "linop", "tests", "test_linop.py")


def _load_retargeted(path, target):
    src = open(path).read().replace("refpykrylov", target)
    mod = types.ModuleType("reference_test_linop_on_" + target)
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


@pytest.mark.skipif(not os.path.exists(REF_TESTS), reason="oracle/_ref not built (python oracle/make_ref.py)")
@pytest.mark.parametrize("target", ["pykrylov_b200", "pykrylov"])
def test_reference_linop_testsuite_passes_on_this_package(target):
    mod = _load_retargeted(REF_TESTS, target)
    suite = unittest.TestSuite()
    n_device = 0
    for case in unittest.defaultTestLoader.loadTestsFromModule(mod):
        for t in case:
            if type(t).__name__ == "test_CoordLinearOperator":      # device operator here
                n_device += 1
                continue
            suite.addTest(t)
    res = unittest.TestResult()
    suite.run(res)
    assert res.testsRun >= 15 and n_device == 2
    assert not res.failures and not res.errors, [(str(t), e[-400:]) for t, e in res.failures + res.errors]
