"""The reference's OWN operator tests, run against this package.

oracle/make_ref.py keeps a transliterated copy of the reference's test files next to the
runnable reference (oracle/_ref, git-ignored, present wherever the oracle was built).  Here
`pykrylov/linop/tests/test_linop.py` (400 lines: T/H inference, algebra and error types,
dtype promotion over allowed_types^2, Identity / Diagonal / Zero / Reduced /
SymmetricallyReduced / linop_from_ndarray) is re-targeted from the reference package to
`pykrylov_b200` by renaming the import and executed unchanged otherwise.  Everything in it is
host-side operator glue (closures run on the host), so it needs no GPU -- except the two
`CoordLinearOperator` cases, whose operator is a CSR in HBM in this package.  Those two, and
the reference's CG test file (`pykrylov/cg/tests/test_diagdom.py`: Poisson 1-D up to n = 10 000
and 2-D up to 500 x 500 through closure-defined operators, i.e. the host-callback bridge with
the vectors and fused AXPY/dot kernels on the device), run here against the host emulation of
the device logic (conftest.emu_ctx, tests/emu) so that they are exercised on every machine.
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


REF_CG_TESTS = os.path.join(ROOT, "oracle", "_ref", "refpykrylov", "cg", "tests", "test_diagdom.py")


@pytest.mark.skipif(not os.path.exists(REF_TESTS), reason="oracle/_ref not built (python oracle/make_ref.py)")
def test_reference_device_dependent_testsuites_pass_on_the_emulated_device(emu_ctx, capsys):
    """CoordLinearOperator (device COO -> CSR operator) and the reference's own CG tests."""
    suite = unittest.TestSuite()
    lin = _load_retargeted(REF_TESTS, "pykrylov_b200")
    for case in unittest.defaultTestLoader.loadTestsFromModule(lin):
        for t in case:
            if type(t).__name__ == "test_CoordLinearOperator":
                suite.addTest(t)
    suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(_load_retargeted(REF_CG_TESTS, "pykrylov_b200")))
    res = unittest.TestResult()
    suite.run(res)
    capsys.readouterr()
    assert res.testsRun == 4                       # 2 Coord cases + Poisson1D + Poisson2D
    assert not res.failures and not res.errors, [(str(t), e[-600:]) for t, e in res.failures + res.errors]
