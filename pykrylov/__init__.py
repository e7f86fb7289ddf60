"""Drop-in alias: ``import pykrylov`` resolves to the B200-native engine.

``from pykrylov.linop import PysparseLinearOperator``, ``from pykrylov.cgs import CGS`` ...
(the imports of the reference's examples/bmark.py:6-9) work unchanged; every
sub-package is the corresponding pykrylov_b200 module.
"""
import importlib
import sys

import pykrylov_b200 as _impl

__version__ = _impl.__version__

for _name in ("generic", "linop", "cg", "cgs", "tfqmr", "bicgstab", "minres", "symmlq", "lls", "gallery", "tools"):
    _mod = importlib.import_module("pykrylov_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
    for _sub in list(sys.modules):
        if _sub.startswith("pykrylov_b200." + _name + "."):
            sys.modules[__name__ + _sub[len("pykrylov_b200"):]] = sys.modules[_sub]
