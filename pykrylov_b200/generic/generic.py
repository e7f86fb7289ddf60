"""KrylovMethod: the attribute / keyword contract every solver keeps.

Same surface as the reference's pykrylov/generic/generic.py:11-98: the constructor
reads only ``abstol``, ``reltol``, ``precon`` and ``logger`` (anything else is
silently ignored, generic.py:74-77 -- the reference's own tests rely on that),
results are left as attributes.  New, engine-specific keywords default to the
reference's semantics:

  check_interval : iterations enqueued on the GPU between two reads of the
                   device status block (the stopping tests themselves run on
                   device every iteration, so results do not depend on it).
  context        : pykrylov_b200.device.Context to run on (default: cuda:LOCAL_RANK).
"""
import logging

null_log = logging.getLogger("krylov")
null_log.setLevel(logging.INFO)
null_log.addHandler(logging.NullHandler())


class KrylovMethod(object):

    def __init__(self, op, **kwargs):
        self.prefix = "Generic: "
        self.name = "Generic Krylov Method (must be subclassed)"
        self.op = op
        self.abstol = kwargs.get("abstol", 1.0e-8)
        self.reltol = kwargs.get("reltol", 1.0e-6)
        self.precon = kwargs.get("precon", None)
        self.logger = kwargs.get("logger", null_log)
        self.check_interval = int(kwargs.get("check_interval", 32))
        self.context = kwargs.get("context", None)

        self.residNorm = None
        self.residNorm0 = None
        self.residHistory = []
        self.nMatvec = 0
        self.nIter = 0
        self.converged = False
        self.bestSolution = None
        self.x = self.bestSolution

    def _write(self, msg):
        self.logger.info(msg)

    def solve(self, rhs, **kwargs):
        raise NotImplementedError("This method must be subclassed")
