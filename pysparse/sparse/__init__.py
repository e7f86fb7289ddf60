from . import pysparseMatrix      # noqa: F401
