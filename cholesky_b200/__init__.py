"""cholesky_b200: B200-native FP64 sparse Cholesky numeric factorization behind the C-level surface of
syamajala/cholesky (see include/cholesky.h)."""
from .engine import Cholesky, CholeskyError, read_vector, write_solution  # noqa: F401
