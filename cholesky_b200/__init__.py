"""cholesky_b200: B200-native FP64 sparse Cholesky numeric factorization behind the C-level surface of
syamajala/cholesky (see include/cholesky.h)."""
from .engine import (Cholesky, CholeskyError, factor_binary_to_mtx, read_factor_binary, read_vector,  # noqa: F401
                     write_solution)
