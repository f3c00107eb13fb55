"""metabuli_b200 — B200-native `metabuli classify` hot path (CUDA sm_100a behind a C-ABI).

Package layout: csrc/ (CUDA kernels + C-ABI + C++ host), _lib/ (built artefacts), and the thin Python
plumbing used by the tests and bench.py (ctypes binding, DB readers, TSV writer)."""
from ._ffi import MblError, load_library  # noqa: F401
from .classifier import Classifier, ClassifyOptions  # noqa: F401
from .dbio import load_database  # noqa: F401
from .fastx import read_fastx  # noqa: F401
