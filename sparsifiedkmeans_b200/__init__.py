"""sparsifiedkmeans_b200 -- B200-native engine for the sparsified K-means hot path.

Host-side mirror of the reference's interface (same names, argument meaning and error
behaviour) over libskm_b200.so, hand-written sm_100a CUDA behind a C ABI
(include/skm_b200.h).  There is no CPU fallback.
"""
from .engine import (Context, Dataset, IterStats, Lloyd, dct_mix, default_context,     # noqa: F401
                     fwht_f32_inplace, lloyd_step_host, mix_hadamard, sample_rows, sample_rows_general,
                     second_pass)
from .find_cluster_assignments import findClusterAssignments                  # noqa: F401
from .kmeans import (Arthur_initialization, KMeansError, kmeans_sparsified,    # noqa: F401
                     randsample_block, randsample_fixedNumberEntries)
from .multi import MultiContext, MultiDataset, MultiLloyd                      # noqa: F401
from .ops import (SparseMatrixColumnNormSq, SparseMatrixInnerProduct,         # noqa: F401
                  SparseMatrixMinusCluster, hadamard, hadamard_pthreads)

__version__ = "0.1.0"
