"""CPU oracle for the sparsified K-means hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this package.  The product (sparsifiedkmeans_b200) never does and
fails loudly when its CUDA library is missing.

Contents
  skm_oracle.c   plain-C restatement of the reference's native arithmetic
                 (SparseMatrixMinusCluster / InnerProduct / ColumnNormSq /
                 hadamard, centroid update), wrapped by `oracle.cport`
  host_ref.py    float64 numpy restatement of the reference's MATLAB host logic
                 (findClusterAssignments.m, kmeans_sparsified.m Lloyd loop,
                 Arthur_initialization.m, mix/unmix, sampler contract)
  refmex.py      ctypes driver for the reference's own C files compiled
                 unmodified into oracle/_ref (the strongest checker)
  mexstub/       stub mex.h + runtime used to build oracle/_ref and to test mex/
Parity status: pinned against oracle/_ref and tests/golden (see tests/test_oracle.py).
"""
