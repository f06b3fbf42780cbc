"""Parity oracle for the GBD-PCG hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``mpcgpu_b200/``, ``include/``) never does.

* :mod:`oracle.pcg`    -- ctypes binding of ``pcg_oracle.c`` (fp32/fp64 restatement of
  ``GBD-PCG/include/pcg.cuh:98-217``) plus an fp64 direct block-tridiagonal solve for ground truth.
* :mod:`oracle.qdldl`  -- ctypes binding of ``oracle/_ref/libqdldl_ref.so`` (the reference's own
  ``qdldl/src/qdldl.c`` + ``qdldl_driver.c``), the CPU baseline.
* :mod:`oracle.refgpu` -- ctypes binding of ``oracle/_ref/libref_gbdpcg.so`` (the reference's own
  ``pcg<>`` kernel compiled for sm_100a), for GPU A/B parity.
"""
