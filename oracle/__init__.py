"""CPU oracle for the atlasqtl CAVI hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``atlasqtl_b200``)
never does: it fails loudly when its CUDA library is missing instead of falling
back to anything in here.

Contents
--------
``cavi_oracle.c``   C restatement of the sweep (dual form = the reference's loop nest,
                    primal form, blocked primal form).
``ref_wrap.cpp``    extern "C" doorway to the reference's own ``src/coreLoop.cpp``,
                    compiled unmodified against ``shim/RcppEigen.h`` into ``_ref/``.
``native.py``       ctypes loaders for both libraries.
``vb_oracle.py``    NumPy/SciPy restatement of the R outer loop and the ELBO
                    (``R/atlasqtl_global_local_core.R``, ``R/update_vb.R``, ``R/elbo.R``).
``prepare_oracle.py``  literal restatement of ``prepare_data_``'s X / Y work (``scale``, ``rm_constant_``,
                    ``rm_collinear_``; ``R/prepare_atlasqtl.R:57-83``, ``R/utils.R:276-343``).

Parity status: the sweep is pinned against the reference's own C++ (``_ref``); the outer
loop is a restatement with no R available to run -- "parity unpinned" for that part, see
``vb_oracle.py``'s header and DESIGN.md.
"""
