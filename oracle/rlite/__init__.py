"""rlite: a small evaluator for the subset of R the reference's R/*.R files use -- TEST INFRASTRUCTURE ONLY.

Why it exists: the outer VB loop, the ELBO, the pre-processing and the bFDR summary of the reference are R code, and
this image has no R.  Instead of trusting a hand restatement, the reference's unmodified sources are parsed and
executed here (parser.py: tokeniser + Pratt parser; interp.py: environments, closures, lazy defaults, replacement
calls; base.py: the ~150 base / stats / gsl functions those files call, NumPy / SciPy underneath; reference.py: loads
/root/reference/R/*.R and binds `.Call` to the reference's compiled src/coreLoop.cpp).  The outputs are committed as
tests/golden/rlite_*.npz and pin oracle/vb_oracle.py, the product's host loop and the CUDA path
(tests/test_rlite.py, tests/test_gpu_rlite_golden.py).

Only tests/ and tests/golden/make_rlite_golden.py may import this package; the product never does
(tests/test_cabi.py::test_product_never_imports_the_oracle).
"""
