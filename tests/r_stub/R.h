/* TEST INFRASTRUCTURE ONLY: stand-in declarations for the handful of R C-API names bindings/R/atlasqtl_b200_shim.c uses,
 * so that the shim can be syntax- and type-checked (`gcc -fsyntax-only`, tests/test_cabi.py) and -- together with the
 * miniature runtime rstub_runtime.c -- compiled, loaded and executed (tests/r_shim_real.py) in an image without R.
 * Not R's headers. */
#ifndef AQ_R_STUB_H
#define AQ_R_STUB_H
#include <stddef.h>
#endif
