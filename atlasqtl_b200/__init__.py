"""atlasqtl_b200: B200-native CUDA implementation of atlasqtl's CAVI sweep hot path.

Host-side mirror of the reference's R interface for that path (`atlasqtl`, `set_hyper`, `set_init`,
`assign_bFDR`) over the C-ABI library `libatlasqtl_b200.so` (include/atlasqtl_b200.h).
There is no CPU fallback: every compute entry point raises if the CUDA library is missing.
"""
from .api import atlasqtl  # noqa: F401
from .hyper_init import set_hyper, set_init  # noqa: F401
from .summarise import assign_bFDR  # noqa: F401
