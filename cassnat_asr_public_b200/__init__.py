"""cassnat_asr_public_b200 — B200-native (sm_100a) CASS-NAT alignment-conditioned decoding path.

Python host mirror of the reference's `src/models/cassnat.py` API over a C-ABI CUDA library
(include/cassnat_b200.h).  No Triton, no CPU fallback: the ops raise without the built library
and a CUDA device.
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
