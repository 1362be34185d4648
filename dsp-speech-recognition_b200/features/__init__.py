"""Drop-in replacement of the reference `features` package (AuCson/DSP-Speech-Recognition, features/__init__.py:1-6).

Same sub-modules, same function names, signatures, return types and flat-namespace resolution order (the star
imports below mirror the reference's, so duplicates resolve identically: `features.preemphasis` is
preprocess.preemphasis, `features.center_clip` / `features.max_pitch` are pitch's).  Every arithmetic entry point
runs hand-written sm_100a CUDA kernels through libdspfe.so (include/dspfe.h); there is no CPU fallback.
"""
from .base import *  # noqa: F401,F403
from .sigproc import *  # noqa: F401,F403
from .pitch import *  # noqa: F401,F403
from .endpoint import *  # noqa: F401,F403
from .pitch import *  # noqa: F401,F403
from .preprocess import *  # noqa: F401,F403
