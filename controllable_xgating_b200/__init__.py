"""controllable_xgating_b200 — B200-native (sm_100a) implementation of the gated-fusion caption decoder
of vsislab/Controllable_XGating (caption_src/SAModel.py + sub_modules.py + CaptionModel.py).

Public surface = the reference's: SAModel, CaptionModel, LanguageModelCriterion, ClassiferCriterion,
RewardCriterion, to_contiguous.  All arithmetic runs in libxgating.so (include/xgating.h); importing this
package without the built library raises.
"""
from . import _lib

_lib.load()   # fail loudly at import if the CUDA library has not been built

from .CaptionModel import CaptionModel  # noqa: E402
from .SAModel import (ClassiferCriterion, LanguageModelCriterion, RewardCriterion, SAModel,  # noqa: E402
                      to_contiguous)

from .optim import FusedAdam  # noqa: E402

__all__ = ["SAModel", "CaptionModel", "LanguageModelCriterion", "ClassiferCriterion", "RewardCriterion",
           "to_contiguous", "FusedAdam"]
__version__ = "0.1.0"
