"""Put this directory first on sys.path and the reference's `starttrain.py` / `eval.py` / `eval_utils.py`
(`from SAModel import *`) pick up the B200 implementation unchanged."""
import numpy as np  # noqa: F401  (the reference's SAModel star-exports these via data_io / torch imports)
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.autograd import Variable  # noqa: F401

from controllable_xgating_b200.SAModel import *  # noqa: F401,F403
from controllable_xgating_b200.SAModel import (ClassiferCriterion, LanguageModelCriterion, RewardCriterion,  # noqa: F401
                                               SAModel, to_contiguous)
