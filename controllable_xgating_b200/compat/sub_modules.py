"""`from sub_modules import *` shim (see compat/SAModel.py)."""
from controllable_xgating_b200.sub_modules import (EncoderLstm_two_fc, Fusion, Gate, LSTMCore_two_layer_gate,  # noqa: F401
                                                   to_contiguous, two_inputs_lstmcell)
