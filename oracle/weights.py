"""Synthetic weights/inputs used by the oracle tests: re-exported from counting_detr_b200.synthetic so that the
reference (when goldens are generated), the oracle and the CUDA path all consume the very same tensors."""
from counting_detr_b200.synthetic import (SynthCfg, default_args, make_inputs, make_state_dict,  # noqa: F401
                                          uniform)
