"""Synthetic checkpoints for the oracle (TEST INFRASTRUCTURE): the generator itself lives in tools/synth_weights.py so
that bench.py's in-run parity check can build the same seeded weights without importing anything from oracle/."""
from tools.synth_weights import *  # noqa: F401,F403
from tools.synth_weights import CONFIGS, MOE_CONFIGS, checksum, iter_state_dict, synth_state_dict, synth_tensor, tensor_specs  # noqa: F401
