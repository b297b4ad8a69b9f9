"""ractip_b200 -- B200-native probability stage for RactIP (McCaskill inside/outside,
two-strand co-fold / duplex partition functions, unpaired-window accessibilities).

The arithmetic lives in hand-written sm_100a CUDA kernels behind the C ABI of
include/ractip_prob.h (libractip_prob.so, built in-tree by ractip_b200/build.py).
Importing this package never falls back to a CPU implementation: using the stage
without the built library or without a GPU raises.
"""
from .stage import (DeviceBatch, PairProbabilities, PairRecords, ProbabilityStage, RpError,
                    bp_offsets, default_model, default_opts, zscore_shuffles)

__all__ = ["DeviceBatch", "PairProbabilities", "PairRecords", "ProbabilityStage", "RpError",
           "bp_offsets", "default_model", "default_opts", "zscore_shuffles"]
