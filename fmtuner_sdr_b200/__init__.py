"""fmtuner_sdr_b200 — B200-native channel-batched FM stereo + RDS engine.

Python is plumbing only: `Engine` is a ctypes view of the C ABI in include/fmgpu.h
(libfmgpu.so, hand-written sm_100a CUDA). There is no CPU fallback; creating an
engine without a CUDA device raises.
"""
from .engine import (RESET_AFPOST, RESET_ALL, RESET_DECIM, RESET_DEMOD, RESET_DSP, RESET_RDS, RESET_STEREO,  # noqa: F401
                     Channelizer, Engine, EngineError, LevelSums, SignalLevel, SynthParams, GROUP_DTYPE, STATUS_DTYPE, XdrRdsFormatter, lib_path,  # noqa: F401
                     load_library, make_config, measure_fp32_tflops, synth_iq)
