"""soapnuke_b200 — B200-native FASTQ filter engine (drop-in for the SOAPnuke `filter` hot path).

The product is the CUDA/C++ shared library built from csrc/ + host/ (see include/snk_engine.h);
this Python package only holds the ctypes mirror of that ABI, the synthetic data generator and
the build helper used by tests and bench.py.
"""
from . import abi, synth  # noqa: F401
