"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see snk_oracle.c).
Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs; never from the product."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "SOAPnuke")

import sys
sys.path.insert(0, os.path.dirname(HERE))
from soapnuke_b200 import abi  # noqa: E402  (POD definitions only)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "snk_oracle.c")):
            build()
        L = C.CDLL(LIB)
        L.orc_adapter_pos.restype = C.c_int
        L.orc_adapter_pos.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_int]
        L.orc_filter_pe.restype = C.c_int
        L.orc_id_flags.restype = C.c_uint32
        L.orc_id_flags.argtypes = [C.POINTER(abi.Params), C.c_char_p, C.c_int]
        L.orc_filter_pe.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Batch), C.POINTER(abi.Batch),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32)]
        L.orc_filter_se.restype = C.c_int
        L.orc_filter_se.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Batch), C.c_void_p, C.c_void_p,
                                    C.c_uint64, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def adapter_pos(read: bytes, adapter: bytes, ada_mis=2, ada_mr=0.5, ada_edge=6):
    return lib().orc_adapter_pos(read, len(read), adapter, len(adapter), ada_mis, ada_mr, ada_edge)


def id_flags(params, ids):
    """SNK_PRE_TILE / SNK_PRE_FOV bits (uint16 array) of the record ids (list of bytes), to be OR-ed into len[]."""
    return np.array([lib().orc_id_flags(C.byref(params), i, len(i)) for i in ids], dtype=np.uint16)


def new_stats(params):
    return np.zeros(params.n_slots * abi.SLOT_WORDS, dtype=np.uint64)


def filter_pe(params, d, stats=None, first_index=0):
    """d: dict from synth.gen_pairs. Returns (res1, res2, stats, err)."""
    n = d["seq1"].shape[0]
    b1 = abi.make_batch(d["seq1"], d["qual1"], d["len1"])
    b2 = abi.make_batch(d["seq2"], d["qual2"], d["len2"])
    r1 = np.zeros(n, dtype=abi.RESULT_DTYPE)
    r2 = np.zeros(n, dtype=abi.RESULT_DTYPE)
    if stats is None:
        stats = new_stats(params)
    err = C.c_uint32(0)
    rc = lib().orc_filter_pe(C.byref(params), C.byref(b1), C.byref(b2), r1.ctypes.data, r2.ctypes.data,
                             stats.ctypes.data, first_index, C.byref(err))
    assert rc == 0
    return r1, r2, stats, err.value


def filter_se(params, d, stats=None, first_index=0):
    n = d["seq1"].shape[0]
    b1 = abi.make_batch(d["seq1"], d["qual1"], d["len1"])
    r1 = np.zeros(n, dtype=abi.RESULT_DTYPE)
    if stats is None:
        stats = new_stats(params)
    err = C.c_uint32(0)
    rc = lib().orc_filter_se(C.byref(params), C.byref(b1), r1.ctypes.data, stats.ctypes.data, first_index, C.byref(err))
    assert rc == 0
    return r1, stats, err.value


def have_reference():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(args, cwd=None, timeout=600, module="filter"):
    """Run the unmodified reference binary: `SOAPnuke <module> <args>` (filter | filtersRNA)."""
    return subprocess.run([REF_BIN, module] + list(args), cwd=cwd, timeout=timeout,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE)
