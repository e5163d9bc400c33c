"""CPU tier: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol
include/snk_engine.h declares. No compute entry point is exercised here (they need a device and
there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT, abi


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "snk_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snk_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_mirror_agree():
    assert declared_symbols() == abi.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol(engine_lib):
    for name in declared_symbols():
        assert hasattr(engine_lib, name), f"{name} missing from libsnk_engine.so"
    assert engine_lib.snk_abi_version() == abi.ABI_VERSION
    assert engine_lib.snk_stats_slot_words() == abi.SLOT_WORDS


def test_struct_sizes_match_the_header(tmp_path):
    """sizeof/offsetof as the C compiler sees them == the ctypes mirror."""
    import subprocess
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "snk_engine.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(snk_params), sizeof(snk_batch), sizeof(snk_read_result),'
                    ' offsetof(snk_params, adapter), offsetof(snk_params, slot_block), offsetof(snk_params, n_slots),'
                    ' sizeof(snk_text_format), sizeof(snk_text_meta), offsetof(snk_text_meta, flags));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.Params), C.sizeof(abi.Batch), C.sizeof(abi.ReadResult), abi.Params.adapter.offset,
            abi.Params.slot_block.offset, abi.Params.n_slots.offset, C.sizeof(abi.TextFormat), C.sizeof(abi.TextMeta),
            abi.TextMeta.flags.offset]
    assert got == want


def test_params_check_rejects_what_the_reference_would_crash_on(engine_lib):
    p = abi.make_params(is_pe=True, adapter1="ACGTACGTACGT")
    assert engine_lib.snk_params_check(C.byref(p)) == 0
    p.ada_mis[0] = -1                      # (adptLen-5)/(adaMis+1): division by zero in read_filter.cpp:714
    assert engine_lib.snk_params_check(C.byref(p)) != 0
    assert b"adaMis" in engine_lib.snk_last_error()
    p = abi.make_params(is_pe=True)
    p.n_slots = 0
    assert engine_lib.snk_params_check(C.byref(p)) != 0


def test_engine_refuses_to_run_without_a_gpu(engine_lib):
    """No silent CPU path: creating an engine on a box without CUDA must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = abi.make_params(is_pe=True)
    h = C.c_void_p()
    assert engine_lib.snk_engine_create(C.byref(p), 0, C.byref(h)) != 0
    assert b"no CUDA device" in engine_lib.snk_last_error() or b"CUDA" in engine_lib.snk_last_error()


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under soapnuke_b200/ may reference it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "soapnuke_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_py|liboracle|snk_oracle|orc_filter|libcoretest", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
