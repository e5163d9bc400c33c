"""CPU tier for the warp-specialised kernel (soapnuke_b200/csrc/ws_kernel.cuh): its histogram work units (ws_core.cuh:
quality items on owner-computes cells, base items on vertical bit-sliced counters fed by indicator planes, per-item
flush) and its tile shapes, replayed on the CPU by tests/coretest (coretest_filter_ws), must reproduce the oracle bit
for bit - per-read records and every statistics word."""
import os

import numpy as np
import pytest

from helpers import A1, A2, CFG2_KW, abi, assert_same, core_replay_ws, oracle_run, synth

CONFIGS = [
    # name, pe, n, L, gen kwargs, params kwargs, replay kwargs
    ("cfg1_se150_default", False, 20000, 150, dict(seed=1001), dict(), dict()),
    ("cfg2_pe150_all", True, 20000, 150, dict(seed=1002), dict(CFG2_KW, threads=3, patch_size=250), dict()),
    ("cfg2_pe150_all_wpg8", True, 12000, 150, dict(seed=1012), dict(CFG2_KW, threads=2, patch_size=400), dict(wpg=8, grid=7)),
    ("cfg2_pe150_discard_first", True, 10000, 150, dict(seed=1003), dict(adapter1=A1, adapter2=A2), dict(first=123456, wpg=2)),
    ("cfg4_se50_adapter", False, 20000, 50, dict(seed=1004, adapter1=synth.SRNA_ADAPTER3, insert_range=(15, 35)),
     dict(adapter1=synth.SRNA_ADAPTER3.decode(), ada_trim=True, min_read_length=15), dict(grid=5)),
    ("cfg5_pe250_polyg", True, 8000, 250, dict(seed=1005, polyg_frac=0.3), dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10), dict()),
    ("pe120_varlen_hardtrim_headcuts", True, 8000, 120, dict(seed=7, var_len=True),
     dict(adapter1=A1, adapter2=A2, ada_trim=True, hard_trim=(3, 2, 5, 1), polyX_num=8, threads=2, patch_size=50), dict(wpg=1, grid=5)),
    ("pe150_hardtrim_long_headcut", True, 6000, 150, dict(seed=17), dict(adapter1=A1, adapter2=A2, ada_trim=True, hard_trim=(37, 2, 70, 33), min_read_length=20),
     dict()),
    ("pe120_small_qb_checked_tiles", True, 4000, 120, dict(seed=8, var_len=True), dict(adapter1=A1, adapter2=A2, ada_trim=True), dict(qb=8)),
    ("pe100_meanq", True, 4000, 100, dict(seed=9), dict(adapter1=A1, adapter2=A2, mean_quality=20), dict()),
    ("se150_two_adapters_lowercase", False, 6000, 150, dict(seed=10), dict(adapter1=[A2.lower(), A1], ada_trim=True), dict()),
    ("pe150_minlen_off", True, 6000, 150, dict(seed=11), dict(CFG2_KW, min_read_length=-1, max_read_length=140), dict()),
    ("se_tiny_ragged", False, 37, 75, dict(seed=14, var_len=True), dict(adapter1=A1), dict()),
    ("se250_trim", False, 3000, 250, dict(seed=15, var_len=True), dict(adapter1=A1, ada_trim=True, trim_bad_head=(20, 10), trim_bad_tail=(20, 30)), dict()),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_ws_replay_matches_oracle(cfg):
    name, pe, n, L, gkw, pkw, rkw = cfg
    d = synth.gen_pairs(n, L=L, se=not pe, **gkw)
    p = abi.make_params(is_pe=pe, **pkw)
    o1, o2, ost, oerr = oracle_run(p, d, first=rkw.get("first", 0))
    got = core_replay_ws(p, d, **rkw)
    assert got is not None, "shape must be served by the warp-specialised kernel"
    c1, c2, cst, cerr = got
    assert oerr == cerr == 0
    assert_same((c1, c2, cst), (o1, o2, ost), name)


def test_ws_replay_filtersRNA():
    d = synth.gen_srna(12000, L=50, seed=321, var_len=True)
    p = abi.make_params(is_pe=False, srna=True, adapter1=synth.SRNA_ADAPTER5, adapter2=synth.SRNA_ADAPTER3, threads=2, patch_size=100,
                        ada_trim=True, polyG_tail=6, min_read_length=18, max_read_length=49)
    o1, _, ost, oerr = oracle_run(p, d)
    c1, _, cst, cerr = core_replay_ws(p, d, grid=5)
    assert oerr == cerr == 0
    assert_same((c1, None, cst), (o1, None, ost), "filtersRNA")


def test_ws_replay_flushes_before_the_counters_wrap(monkeypatch):
    """Vertical counters and 16-bit cells are flushed on a record budget; a tiny budget forces many flushes."""
    monkeypatch.setenv("SNK_CORETEST_FLUSH_EVERY", "300")
    d = synth.gen_pairs(9000, L=150, seed=33)
    p = abi.make_params(is_pe=True, threads=2, patch_size=500, **CFG2_KW)
    o1, o2, ost, _ = oracle_run(p, d)
    c1, c2, cst, _ = core_replay_ws(p, d, grid=2)
    assert_same((c1, c2, cst), (o1, o2, ost), "frequent flushes")


def test_ws_shape_is_not_offered_for_long_rows():
    d = synth.gen_pairs(200, L=400, seed=12)
    p = abi.make_params(is_pe=True, adapter1=A1, adapter2=A2)
    assert core_replay_ws(p, d) is None
