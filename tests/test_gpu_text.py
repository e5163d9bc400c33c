"""GPU tier of the FASTQ text path, through the C ABI: raw FASTQ text in,
clean FASTQ text out. The device must index and pack the text into the rows the SoA entry point
would have been given (same per-read records, same statistics as the oracle) and emit exactly the
bytes the reference's output_fastqs writes (model: helpers.ref_clean_text, pinned end to end
against the reference binary by tests/test_cli_gpu.py)."""
import ctypes as C

import numpy as np
import pytest

from helpers import A1, A2, CFG2_KW, Engine, abi, assert_same, fastq_text, oracle_run, ref_clean_text, synth
from test_text_replay import weird_ids

pytestmark = pytest.mark.gpu


def texts_of(d, ids, mates, **kw):
    return [fastq_text(ids[m], d[f"seq{m + 1}"], d[f"qual{m + 1}"], d[f"len{m + 1}"], **kw) for m in range(mates)]


@pytest.mark.parametrize("opts", [dict(), dict(pe_info=1), dict(fasta=1, pe_info=1, id_mode=1), dict(id_mode=2), dict(id_mode=1)],
                         ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()) or "plain")
def test_text_path_matches_oracle_and_reference_format(engine_lib, opts):
    n = 40000
    d = synth.gen_pairs(n, L=150, seed=1002)
    p = abi.make_params(is_pe=True, threads=3, patch_size=1000, **CFG2_KW)
    o1, o2, ost, oerr = oracle_run(p, d)
    ids = [weird_ids(n, 0), weird_ids(n, 1)]
    texts = texts_of(d, ids, 2)
    with Engine(engine_lib, p) as e:
        meta, outs, offs, ress = e.filter_text(texts, n, 160, **opts)
        st = e.stats()
        flags, _ = e.error_flags()
    assert meta.flags == 0 and flags == oerr == 0
    assert meta.max_len == 150 and meta.kept == int((o1["category"] == 0).sum())
    assert_same((ress[0], ress[1], st), (o1, o2, ost), "text path")
    for m, res in ((0, o1), (1, o2)):
        want, want_off = ref_clean_text(ids[m], d[f"seq{m + 1}"], d[f"qual{m + 1}"], res, m, **opts)
        assert np.array_equal(offs[m], want_off)
        assert outs[m] == want, f"clean text of mate {m + 1} differs"


def test_text_path_se_crlf_phred64_varlen(engine_lib):
    n = 30000
    d = synth.gen_pairs(n, L=100, seed=9, se=True, var_len=True)
    d["qual1"] = np.where(d["qual1"] > 0, d["qual1"] + 31, 0).astype(np.uint8)        # phred 64 input
    p = abi.make_params(is_pe=False, adapter1=A1, ada_trim=True, quality_phred=64, out_quality_phred=33, polyG_tail=10)
    o1, _, ost, oerr = oracle_run(p, d)
    ids = [synth.read_ids(n, 1)]
    texts = texts_of(d, ids, 1, eol=b"\r\n")
    with Engine(engine_lib, p) as e:
        meta, outs, offs, ress = e.filter_text(texts, n, 112, strip=2)
        st = e.stats()
    assert meta.flags == 0 and oerr == 0
    assert_same((ress[0], None, st), (o1, None, ost), "se text path")
    want, want_off = ref_clean_text(ids[0], d["seq1"], d["qual1"], o1, 0, qshift=-31)
    assert np.array_equal(offs[0], want_off) and outs[0] == want


def test_text_path_batches_lanes_and_slots(engine_lib):
    """Several text batches on rotating lanes with a running first_index == one oracle run."""
    n = 50000
    d = synth.gen_pairs(n, L=100, seed=21)
    p = abi.make_params(is_pe=True, threads=4, patch_size=20, **CFG2_KW)
    o1, o2, ost, _ = oracle_run(p, d)
    ids = [synth.read_ids(n, 1), synth.read_ids(n, 2)]
    cuts = [0, 1, 1000, 1003, 24200, 50000]
    with Engine(engine_lib, p) as e:
        for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            sub = {key: np.ascontiguousarray(v[a:b]) for key, v in d.items() if isinstance(v, np.ndarray)}
            texts = texts_of(sub, [ids[0][a:b], ids[1][a:b]], 2, last_newline=True)
            meta, outs, offs, ress = e.filter_text(texts, b - a, 112, first=a, lane=k % 3)
            assert meta.flags == 0
            assert np.array_equal(ress[0], o1[a:b]) and np.array_equal(ress[1], o2[a:b])
            want, _ = ref_clean_text(ids[1][a:b], sub["seq2"], sub["qual2"], o2[a:b], 1)
            assert outs[1] == want
        st = e.stats()
    assert_same((o1, o2, st), (o1, o2, ost), "composed text batches")


def test_text_path_flags_and_stride_retry(engine_lib):
    n = 5000
    d = synth.gen_pairs(n, L=150, seed=4, se=True)
    p = abi.make_params(is_pe=False)
    o1, _, ost, _ = oracle_run(p, d)
    ids = [synth.read_ids(n, 1)]
    text = texts_of(d, ids, 1)[0]
    with Engine(engine_lib, p) as e:
        meta, *_ = e.filter_text([text], n, 64)                       # rows too narrow: nothing may be counted
        assert meta.flags == abi.TEXT_STRIDE_OVERFLOW and meta.max_len == 150
        assert not e.stats().any()
        meta, *_ = e.filter_text([text], n - 1, 160)                  # more lines than 4n
        assert meta.flags & abi.TEXT_LINE_COUNT
        assert not e.stats().any()
        meta, *_ = e.filter_text([text[:-1]], n, 160)                 # last line without newline loses a character
        assert meta.flags & abi.TEXT_LEN_MISMATCH and meta.bad_record == n - 1
        e.check(e.lib.snk_engine_stats_reset(e.h))
        meta, outs, offs, ress = e.filter_text([text], n, 160)        # the retry with the right stride
        assert meta.flags == 0
        assert_same((ress[0], None, e.stats()), (o1, None, ost), "after retry")


def test_text_path_many_segments_equals_soa_path(engine_lib):
    """A batch whose text spans hundreds of 32 KB segments: same records and statistics as the SoA entry point."""
    n = 300000
    d = synth.gen_pairs(n, L=150, seed=1002)
    p = abi.make_params(is_pe=True, threads=8, patch_size=0, **CFG2_KW)
    ids = [synth.read_ids(n, 1), synth.read_ids(n, 2)]
    texts = texts_of(d, ids, 2)
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d)
        st_soa = e.stats()
    with Engine(engine_lib, p) as e:
        meta, outs, offs, ress = e.filter_text(texts, n, 160)
        st = e.stats()
    assert meta.flags == 0
    assert_same((ress[0], ress[1], st), (r1, r2, st_soa), "text vs SoA")
    for m, res in ((0, r1), (1, r2)):
        assert outs[m] == synth.clean_fastq_bytes(d[f"seq{m + 1}"], d[f"qual{m + 1}"], d[f"len{m + 1}"], res, m + 1)


def test_text_path_double_pe_suffix_with_nothing_dropped(engine_lib):
    """pe_info = 2 (the reference's double "/1/1" suffix when the trim files are on, peprocess.cpp:1460-1475) adds 4 bytes
    per kept record: with lenient filters and no drops the clean text is LONGER than the raw text, and both mates' outputs
    must still come back intact (the device output buffer is sized for it)."""
    n = 60000
    d = synth.gen_pairs(n, L=150, seed=31)
    p = abi.make_params(is_pe=True, low_qual_ratio=-1, n_ratio=-1, min_read_length=1, threads=2, patch_size=1000)
    o1, o2, ost, oerr = oracle_run(p, d)
    assert int((o1["category"] == 0).sum()) == n, "case must keep every pair"
    ids = [synth.read_ids(n, 1), synth.read_ids(n, 2)]
    texts = texts_of(d, ids, 2)
    with Engine(engine_lib, p) as e:
        meta, outs, offs, ress = e.filter_text(texts, n, 160, pe_info=2)
        st = e.stats()
    assert meta.flags == 0 and meta.kept == n
    assert meta.out_bytes[0] == len(texts[0]) + 4 * n and meta.out_bytes[1] == len(texts[1]) + 4 * n
    assert_same((ress[0], ress[1], st), (o1, o2, ost), "pe_info 2")
    for m, res in ((0, o1), (1, o2)):
        want, want_off = ref_clean_text(ids[m], d[f"seq{m + 1}"], d[f"qual{m + 1}"], res, m, pe_info=2)
        assert np.array_equal(offs[m], want_off)
        assert outs[m] == want, f"clean text of mate {m + 1} differs"
    with Engine(engine_lib, p) as e:                       # the ABI rejects values the formatter has no room for
        fmt = abi.TextFormat(strip=1, pe_info=3)
        buf = np.frombuffer(texts[0], dtype=np.uint8).copy()
        rc = e.lib.snk_filter_pe_text_async(e.h, 0, buf.ctypes.data, len(texts[0]), buf.ctypes.data, len(texts[0]), n, 160, C.byref(fmt), 0)
        assert rc != 0 and b"bad text format" in e.lib.snk_last_error()
