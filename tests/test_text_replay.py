"""CPU tier of the FASTQ text path: the device functions of
soapnuke_b200/csrc/text_core.cuh compiled as plain C++ (tests/coretest, test only) must index and
pack FASTQ text into exactly the rows the generator made, and format the clean records exactly as
the reference's output_fastqs / preOutput / index removal do (test-side model in helpers.py, which
tests/test_cli_gpu.py pins against the reference binary end to end)."""
import numpy as np
import pytest

from helpers import (A1, A2, CFG2_KW, abi, fastq_text, oracle_run, ref_clean_text, ref_id_transform, synth,
                     text_replay_format, text_replay_index_pack)


def weird_ids(n, mate):
    rng = np.random.default_rng(77 + mate)
    out = []
    for i in range(n):
        k = rng.integers(0, 8)
        base = b"@FCD1PB1ACXX:4:1101:%d:%d" % (i // 1000, i % 1000)
        if k == 0:
            rid = base + b"#GAAGCACG/%d" % (mate + 1)
        elif k == 1:
            rid = base + b"#AC#GT/x/%d#tail" % (mate + 1)
        elif k == 2:
            rid = b"@noseparators%d" % i
        elif k == 3:
            rid = b"@HISEQ:310:C5MH9ANXX:1:1101:%d:2043 %d:N:0:TCGGTCAC" % (i, mate + 1)
        elif k == 4:
            rid = b"X@" + base[1:] + b"/%d" % (mate + 1)      # '@' not in front
        elif k == 5:
            rid = b"@" + bytes(rng.integers(33, 127, size=int(rng.integers(1, 90)), dtype=np.uint8))
        elif k == 6:
            rid = b"@"
        else:
            rid = base + b"/%d" % (mate + 1)
        out.append(rid)
    return out


@pytest.mark.parametrize("eol,strip,last_nl", [(b"\n", 1, True), (b"\r\n", 2, True), (b"\n", 1, False)])
@pytest.mark.parametrize("L,var", [(150, False), (50, True), (250, True)])
def test_index_and_pack_rebuild_the_rows(eol, strip, last_nl, L, var):
    d = synth.gen_pairs(700, L=L, seed=L + strip, se=True, var_len=var)
    ids = weird_ids(700, 0)
    text = fastq_text(ids, d["seq1"], d["qual1"], d["len1"], eol=eol, last_newline=last_nl)
    stride = d["seq1"].shape[1]
    flags, line_off, S, Q, Ln, mx, _ = text_replay_index_pack(text, 700, strip, stride)
    want_S, want_Q = d["seq1"].copy(), d["qual1"].copy()
    if not last_nl:            # the reference strips one character from every line, newline or not:
        want_Q[-1, int(d["len1"][-1]) - 1] = 0      # the last quality line loses a real character
        assert flags == 2      # ... so the last record's sequence and quality lengths differ
    else:
        assert flags == 0
    assert np.array_equal(Ln, d["len1"])
    assert mx == d["len1"].max()
    assert np.array_equal(S, want_S) and np.array_equal(Q, want_Q)
    # ids are where line_off says
    for i in (0, 1, 350, 699):
        a, b = int(line_off[4 * i]), int(line_off[4 * i + 1])
        assert text[a:b - strip] == ids[i] or (i == 699 and not last_nl)


def test_pack_flags():
    d = synth.gen_pairs(64, L=100, seed=3, se=True)
    ids = synth.read_ids(64, 1)
    text = fastq_text(ids, d["seq1"], d["qual1"], d["len1"])
    assert text_replay_index_pack(text, 64, 1, 64)[0] & 1                     # stride overflow
    assert text_replay_index_pack(text, 63, 1, 112)[0] == 4                   # more lines than 4n
    assert text_replay_index_pack(text[:-200], 64, 1, 112)[0] == 4            # fewer lines
    bad = text.replace(b"\n+\n", b"\n+\nI", 1)
    assert text_replay_index_pack(bad, 64, 1, 112)[0] == 2                    # seq/qual length mismatch


@pytest.mark.parametrize("opts", [dict(), dict(pe_info=1), dict(fasta=1), dict(fasta=1, pe_info=1, id_mode=1), dict(id_mode=1),
                                  dict(id_mode=2, pe_info=1), dict(qshift=31), dict(id_mode=2, fasta=1), dict(pe_info=2), dict(pe_info=2, fasta=1, id_mode=1)],
                         ids=lambda o: "-".join(f"{k}{v}" for k, v in o.items()) or "plain")
@pytest.mark.parametrize("lanes", [1, 32])
def test_format_matches_reference_model(opts, lanes):
    n = 1500
    d = synth.gen_pairs(n, L=150, seed=5)
    p = abi.make_params(is_pe=True, **CFG2_KW)
    r1, r2, _, err = oracle_run(p, d)
    assert err == 0 and 0 < (r1["category"] == 0).sum() < n
    for mate, res in ((0, r1), (1, r2)):
        ids = weird_ids(n, mate)
        S, Q, Ln = d[f"seq{mate + 1}"], d[f"qual{mate + 1}"], d[f"len{mate + 1}"]
        text = fastq_text(ids, S, Q, Ln)
        flags, line_off, PS, PQ, PL, _, buf = text_replay_index_pack(text, n, 1, S.shape[1])
        assert flags == 0 and np.array_equal(PS, S)
        got, got_off = text_replay_format(buf, len(text), line_off, PS, PQ, res, mate, 1, lanes=lanes, **opts)
        want, want_off = ref_clean_text(ids, S, Q, res, mate, **opts)
        assert np.array_equal(got_off, want_off)
        assert got == want


def test_id_transform_model_examples():
    assert ref_id_transform(b"@FCD1PB1ACXX:4:1101:1799:2201#GAAGCACG/2", 1) == b"@FCD1PB1ACXX:4:1101:1799:2201/2"
    assert ref_id_transform(b"@HISEQ:310:C5MH9ANXX:1:1101:3517:2043 2:N:0:TCGGTCAC", 2) == b"@HISEQ:310:C5MH9ANXX:1:1101:3517:2043 2:N:0"
