"""CPU tier of the kernel parity tests: the device functions of soapnuke_b200/csrc/filter_core.cuh,
compiled as plain C++ and driven through the kernel's tile/phase structure (tests/coretest), must
reproduce the oracle bit for bit (per-read records and every statistics word)."""
import numpy as np
import pytest

from helpers import A1, A2, CFG2_KW, abi, assert_same, core_replay, oracle_run, synth

CONFIGS = [
    # name, pe, n, L, gen kwargs, params kwargs, replay kwargs
    ("cfg1_se150_default", False, 20000, 150, dict(seed=1001), dict(), dict()),
    ("cfg2_pe150_all", True, 20000, 150, dict(seed=1002), dict(CFG2_KW, threads=3, patch_size=250), dict()),
    ("cfg2_pe150_discard", True, 10000, 150, dict(seed=1003), dict(adapter1=A1, adapter2=A2), dict(first=123456, tile_r=32)),
    ("cfg4_se50_adapter", False, 20000, 50, dict(seed=1004, adapter1=synth.SRNA_ADAPTER3, insert_range=(15, 35)),
     dict(adapter1=synth.SRNA_ADAPTER3.decode(), ada_trim=True, min_read_length=15), dict(grid=5)),
    ("cfg5_pe250_polyg", True, 8000, 250, dict(seed=1005, polyg_frac=0.3), dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10), dict()),
    ("pe120_varlen_hardtrim", True, 8000, 120, dict(seed=7, var_len=True),
     dict(adapter1=A1, adapter2=A2, ada_trim=True, hard_trim=(3, 2, 5, 1), polyX_num=8, threads=2, patch_size=50), dict(tile_r=64, grid=5)),
    ("pe120_small_qb", True, 4000, 120, dict(seed=8, var_len=True), dict(adapter1=A1, adapter2=A2, ada_trim=True), dict(qb=8)),
    ("pe100_phred64_out33", True, 4000, 100, dict(seed=9), dict(adapter1=A1, adapter2=A2, mean_quality=20), dict()),
    ("se150_two_adapters_lowercase", False, 6000, 150, dict(seed=10), dict(adapter1=[A2.lower(), A1], ada_trim=True), dict()),
    ("pe150_minlen_off", True, 6000, 150, dict(seed=11), dict(CFG2_KW, min_read_length=-1, max_read_length=140), dict()),
    ("pe400_long_reads", True, 1500, 400, dict(seed=12), dict(adapter1=A1, adapter2=A2, ada_trim=True), dict(tile_r=32)),
    # stride 1008: the CTA is capped at 1024 threads -> one phase-B unit per item instead of two
    ("pe1000_max_len", True, 300, 1000, dict(seed=13), dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10, trim_bad_head=(20, 10)), dict()),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_core_replay_matches_oracle(cfg):
    name, pe, n, L, gkw, pkw, rkw = cfg
    d = synth.gen_pairs(n, L=L, se=not pe, **gkw)
    p = abi.make_params(is_pe=pe, **pkw)
    o1, o2, ost, oerr = oracle_run(p, d, first=rkw.get("first", 0))
    c1, c2, cst, cerr = core_replay(p, d, **rkw)
    assert oerr == cerr == 0
    assert_same((c1, c2, cst), (o1, o2, ost), name)


SRNA_CONFIGS = [
    ("srna_discard_L44", 12000, 44, dict(), dict()),
    ("srna_trim_polyg_varlen", 12000, 50, dict(var_len=True), dict(ada_trim=True, polyG_tail=6, highA_ratio=0.6, polyX_num=12)),
    ("srna_trim_hard_params", 8000, 75, dict(), dict(ada_trim=True, hard_trim=(2, 1), min_read_length=15, max_read_length=60,
                                                     ada_rctg=7, ada_rar=0.7, ada_rma=6, ada_rer=0.3, ada_rmm=3)),
    ("srna_minlen_off", 6000, 50, dict(var_len=True), dict(ada_trim=True, min_read_length=-1, max_read_length=-1)),
]


def test_filtersRNA_other_adapter_shapes():
    """Adapters with an N or longer than 64 bases cannot use the bit-plane aligners (byte loops);
    adapters of 33..64 bases use the 64-bit windows."""
    d = synth.gen_srna(6000, L=50, seed=77, var_len=True)
    a5 = synth.SRNA_ADAPTER5[:10] + b"N" + synth.SRNA_ADAPTER5[11:]
    a3 = synth.SRNA_ADAPTER3[:14] + b"N" + synth.SRNA_ADAPTER3[15:]
    for ada5, ada3 in ((a5, a3), (synth.SRNA_ADAPTER5 * 3, synth.SRNA_ADAPTER3),
                       (synth.SRNA_ADAPTER5[:6] + synth.SRNA_ADAPTER5 + synth.SRNA_ADAPTER5[:14], synth.SRNA_ADAPTER3 + synth.SRNA_ADAPTER3)):
        p = abi.make_params(is_pe=False, srna=True, adapter1=ada5, adapter2=ada3, ada_trim=True, min_read_length=18, max_read_length=49)
        o1, _, ost, oerr = oracle_run(p, d)
        c1, _, cst, cerr = core_replay(p, d)
        assert oerr == cerr == 0
        assert (o1["category"] == 0).sum() > 500
        assert_same((c1, None, cst), (o1, None, ost), "srna byte path")


@pytest.mark.parametrize("cfg", SRNA_CONFIGS, ids=[c[0] for c in SRNA_CONFIGS])
def test_core_replay_matches_oracle_filtersRNA(cfg):
    """filtersRNA: sRNA_findAdapter / sRNA_hasAdapter, the cut at the 3' adapter and sRNA_discard."""
    name, n, L, gkw, pkw = cfg
    d = synth.gen_srna(n, L=L, seed=len(name) * 31 + L, **gkw)
    kw = dict(min_read_length=18, max_read_length=49)
    kw.update(pkw)
    p = abi.make_params(is_pe=False, srna=True, adapter1=synth.SRNA_ADAPTER5, adapter2=synth.SRNA_ADAPTER3, threads=2, patch_size=100, **kw)
    o1, _, ost, oerr = oracle_run(p, d)
    c1, _, cst, cerr = core_replay(p, d, grid=5)
    assert oerr == cerr == 0
    assert (np.bincount(o1["category"], minlength=12) > 0).sum() >= 4
    assert_same((c1, None, cst), (o1, None, ost), name)


@pytest.mark.parametrize("pe", [True, False], ids=["pe", "se"])
def test_tile_fov_flags_in_len(pe):
    """SNK_PRE_TILE / SNK_PRE_FOV bits of len[] (the ids never cross the SoA boundary): first test of the
    discard cascade, plain counters, only mate 1's flags count for a pair."""
    import oracle_py as orc
    n = 8000
    d = synth.gen_pairs(n, L=100, seed=61, se=not pe, var_len=True)
    p = abi.make_params(is_pe=pe, adapter1=A1, adapter2=A2 if pe else None, ada_trim=True, tile="1102,2201", fov="C002R003", threads=2, patch_size=50)
    ids = [a if i % 3 else b for i, (a, b) in enumerate(zip(synth.tile_ids(n, 1), synth.fov_ids(n, 1)))]
    d["len1"] = d["len1"] | orc.id_flags(p, ids)
    if pe:
        d["len2"] = d["len2"] | np.where(np.arange(n) % 11 == 0, abi.PRE_TILE, 0).astype(np.uint16)     # mate 2's flags are ignored
    o1, o2, ost, oerr = oracle_run(p, d)
    c1, c2, cst, cerr = core_replay(p, d, grid=5)
    assert oerr == cerr == 0
    cats = np.bincount(o1["category"], minlength=14)
    assert cats[12] > 200 and cats[13] > 50 and cats[0] > 2000, cats
    assert_same((c1, c2, cst), (o1, o2, ost), "tile/fov")


CONTAM_PLANTS = [synth.CONTAM1, synth.CONTAM2, synth.CONTAM3, synth.revcomp(synth.CONTAM1), synth.revcomp(synth.CONTAM3)]
CONTAM_CONFIGS = [
    ("contam_pe_single", True, 6000, 100, dict(), dict(adapter1=A1, adapter2=A2, ada_trim=True, contam1=synth.CONTAM1.decode(), contam2=synth.CONTAM2.decode())),
    ("contam_se_list_varlen", False, 6000, 120, dict(var_len=True),
     dict(contam1=f"{synth.CONTAM1.decode()},{synth.CONTAM3.decode()}", ct_match_r="0.3,0.5")),
    ("contam_pe_list_budgets", True, 5000, 150, dict(),
     dict(adapter1=A1, adapter2=A2, contam1=",".join(c.decode() for c in (synth.CONTAM3, synth.CONTAM2, synth.CONTAM1)),
          contam2=",".join([synth.CONTAM1.decode(), synth.CONTAM2.decode(), synth.CONTAM3.decode()[:20]]), ct_match_r="0.2,0.6,0.9",
          ada_mis=(1, 3), ada_edge=(4, 8))),
    ("contam_short_reads_tiny_thr", True, 4000, 40, dict(), dict(contam1=synth.CONTAM1.decode(), contam2=synth.CONTAM2.decode(), ct_match_r="0.1",
                                                                  min_read_length=20)),     # reads shorter than the contaminant, segThr-6 <= 0
    ("contam_trim_mode", True, 3000, 100, dict(), dict(adapter1=A1, adapter2=A2, ada_trim=True, contam1=synth.CONTAM1.decode(),
                                                       contam2=synth.CONTAM2.decode(), ct_match_r="0.4", contam_trim=True)),
]


_C1, _C2, _C3 = synth.CONTAM1.decode(), synth.CONTAM2.decode(), synth.CONTAM3.decode()
CONTAM_CONFIGS += [
    ("gcontam_pe", True, 5000, 100, dict(), dict(adapter1=A1, adapter2=A2, ada_trim=True, global_contams=f"{_C1},{_C3}", glob_cotm_mR="0.5,0.6", glob_cotm_mM="1,2")),
    ("gcontam_se_with_contam_varlen", False, 5000, 120, dict(var_len=True), dict(global_contams=_C2, glob_cotm_mR="0.4", glob_cotm_mM="0", contam1=_C1)),
    ("gcontam_pe_short_reads", True, 4000, 60, dict(), dict(global_contams=f"{_C3},{_C1[:18]},{_C2}", glob_cotm_mR="0.5,1.0,0.9", glob_cotm_mM="1,0,2", contam2=_C2)),
    ("gcontam_reads_shorter_than_contam", True, 3000, 40, dict(var_len=True), dict(global_contams=f"{_C1}{_C2},{_C3}", glob_cotm_mR="0.3,0.5", glob_cotm_mM="0,1",
                                                                                      min_read_length=10)),
]


@pytest.mark.parametrize("cfg", CONTAM_CONFIGS, ids=[c[0] for c in CONTAM_CONFIGS])
def test_core_replay_matches_oracle_contam(cfg):
    """Contaminant sequences (hasContam / hasContams): per-offset budget / run tables, N handling, discard order."""
    name, pe, n, L, gkw, pkw = cfg
    d = synth.add_contams(synth.gen_pairs(n, L=L, seed=len(name) * 13, se=not pe, **gkw), CONTAM_PLANTS, seed=L)
    p = abi.make_params(is_pe=pe, threads=2, patch_size=60, **pkw)
    o1, o2, ost, oerr = oracle_run(p, d)
    c1, c2, cst, cerr = core_replay(p, d, grid=4)
    assert oerr == cerr == 0
    if not pkw.get("contam_trim"):
        assert ((o1["category"] == 14) | (o1["category"] == 15)).sum() > n // 50
    assert_same((c1, c2, cst), (o1, o2, ost), name)


def test_mixed_checked_and_unchecked_tiles(monkeypatch):
    """A few records with qualities above the shared-memory bins: their tiles take the checked
    histogram path (out-of-bin qualities go straight to the global tables, mirrored into the clean
    table and taken back by delta entries), all other tiles the unchecked walk; head and tail trims
    on, and a small flush period so that raw/delta cells are flushed many times."""
    from helpers import high_quality_mix
    monkeypatch.setenv("SNK_CORETEST_FLUSH_EVERY", "300")
    d = high_quality_mix()
    p = abi.make_params(is_pe=True, adapter1=A1, adapter2=A2, ada_trim=True, trim_bad_head=(25, 12), trim_bad_tail=(25, 40),
                        hard_trim=(2, 0, 0, 3), polyG_tail=8, threads=3, patch_size=40)
    o1, o2, ost, oerr = oracle_run(p, d)
    c1, c2, cst, cerr = core_replay(p, d, tile_r=16, grid=7)
    assert oerr == cerr == 0
    assert (o1["category"] == 0).sum() > 1000 and (o1["head_cut"] > 2).sum() > 50
    assert_same((c1, c2, cst), (o1, o2, ost), "mixed tiles")


def test_every_byte_value_is_classified_like_the_oracle():
    """Unrecognized bases (anything but ACGTN, any case) must raise the error flag, exactly the
    characters the reference's switch rejects (read_filter.cpp:270-285)."""
    L = 48
    for b in range(1, 256):
        if b in (10, 13):
            continue
        S = np.zeros((1, 48), dtype=np.uint8); Q = np.zeros_like(S)
        S[0, :L] = np.frombuffer(b"ACGT" * 12, dtype=np.uint8)
        S[0, 17] = b
        Q[0, :L] = ord("I")
        d = dict(seq1=S, qual1=Q, len1=np.array([L], dtype=np.uint16))
        p = abi.make_params(is_pe=False)
        _, _, _, oerr = oracle_run(p, d)
        _, _, _, cerr = core_replay(p, d)
        assert (oerr & 1) == (cerr & 1), f"byte {b:#x}: oracle {oerr} replay {cerr}"
        assert bool(oerr & 1) == (chr(b) not in "ACGTNacgtn")


def test_every_quality_byte_value_flags_like_the_oracle():
    """Quality bytes below the Phred base or above the table's bins raise the bad-quality flag on both
    sides; bytes inside [phred, phred+64) never do, whether or not they fit the shared-memory bins."""
    L = 40
    for phred in (33, 64):
        for b in list(range(1, 10)) + list(range(10, 256, 3)) + [phred - 1, phred, phred + 41, phred + 42, phred + 63, phred + 64, 127, 128, 255]:
            if b in (10, 13) or b > 255:
                continue
            S = np.zeros((3, 48), dtype=np.uint8); Q = np.zeros_like(S)
            S[:, :L] = np.frombuffer(b"ACGT" * 10, dtype=np.uint8)
            Q[:, :L] = phred + 30
            Q[1, 23] = b
            d = dict(seq1=S, qual1=Q, len1=np.array([L, L, L - 7], dtype=np.uint16))
            p = abi.make_params(is_pe=False, quality_phred=phred)
            o1, _, ost, oerr = oracle_run(p, d)
            c1, _, cst, cerr = core_replay(p, d)
            assert oerr == cerr, f"phred {phred} byte {b}: oracle {oerr} replay {cerr}"
            assert bool(oerr & 2) == (not (phred <= b < phred + 64)), (phred, b, oerr)
            if oerr == 0:
                assert_same((c1, None, cst), (o1, None, ost), f"phred {phred} byte {b}")


def test_batches_and_first_index_compose():
    """Feeding the input as several batches with running first_index gives the same tables as one call."""
    d = synth.gen_pairs(9000, L=100, seed=21)
    p = abi.make_params(is_pe=True, threads=4, patch_size=20, **CFG2_KW)
    o1, o2, ost, _ = oracle_run(p, d)
    st = np.zeros_like(ost)
    import ctypes as C
    from helpers import load_coretest
    lib = load_coretest()
    cuts = [0, 1000, 1003, 4200, 9000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        sub = {k: (v[a:b] if isinstance(v, np.ndarray) else v) for k, v in d.items()}
        sub = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in sub.items()}
        b1 = abi.make_batch(sub["seq1"], sub["qual1"], sub["len1"]); b2 = abi.make_batch(sub["seq2"], sub["qual2"], sub["len2"])
        r1 = np.zeros(b - a, dtype=abi.RESULT_DTYPE); r2 = np.zeros(b - a, dtype=abi.RESULT_DTYPE)
        err = C.c_uint32(0)
        lib.coretest_filter(C.byref(p), C.byref(b1), C.byref(b2), r1.ctypes.data, r2.ctypes.data, st.ctypes.data, a, C.byref(err), 0, 4, -1)
        assert np.array_equal(r1, o1[a:b]) and np.array_equal(r2, o2[a:b])
    assert np.array_equal(st, ost)


def adversarial_adapter_reads(n, L, adapter, seed):
    """Reads salted with mutated / truncated / shifted copies of the adapter, N's and lowercase:
    every branch of adapter_pos (phase 1 shifts, phase 2 budgets and run-accepts, phase 3 overhangs)."""
    rng = np.random.default_rng(seed)
    A = len(adapter)
    ada = np.frombuffer(adapter, dtype=np.uint8)
    stride = synth.stride_for(L)
    S = np.zeros((n, stride), dtype=np.uint8)
    Q = np.zeros((n, stride), dtype=np.uint8)
    Ln = rng.integers(max(A - 1, 25), L + 1, size=n).astype(np.uint16)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(n):
        l = int(Ln[i])
        s = acgt[rng.integers(0, 4, size=l)].copy()
        mode = rng.integers(0, 6)
        if mode < 5:
            piece = ada.copy()
            for _ in range(int(rng.integers(0, 5))):          # 0..4 substitutions
                piece[rng.integers(0, A)] = acgt[rng.integers(0, 4)]
            if mode == 0:       # anywhere inside
                off = int(rng.integers(0, max(1, l - A + 1)))
            elif mode == 1:     # hanging over the 3' end
                off = int(rng.integers(max(0, l - A + 1), l))
            elif mode == 2:     # starting before the read
                off = -int(rng.integers(1, 8))
            elif mode == 3:     # only a prefix of the adapter, inside
                piece = piece[: int(rng.integers(6, A))]
                off = int(rng.integers(0, max(1, l - len(piece) + 1)))
            else:               # exactly at the boundary lengths
                off = l - int(rng.integers(4, 9))
            a0, b0 = max(off, 0), min(l, off + len(piece))
            if b0 > a0:
                s[a0:b0] = piece[a0 - off:b0 - off]
        if rng.random() < 0.2:
            s[rng.integers(0, l, size=2)] = ord("N")
        if rng.random() < 0.1:
            j = int(rng.integers(0, l))
            s[j] = s[j] | 0x20
        S[i, :l] = s
        Q[i, :l] = rng.integers(35, 74, size=l)
    return dict(seq1=S, qual1=Q, len1=Ln, n=n, L=L, stride=stride)


@pytest.mark.parametrize("adapter,L,kw", [
    (synth.ADAPTER1, 100, dict()),
    (synth.ADAPTER2, 150, dict()),
    (synth.SRNA_ADAPTER3, 50, dict()),
    (synth.ADAPTER1, 150, dict(ada_mis=(3, 3), ada_mr=(0.4, 0.4), ada_edge=(4, 4))),
    (synth.ADAPTER2, 100, dict(ada_mis=(0, 0), ada_mr=(0.9, 0.9), ada_edge=(10, 10))),
    (synth.ADAPTER1, 70, dict(ada_mis=(1, 1), ada_mr=(0.1, 0.1), ada_edge=(6, 6))),
    (b"ACGTTGCAAC", 60, dict(ada_mis=(1, 1), ada_mr=(0.5, 0.5), ada_edge=(3, 3))),
    (synth.ADAPTER1 + synth.ADAPTER2[:31], 150, dict()),                      # 63 bases: both plane words in use
    (synth.ADAPTER1 + synth.ADAPTER2, 150, dict()),                           # 74 bases: byte-wise path
    (b"AAGTCGGAGGCCAAGCGGTCTTAGGNAGACAA", 100, dict()),                       # adapter with N: byte-wise path
], ids=["a32", "a42", "a21", "a32_loose", "a42_strict", "a32_shortrun", "a10", "a63", "a74_bytes", "a32_withN_bytes"])
def test_adapter_matcher_adversarial(adapter, L, kw):
    d = adversarial_adapter_reads(6000, L, adapter[:42] if len(adapter) > 64 else adapter, seed=len(adapter) * 7 + L)
    for trim in (True, False):
        p = abi.make_params(is_pe=False, adapter1=adapter.decode(), ada_trim=trim, min_read_length=10, **kw)
        o1, _, ost, oerr = oracle_run(p, d)
        c1, _, cst, cerr = core_replay(p, d)
        assert (o1["adacut_pos"] >= 0).sum() > 500, "generator should produce many adapter hits"
        assert_same((c1, None, cst), (o1, None, ost), f"adapter len {len(adapter)} trim={trim}")


def random_case(seed, scale=1):
    """Random option set + random batch (the offline twin of this generator, run against the reference binary
    over several hundred seeds, found no discrepancy outside the reference's undefined behaviour)."""
    import random
    rnd = random.Random(seed)
    pe = rnd.random() < 0.6
    L = rnd.choice([17, 36, 50, 75, 100, 150, 151, 250])
    n = rnd.choice([300, 700, 1100]) * scale
    kw = dict(threads=rnd.choice([1, 2, 3, 5]), patch_size=rnd.choice([7, 20, 33]) * scale)
    if rnd.random() < 0.8:
        kw["adapter1"] = rnd.choice([A1, [A2.lower(), A1], A1[:10], A1 + A2[:31]])
        if pe:
            kw["adapter2"] = A2
        kw["ada_trim"] = rnd.random() < 0.6
    kw["low_qual"] = rnd.choice([2, 5, 10, 20])
    kw["low_qual_ratio"] = rnd.choice([0.1, 0.3, 0.5, 0.9, -1.0])
    kw["mean_quality"] = rnd.choice([-1, 10, 20, 30])
    kw["n_ratio"] = rnd.choice([0.01, 0.05, 0.2, -1.0])
    kw["highA_ratio"] = rnd.choice([-1.0, 0.3, 0.5, 0.8])
    kw["polyG_tail"] = rnd.choice([-1.0, 3, 8, 15])
    kw["polyX_num"] = rnd.choice([-1, 1, 6, 12, 40])
    kw["min_read_length"] = rnd.choice([-1, 0, 10, 30, 60])
    kw["max_read_length"] = rnd.choice([-1, -1, 40, 100, 140])
    if rnd.random() < 0.5:
        kw["trim_bad_head"] = (rnd.choice([10, 20, 30]), rnd.choice([0, 5, 10, 400]))
    if rnd.random() < 0.5:
        kw["trim_bad_tail"] = (rnd.choice([10, 20, 30]), rnd.choice([0, 5, 20, 400]))
    if rnd.random() < 0.4:
        kw["hard_trim"] = tuple(rnd.choice([0, 2, 7, 60]) for _ in range(4))
    if rnd.random() < 0.3:
        kw["ada_mis"] = (rnd.choice([0, 1, 3, 7]), rnd.choice([0, 2, 4]))
    if rnd.random() < 0.3:
        kw["ada_edge"] = (rnd.choice([0, 3, 6, 10]), rnd.choice([4, 6, 12, 40]))
    if rnd.random() < 0.3:
        kw["ada_mr"] = (rnd.choice([0.1, 0.3, 0.5, 0.8, 1.0]), rnd.choice([0.4, 0.5, 0.9]))
    if rnd.random() < 0.25:
        kw["contam1"] = rnd.choice([_C1, f"{_C1},{_C3}"])
        kw["ct_match_r"] = "0.3" if "," not in kw["contam1"] else "0.3,0.6"
        if pe:
            kw["contam2"] = _C2 if "," not in kw["contam1"] else f"{_C2},{_C1}"
        kw["contam_trim"] = rnd.random() < 0.2
    if rnd.random() < 0.2:
        kw["global_contams"] = _C3; kw["glob_cotm_mR"] = rnd.choice(["0.4", "0.7"]); kw["glob_cotm_mM"] = rnd.choice(["0", "2"])
    kw["quality_phred"] = 33
    kw["max_base_quality"] = rnd.choice([42, 42, 30, 63])
    if not pe and rnd.random() < 0.25:        # the filtersRNA module
        for k in ("contam1", "contam2", "ct_match_r", "contam_trim", "global_contams", "glob_cotm_mR", "glob_cotm_mM", "mean_quality", "n_ratio"):
            kw.pop(k, None)
        kw.update(srna=True, adapter1=rnd.choice([synth.SRNA_ADAPTER5, synth.SRNA_ADAPTER5[:12], synth.SRNA_ADAPTER5 + synth.SRNA_ADAPTER5[:20]]),
                  adapter2=rnd.choice([synth.SRNA_ADAPTER3, synth.SRNA_ADAPTER3 + synth.SRNA_ADAPTER3, synth.SRNA_ADAPTER3[:9]]),
                  ada_rctg=rnd.choice([4, 6, 7]), ada_rar=rnd.choice([0.5, 0.8, 0.95]), ada_rma=rnd.choice([3, 5, 8]),
                  ada_rer=rnd.choice([0.2, 0.4, 0.7]), ada_rmm=rnd.choice([0, 2, 4, 6]))
        kw.pop("ada_mis", None); kw.pop("ada_edge", None); kw.pop("ada_mr", None)
        d = synth.gen_srna(n, L=max(L, 36), seed=seed, var_len=rnd.random() < 0.5)
    else:
        d = synth.gen_pairs(n, L=L, seed=seed, se=not pe, var_len=rnd.random() < 0.5, polyg_frac=rnd.choice([0.04, 0.3]))
    if "contam1" in kw or "global_contams" in kw:
        synth.add_contams(d, CONTAM_PLANTS, seed=seed)
    if rnd.random() < 0.3:                    # tile / fov flags as a caller of the SoA entry points would set them
        fl = np.random.default_rng(seed).choice(np.array([0, 0, 0, abi.PRE_TILE, abi.PRE_FOV, abi.PRE_TILE | abi.PRE_FOV], dtype=np.uint16), size=n)
        d["len1"] = d["len1"] | fl
    return pe, kw, d, dict(first=rnd.choice([0, 12345, 10**9]), tile_r=rnd.choice([0, 8, 24]), grid=rnd.choice([1, 3, 7]))


@pytest.mark.parametrize("seed", range(40))
def test_random_options_replay_matches_oracle(seed):
    pe, kw, d, rkw = random_case(seed)
    p = abi.make_params(is_pe=pe, **kw)
    o1, o2, ost, oerr = oracle_run(p, d, first=rkw["first"])
    c1, c2, cst, cerr = core_replay(p, d, **rkw)
    assert oerr == cerr
    if oerr == 0:
        assert_same((c1, c2, cst), (o1, o2, ost), f"seed {seed}: {kw}")
