"""GPU tier: the CUDA engine, called through the C ABI exactly as a reference-side binding would,
must reproduce the oracle bit for bit: per-read result records, every statistics word, the clean
FASTQ and the report files. Bit-exact is the bar for all of it (integer/byte work; the five fp32
ratio predicates only feed discrete decisions)."""
import ctypes as C
import filecmp
import glob
import gzip
import os

import numpy as np
import pytest

from helpers import A1, A2, CFG2_KW, Engine, abi, assert_same, oracle_run, synth

pytestmark = pytest.mark.gpu

CONFIGS = [
    ("cfg1_se150_default", False, 60000, 150, dict(seed=1001), dict()),
    ("cfg2_pe150_all", True, 60000, 150, dict(seed=1002), dict(CFG2_KW, threads=3, patch_size=1000)),
    ("cfg2_pe150_discard", True, 30000, 150, dict(seed=1003), dict(adapter1=A1, adapter2=A2)),
    ("cfg4_se50_adapter", False, 60000, 50, dict(seed=1004, adapter1=synth.SRNA_ADAPTER3, insert_range=(15, 35)),
     dict(adapter1=synth.SRNA_ADAPTER3.decode(), ada_trim=True, min_read_length=15)),
    ("cfg5_pe250_polyg", True, 30000, 250, dict(seed=1005, polyg_frac=0.3), dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10)),
    ("pe120_varlen_hardtrim", True, 20000, 120, dict(seed=7, var_len=True),
     dict(adapter1=A1, adapter2=A2, ada_trim=True, hard_trim=(3, 2, 5, 1), polyX_num=8, threads=2, patch_size=50)),
    ("se150_two_adapters_lowercase", False, 20000, 150, dict(seed=10), dict(adapter1=[A2.lower(), A1], ada_trim=True)),
    ("pe150_minlen_off", True, 20000, 150, dict(seed=11), dict(CFG2_KW, min_read_length=-1, max_read_length=140)),
    ("pe400_long_reads", True, 4000, 400, dict(seed=12), dict(adapter1=A1, adapter2=A2, ada_trim=True)),
    ("pe1000_max_len", True, 600, 1000, dict(seed=13), dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10)),
    ("se_tiny_ragged", False, 37, 75, dict(seed=14, var_len=True), dict(adapter1=A1)),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_engine_matches_oracle(cfg, engine_lib):
    name, pe, n, L, gkw, pkw = cfg
    d = synth.gen_pairs(n, L=L, se=not pe, **gkw)
    p = abi.make_params(is_pe=pe, **pkw)
    o1, o2, ost, oerr = oracle_run(p, d)
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d)
        st = e.stats()
        flags, _ = e.error_flags()
    assert flags == oerr == 0
    assert_same((r1, r2, st), (o1, o2, ost), name)


@pytest.mark.parametrize("cfg", [c for c in CONFIGS if c[3] <= 250], ids=[c[0] for c in CONFIGS if c[3] <= 250])
def test_warp_specialised_kernel_matches_oracle(cfg, engine_lib, monkeypatch):
    """The experimental warp-specialised kernel (SNK_KERNEL=ws: TMA producer / scan warps / histogram warps handing
    tiles over by mbarriers, rows up to 256 bytes) must be bit-exact too; the production kernel is filter_kernel."""
    monkeypatch.setenv("SNK_KERNEL", "ws")
    name, pe, n, L, gkw, pkw = cfg
    d = synth.gen_pairs(n, L=L, se=not pe, **gkw)
    p = abi.make_params(is_pe=pe, **pkw)
    o1, o2, ost, oerr = oracle_run(p, d)
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d)
        st = e.stats()
        flags, _ = e.error_flags()
    assert flags == oerr == 0
    assert_same((r1, r2, st), (o1, o2, ost), name + " (ws kernel)")


@pytest.mark.parametrize("cfg", [
    ("srna_discard_L44", 60000, 44, dict(), dict()),
    ("srna_trim_polyg_varlen", 60000, 50, dict(var_len=True), dict(ada_trim=True, polyG_tail=6, highA_ratio=0.6, polyX_num=12)),
    ("srna_trim_hard_params", 30000, 75, dict(), dict(ada_trim=True, hard_trim=(2, 1), min_read_length=15, max_read_length=60,
                                                      ada_rctg=7, ada_rar=0.7, ada_rma=6, ada_rer=0.3, ada_rmm=3)),
], ids=lambda c: c[0])
def test_engine_matches_oracle_filtersRNA(cfg, engine_lib):
    """filtersRNA module on the engine (sRNA_findAdapter, sRNA_hasAdapter, cut at the 3' adapter, sRNA_discard)."""
    name, n, L, gkw, pkw = cfg
    d = synth.gen_srna(n, L=L, seed=len(name) * 31 + L, **gkw)
    kw = dict(min_read_length=18, max_read_length=49)
    kw.update(pkw)
    p = abi.make_params(is_pe=False, srna=True, adapter1=synth.SRNA_ADAPTER5, adapter2=synth.SRNA_ADAPTER3, threads=2, patch_size=1000, **kw)
    o1, _, ost, oerr = oracle_run(p, d)
    with Engine(engine_lib, p) as e:
        r1, _ = e.filter_host(d)
        st = e.stats()
        flags, _ = e.error_flags()
    assert flags == oerr == 0
    assert_same((r1, None, st), (o1, None, ost), name)


@pytest.mark.parametrize("pe", [True, False], ids=["pe", "se"])
def test_tile_fov_flags_in_len(pe, engine_lib):
    """SNK_PRE_TILE / SNK_PRE_FOV bits of len[] through the SoA entry points (see the CPU tier's test)."""
    import oracle_py as orc
    n = 40000
    d = synth.gen_pairs(n, L=100, seed=61, se=not pe, var_len=True)
    p = abi.make_params(is_pe=pe, adapter1=A1, adapter2=A2 if pe else None, ada_trim=True, tile="1102,2201", fov="C002R003", threads=2, patch_size=500)
    ids = [a if i % 3 else b for i, (a, b) in enumerate(zip(synth.tile_ids(n, 1), synth.fov_ids(n, 1)))]
    d["len1"] = d["len1"] | orc.id_flags(p, ids)
    if pe:
        d["len2"] = d["len2"] | np.where(np.arange(n) % 11 == 0, abi.PRE_TILE, 0).astype(np.uint16)
    o1, o2, ost, oerr = oracle_run(p, d)
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d)
        st = e.stats()
        flags, _ = e.error_flags()
    assert flags == oerr == 0
    assert_same((r1, r2, st), (o1, o2, ost), "tile/fov")


def test_engine_matches_oracle_contam(engine_lib):
    """Contaminant and global contaminant sequences on the engine (the CPU tier's CONTAM_CONFIGS, larger batches)."""
    from test_core_replay import CONTAM_CONFIGS, CONTAM_PLANTS
    for name, pe, n, L, gkw, pkw in CONTAM_CONFIGS:
        d = synth.add_contams(synth.gen_pairs(5 * n, L=L, seed=len(name) * 13, se=not pe, **gkw), CONTAM_PLANTS, seed=L)
        p = abi.make_params(is_pe=pe, threads=2, patch_size=600, **pkw)
        o1, o2, ost, oerr = oracle_run(p, d)
        with Engine(engine_lib, p) as e:
            r1, r2 = e.filter_host(d)
            st = e.stats()
            flags, _ = e.error_flags()
        assert flags == oerr == 0
        assert_same((r1, r2, st), (o1, o2, ost), name)


@pytest.mark.parametrize("seed,scale", [(s, 1) for s in range(24)] + [(s, 40) for s in range(100, 112)])
def test_random_options_engine_matches_oracle(seed, scale, engine_lib):
    """Random option sets (the CPU tier's generator) through the engine: small ragged batches and larger ones."""
    from test_core_replay import random_case
    pe, kw, d, rkw = random_case(seed, scale)
    p = abi.make_params(is_pe=pe, **kw)
    o1, o2, ost, oerr = oracle_run(p, d, first=rkw["first"])
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d, first=rkw["first"])
        st = e.stats()
        flags, _ = e.error_flags()
    assert flags == oerr
    if oerr == 0:
        assert_same((r1, r2, st), (o1, o2, ost), f"seed {seed} x{scale}: {kw}")


def test_mixed_checked_and_unchecked_tiles(engine_lib):
    """Records with qualities above the shared-memory bins scattered through the batch (see the CPU
    tier's test of the same name): checked and unchecked tiles, raw and delta cells must add up."""
    from helpers import high_quality_mix
    d = high_quality_mix(n=40000)
    p = abi.make_params(is_pe=True, adapter1=A1, adapter2=A2, ada_trim=True, trim_bad_head=(25, 12), trim_bad_tail=(25, 40),
                        hard_trim=(2, 0, 0, 3), polyG_tail=8, threads=3, patch_size=40)
    o1, o2, ost, oerr = oracle_run(p, d)
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d)
        st = e.stats()
        flags, _ = e.error_flags()
    assert flags == oerr == 0
    assert_same((r1, r2, st), (o1, o2, ost), "mixed tiles")


def test_batches_compose_and_empty_batch(engine_lib):
    d = synth.gen_pairs(30000, L=100, seed=21)
    p = abi.make_params(is_pe=True, threads=4, patch_size=20, **CFG2_KW)
    o1, o2, ost, _ = oracle_run(p, d)
    cuts = [0, 0, 1000, 1003, 14200, 30000]
    with Engine(engine_lib, p) as e:
        for a, b in zip(cuts[:-1], cuts[1:]):
            sub = {k: (np.ascontiguousarray(v[a:b]) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
            r1, r2 = e.filter_host(sub, first=a)
            assert np.array_equal(r1, o1[a:b]) and np.array_equal(r2, o2[a:b])
        st = e.stats()
        assert_same((o1, o2, st), (o1, o2, ost), "composed batches")
        # reset clears every table
        assert e.lib.snk_engine_stats_reset(e.h) == 0
        assert not e.stats().any()


def test_device_pointer_entry_point_with_torch_memory(engine_lib):
    """snk_filter_pe_device on torch-owned device tensors and torch's current stream."""
    import torch
    d = synth.gen_pairs(50000, L=150, seed=33)
    p = abi.make_params(is_pe=True, **CFG2_KW)
    o1, o2, ost, _ = oracle_run(p, d)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(v.view(np.int16) if v.dtype == np.uint16 else v).to(dev) for k, v in d.items() if isinstance(v, np.ndarray)}
    out1 = torch.zeros(50000, dtype=torch.int64, device=dev)
    out2 = torch.zeros(50000, dtype=torch.int64, device=dev)
    b1 = abi.Batch(t["seq1"].data_ptr(), t["qual1"].data_ptr(), t["len1"].data_ptr(), 50000, d["stride"])
    b2 = abi.Batch(t["seq2"].data_ptr(), t["qual2"].data_ptr(), t["len2"].data_ptr(), 50000, d["stride"])
    with Engine(engine_lib, p) as e:
        stream = torch.cuda.current_stream().cuda_stream
        e.check(e.lib.snk_filter_pe_device(e.h, C.byref(b1), C.byref(b2), out1.data_ptr(), out2.data_ptr(), 0, C.c_void_p(stream)))
        torch.cuda.synchronize()
        r1 = out1.cpu().numpy().view(abi.RESULT_DTYPE)
        r2 = out2.cpu().numpy().view(abi.RESULT_DTYPE)
        st = e.stats()
        assert e.lib.snk_engine_launch_count(e.h) == 1
    assert_same((r1, r2, st), (o1, o2, ost), "device entry point")


def test_unrecognized_base_and_bad_quality_raise_flags(engine_lib):
    d = synth.gen_pairs(5000, L=150, seed=44, se=True)
    d["seq1"][1234, 17] = ord("X")
    p = abi.make_params(is_pe=False)
    with Engine(engine_lib, p) as e:
        e.filter_host(d)
        flags, idx = e.error_flags()
    assert flags & 1 and idx == 1234          # reference: "Error:unrecognized sequence" + exit(1)
    d = synth.gen_pairs(5000, L=150, seed=45, se=True)
    d["qual1"][77, 3] = 20                    # below the Phred-33 base
    with Engine(engine_lib, p) as e:
        e.filter_host(d)
        flags, idx = e.error_flags()
    assert flags & 2 and idx <= 77


def test_golden_reports_and_clean_fastq_through_the_engine(engine_lib, tmp_path):
    """The committed outputs of the reference binary, reproduced by engine + report writer."""
    from test_oracle import CASES, clean_bytes, load_case
    for name in CASES:
        gd, meta, data, p = load_case(name)
        with Engine(engine_lib, p) as e:
            r1, r2 = e.filter_host(data)
            st = e.stats()
        order = abi.ref_output_order(meta["n"], meta["threads"], 1 << 20, meta["patch_size"], gz_input=False, pe=meta["pe"])
        for m, r in ((1, r1), (2, r2)):
            if r is None:
                continue
            assert clean_bytes(data, r, m, order) == gzip.open(os.path.join(gd, f"c{m}.fq.gz")).read(), f"{name}: clean fq{m}"
        out = tmp_path / name
        out.mkdir()
        fn = engine_lib.snk_report_write_pe if meta["pe"] else engine_lib.snk_report_write_se
        assert fn(C.byref(p), st.ctypes.data, str(out).encode()) == 0
        for f in glob.glob(os.path.join(gd, "*.txt")):
            assert filecmp.cmp(f, str(out / os.path.basename(f)), shallow=False), f"{name}: {os.path.basename(f)}"


def test_full_size_properties(engine_lib):
    """BASELINE config-2 shape at a size the oracle cannot finish in seconds (4 M pairs): check
    size-independent properties, and exact parity on a sampled sub-range."""
    n = 1 << 22
    base = synth.gen_pairs(1 << 18, L=150, seed=1002)
    reps = n // (1 << 18)
    d = {k: (np.ascontiguousarray(np.tile(v, (reps, 1)) if v.ndim == 2 else np.tile(v, reps)) if isinstance(v, np.ndarray) else v)
         for k, v in base.items()}
    p = abi.make_params(is_pe=True, threads=8, nprocs=64, **CFG2_KW)
    with Engine(engine_lib, p) as e:
        r1, r2 = e.filter_host(d)
        st = e.stats().reshape(p.n_slots, abi.SLOT_WORDS)
        flags, _ = e.error_flags()
    assert flags == 0
    tot = st.sum(axis=0)
    kept = int((r1["category"] == 0).sum())

    def gs(f, i):
        return int(tot[abi.slot_file_off(f) + i])
    assert gs(abi.RAW1, abi.GS_READS) == n and gs(abi.RAW2, abi.GS_READS) == n
    assert gs(abi.RAW1, abi.GS_BASES) == int(d["len1"].astype(np.int64).sum())
    assert gs(abi.CLEAN1, abi.GS_READS) == kept == gs(abi.CLEAN2, abi.GS_READS)
    assert gs(abi.CLEAN1, abi.GS_BASES) == int(r1["clean_len"][r1["category"] == 0].astype(np.int64).sum())
    # every discarded pair is counted exactly once
    drops = sum(int(tot[b]) for b in abi.FS_BASE.values())
    assert drops == n - kept
    for f in (abi.RAW1, abi.RAW2, abi.CLEAN1, abi.CLEAN2):
        off = abi.slot_file_off(f)
        bs = tot[off + abi.FILE_BS_OFF: off + abi.FILE_QS_OFF]
        qs = tot[off + abi.FILE_QS_OFF: off + abi.FILE_TS_OFF]
        assert int(bs.sum()) == gs(f, abi.GS_BASES) == int(qs.sum())
        assert int(bs.reshape(-1, 5)[:, 0].sum()) == gs(f, abi.GS_A)
        assert int(qs.reshape(-1, abi.QBINS)[:, 20:].sum()) == gs(f, abi.GS_Q20)
    # periodic input => periodic results; and the first period equals the oracle
    assert np.array_equal(r1[: 1 << 18], r1[(1 << 18):(1 << 19)])
    o1, o2, _, _ = oracle_run(p, base)
    assert np.array_equal(r1[: 1 << 18], o1) and np.array_equal(r2[: 1 << 18], o2)


def test_adapter_matcher_adversarial_on_gpu(engine_lib):
    from test_core_replay import adversarial_adapter_reads
    for adapter, L, kw in ((synth.ADAPTER1, 100, dict()), (synth.ADAPTER2, 150, dict()),
                           (synth.ADAPTER1, 150, dict(ada_mis=(3, 3), ada_mr=(0.4, 0.4), ada_edge=(4, 4))),
                           (synth.ADAPTER1 + synth.ADAPTER2[:31], 150, dict()),
                           (b"AAGTCGGAGGCCAAGCGGTCTTAGGNAGACAA", 100, dict())):
        d = adversarial_adapter_reads(20000, L, adapter, seed=len(adapter) + L)
        p = abi.make_params(is_pe=False, adapter1=adapter.decode(), ada_trim=True, min_read_length=10, **kw)
        o1, _, ost, _ = oracle_run(p, d)
        with Engine(engine_lib, p) as e:
            r1, _ = e.filter_host(d)
            st = e.stats()
        assert_same((r1, None, st), (o1, None, ost), f"adapter len {len(adapter)}")


def test_len_beyond_the_row_raises_the_length_flag(engine_lib):
    """A len[] entry larger than the batch stride (or than SNK_MAX_READ_LEN) is an error of the caller: the row is not
    processed and error bit 3 names the read, instead of silently clamping and counting into a neighbouring table."""
    d = synth.gen_pairs(4000, L=150, seed=46, se=True)
    d["len1"] = d["len1"].copy()
    d["len1"][321] = 161                      # stride is 160
    p = abi.make_params(is_pe=False)
    with Engine(engine_lib, p) as e:
        r1, _ = e.filter_host(d)
        flags, idx = e.error_flags()
        st = e.stats()
    assert flags & 8 and idx == 321
    ok = synth.gen_pairs(4000, L=150, seed=46, se=True)
    keep = np.ones(4000, dtype=bool); keep[321] = False
    o1, _, _, _ = oracle_run(p, ok)
    assert np.array_equal(r1[keep], o1[keep]), "the other reads of the batch are unaffected"
