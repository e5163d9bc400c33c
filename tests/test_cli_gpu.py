"""GPU tier, end to end: the drop-in CLI (soapnuke_b200/bin/SOAPnuke filter ...) against the
unmodified reference binary (oracle/_ref/SOAPnuke) on the same FASTQ files and flags: decompressed
clean FASTQ and every report file must be byte-identical. Falls back to the committed golden
outputs when the reference binary is not available on the box."""
import filecmp
import glob
import gzip
import json
import os
import shutil
import atexit
import concurrent.futures
import subprocess
import tempfile
import threading
import zlib

import pytest

import oracle_py as orc
from helpers import A1, A2, CFG2_FLAGS, ROOT, fastq_text, report_equal, synth

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke")


@pytest.fixture(scope="module")
def cli():
    from soapnuke_b200 import build
    build.build_all()
    assert os.path.exists(CLI)
    return CLI


def read_maybe_gz(path):
    return gzip.open(path).read() if path.endswith(".gz") else open(path, "rb").read()


# Every reference run sleeps in 5 s polling quanta (peprocess.cpp:3039), so the reference sides of ALL cases are
# started together in the background the first time one of them is needed; a test then only waits for its own.
_ALL_CASES = []
_FUTURES = {}
_LOCK = threading.Lock()
_ROOT = None


def reg(cases, **common):
    for c in cases:
        c.update(common)
    _ALL_CASES.extend(cases)
    return cases


def _prepare(root, name, pe, n, L, T, flags, gkw=None, patch=None, gz_in=False, gz_out=False, env=None, cfg=None, index_ids=False,
             module="filter", idfn=None, contams=None, trim=False):
    """Writes the case's input files and runs the reference binary on them."""
    w = os.path.join(root, name)
    os.makedirs(w)
    if module == "filtersRNA":
        d = synth.gen_srna(n, L=L, seed=zlib.crc32(name.encode()) % 100000, **(gkw or {}))
    else:
        d = synth.gen_pairs(n, L=L, seed=zlib.crc32(name.encode()) % 100000, se=not pe, **(gkw or {}))
    if contams:
        synth.add_contams(d, contams, seed=n)
    ext_in = ".fq.gz" if gz_in else ".fq"
    ext_out = ".fq.gz" if gz_out else ".fq"
    def write(path, m):
        if not index_ids:
            return synth.write_fastq(path, d[f"seq{m}"], d[f"qual{m}"], d[f"len{m}"], m, gz=gz_in, ids=idfn(n, m) if idfn else None)
        ids = [b"@FCD1PB1ACXX:4:1101:%d:%d#GAAGCACG/%d" % (i // 1000, i % 1000, m) for i in range(n)]
        data = fastq_text(ids, d[f"seq{m}"], d[f"qual{m}"], d[f"len{m}"])
        (gzip.open(path, "wb", compresslevel=2) if gz_in else open(path, "wb")).write(data)
    write(f"{w}/r1{ext_in}", 1)
    base = ["-1", f"{w}/r1{ext_in}", "-C", "c1" + ext_out, "-T", str(T)]
    if pe:
        write(f"{w}/r2{ext_in}", 2)
        base += ["-2", f"{w}/r2{ext_in}", "-D", "c2" + ext_out]
    if trim:                              # config keys trimFq1/2: every record after trimming (gzip only)
        cfg = list(cfg or []) + ["trimFq1=t1.fq.gz"] + (["trimFq2=t2.fq.gz"] if pe else [])
    if patch or cfg:
        open(f"{w}/cfg.txt", "w").write((f"patch={patch}\n" if patch else "") + "".join(l + "\n" for l in (cfg or [])))
        base += ["-c", f"{w}/cfg.txt"]
    r = orc.run_reference(base + ["-o", f"{w}/ref"] + flags, module=module)
    return dict(w=w, base=base, ext_out=ext_out, ref=r)


def _reference_side(name):
    global _ROOT
    with _LOCK:
        if _ROOT is None:
            _ROOT = tempfile.mkdtemp(prefix="snk_cli_")
            atexit.register(shutil.rmtree, _ROOT, ignore_errors=True)
            pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) // 2)))
            for c in _ALL_CASES:
                _FUTURES[c["name"]] = pool.submit(_prepare, _ROOT, **c)
    return _FUTURES[name].result()


def run_both(cli, tmp, name, pe, flags, env=None, module="filter", trim=False, out_name="mine", **_unused):
    st = _reference_side(name)
    w, base, ext_out, r = st["w"], st["base"], st["ext_out"], st["ref"]
    assert r.returncode == 0, r.stderr.decode()
    e = dict(os.environ)
    e.update(env or {})
    m = subprocess.run([cli, module] + base + ["-o", f"{w}/{out_name}"] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e, timeout=600)
    assert m.returncode == 0, m.stderr.decode()
    for mate in (1, 2) if pe else (1,):
        a = read_maybe_gz(f"{w}/ref/c{mate}{ext_out}")
        b = read_maybe_gz(f"{w}/{out_name}/c{mate}{ext_out}")
        assert a == b, f"{name}: clean fq{mate} differs ({len(a)} vs {len(b)} bytes)"
    if trim:
        for mate in (1, 2) if pe else (1,):
            a = read_maybe_gz(f"{w}/ref/t{mate}.fq.gz")
            b = read_maybe_gz(f"{w}/{out_name}/t{mate}.fq.gz")
            assert len(a) > 0 and a == b, f"{name}: trim fq{mate} differs ({len(a)} vs {len(b)} bytes)"
    reports = sorted(glob.glob(f"{w}/ref/*.txt"))
    assert len(reports) == (10 if pe else 6)
    for f in reports:
        assert report_equal(f, f"{w}/{out_name}/{os.path.basename(f)}"), f"{name}: {os.path.basename(f)} differs"


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary not available")
@pytest.mark.parametrize("case", reg([
    dict(name="pe_cfg2_plain_T1", pe=True, n=30000, L=150, T=1, flags=CFG2_FLAGS),
    dict(name="pe_cfg2_plain_T4_multicycle", pe=True, n=30000, L=150, T=4, flags=CFG2_FLAGS, patch=25),
    dict(name="pe_cfg2_gz_T3", pe=True, n=20000, L=150, T=3, flags=CFG2_FLAGS, patch=40, gz_in=True, gz_out=True),
    dict(name="pe_discard_small_batches", pe=True, n=20000, L=100, T=2, flags=["-f", A1, "-r", A2], env={"SNK_BATCH_READS": "3000"}),
    dict(name="pe_polyg250", pe=True, n=8000, L=250, T=2, flags=["-f", A1, "-r", A2, "-J", "-g", "10"], gkw=dict(polyg_frac=0.3)),
    dict(name="pe_varlen_hardtrim", pe=True, n=12000, L=120, T=2, flags=["-f", A1, "-r", A2, "-J", "-t", "3,2,4,1"], gkw=dict(var_len=True), patch=30),
    dict(name="pe_index_peinfo_seqtype0", pe=True, n=12000, L=100, T=2, flags=["-f", A1, "-r", A2, "-J"], cfg=["index", "pe_info"], index_ids=True),
    dict(name="pe_index_seqtype1_fasta", pe=True, n=12000, L=100, T=3, flags=["-f", A1, "-r", A2, "-J"], patch=17,
         cfg=["index", "seqType=1", "outFileType=fasta"], index_ids=True),
    dict(name="se_fasta_gz", pe=False, n=12000, L=100, T=2, flags=["-f", A1], cfg=["outFileType=fasta", "pe_info"], gz_in=True, gz_out=True),
    dict(name="pe_varlen_growing_stride", pe=True, n=20000, L=150, T=2, flags=["-f", A1, "-r", A2, "-J"], gkw=dict(var_len=True),
         env={"SNK_BATCH_READS": "2048"}, cfg=["outQualSys=1"]),
    dict(name="pe_cfg2_gz_big_members", pe=True, n=120000, L=150, T=4, flags=CFG2_FLAGS, gz_in=True, gz_out=True),
    dict(name="se_default", pe=False, n=30000, L=150, T=1, flags=[]),
    dict(name="se_adapter_T4", pe=False, n=30000, L=100, T=4, flags=["-f", A1, "-J", "-g", "10"], patch=11, gz_in=True),
]), ids=lambda c: c["name"])
def test_cli_matches_reference_binary(cli, tmp_path, case):
    run_both(cli, tmp_path, **case)


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary not available")
@pytest.mark.parametrize("case", reg([
    dict(name="tile_pe", pe=True, n=30000, L=100, T=2, flags=["-f", A1, "-r", A2, "-J"], cfg=["tile=1102,2201,9999"], idfn=synth.tile_ids),
    dict(name="tile_se_gz", pe=False, n=20000, L=100, T=1, flags=["-f", A1], cfg=["tile=1103"], idfn=synth.tile_ids, gz_in=True, gz_out=True),
    dict(name="fov_pe_multicycle", pe=True, n=20000, L=100, T=3, flags=["-f", A1, "-r", A2, "-J"], patch=20,
         cfg=["fov=C002R003,C004R001,C001R005"], idfn=synth.fov_ids),
    dict(name="fov_and_tile_se", pe=False, n=20000, L=100, T=1, flags=[], cfg=["fov=C003R002", "tile=0123"], idfn=synth.fov_ids),
]), ids=lambda c: c["name"])
def test_cli_tile_fov_matches_reference_binary(cli, tmp_path, case):
    """Config keys tile= / fov=: the ids are parsed on the device by the FASTQ text path."""
    run_both(cli, tmp_path, **case)


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary not available")
@pytest.mark.parametrize("case", reg([
    dict(name="trim_pe_T1", pe=True, n=20000, L=100, T=1, flags=CFG2_FLAGS, trim=True),
    dict(name="trim_pe_peinfo_index_multicycle", pe=True, n=12000, L=100, T=3, flags=["-f", A1, "-r", A2, "-J", "-x", "20,10", "-y", "20,30"], patch=17,
         cfg=["index", "pe_info"], index_ids=True, trim=True),
    dict(name="trim_pe_gz_T4_emptied_reads", pe=True, n=20000, L=100, T=4, flags=["-f", A1, "-r", A2, "-J", "-t", "30,30,40,45", "-4", "10"], patch=25,
         gz_in=True, gz_out=True, trim=True, cfg=["outQualSys=1"]),
    dict(name="trim_se_fasta", pe=False, n=15000, L=120, T=2, flags=["-f", A1, "-J", "-g", "8"], gkw=dict(var_len=True),
         cfg=["outFileType=fasta"], trim=True),
    dict(name="trim_pe_small_batches", pe=True, n=30000, L=150, T=2, flags=CFG2_FLAGS, patch=40, trim=True, env={"SNK_BATCH_READS": "3000"}),
]), ids=lambda c: c["name"])
def test_cli_trim_files_match_reference_binary(cli, tmp_path, case):
    """Config keys trimFq1= / trimFq2=: the trim files hold every record as fastq_trim left it, before the discard
    decision; with pe_info the clean ids then carry the mate suffix twice (preOutput runs again on the same record)."""
    run_both(cli, tmp_path, **case)


CT = [synth.CONTAM1, synth.CONTAM2, synth.CONTAM3]
CT1, CT2, CT3 = (c.decode() for c in CT)


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary not available")
@pytest.mark.parametrize("case", reg([
    dict(name="contam_pe_single", pe=True, n=30000, L=100, T=2, flags=["-f", A1, "-r", A2, "-J"], cfg=[f"contam1={CT1}", f"contam2={CT2}"], contams=CT),
    dict(name="contam_se_list_gz", pe=False, n=20000, L=120, T=1, flags=[], cfg=[f"contam1={CT1},{CT3}", "ctMatchR=0.3,0.5"], contams=CT,
         gkw=dict(var_len=True), gz_in=True, gz_out=True),
    dict(name="contam_pe_list_budgets", pe=True, n=20000, L=150, T=3, flags=["-f", A1, "-r", A2], patch=20,
         cfg=[f"contam1={CT3},{CT2},{CT1}", f"contam2={CT1},{CT2},{CT3[:20]}", "ctMatchR=0.2,0.6,0.9", "adaMis=1,3", "adaEdge=4,8"], contams=CT),
    dict(name="contam_trim_mode", pe=True, n=20000, L=100, T=1, flags=["-f", A1, "-r", A2, "-J"],
         cfg=[f"contam1={CT1}", f"contam2={CT2}", "contam_trim", "ctMatchR=0.4"], contams=CT),
    dict(name="gcontam_pe", pe=True, n=20000, L=100, T=2, flags=["-f", A1, "-r", A2, "-J"], patch=20,
         cfg=[f"global_contams={CT1},{CT3}", "glob_cotm_mR=0.5,0.6", "glob_cotm_mM=1,2"], contams=CT + [synth.revcomp(c) for c in CT]),
    dict(name="gcontam_se_with_contam", pe=False, n=20000, L=120, T=1, flags=[], gkw=dict(var_len=True),
         cfg=[f"global_contams={CT2}", "glob_cotm_mR=0.4", "glob_cotm_mM=0", f"contam1={CT1}"], contams=CT + [synth.revcomp(c) for c in CT]),
]), ids=lambda c: c["name"])
def test_cli_contam_matches_reference_binary(cli, tmp_path, case):
    """Config keys contam1= / contam2= / ctMatchR= / contam_trim."""
    run_both(cli, tmp_path, **case)


SA5, SA3 = synth.SRNA_ADAPTER5.decode(), synth.SRNA_ADAPTER3.decode()


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary not available")
@pytest.mark.parametrize("case", reg([
    dict(name="srna_trim_T1", n=40000, L=50, T=1, flags=["-f", SA5, "-r", SA3, "-J"]),
    dict(name="srna_trim_polyg_T3_gz", n=30000, L=50, T=3, flags=["-f", SA5, "-r", SA3, "-J", "-g", "6", "-p", "0.6", "-X", "12"],
         gkw=dict(var_len=True), patch=15, gz_in=True, gz_out=True),
    dict(name="srna_discard_L44", n=20000, L=44, T=2, flags=["-f", SA5, "-r", SA3]),
    dict(name="srna_hard_cfg", n=20000, L=75, T=2, flags=["-f", SA5, "-r", SA3, "-J", "-t", "2,1", "-4", "15"],
         cfg=["maxReadLen=60", "adaRCtg=7", "adaRAr=0.7", "adaRMa=6", "adaREr=0.3", "adaRMm=3"], env={"SNK_BATCH_READS": "3000"}),
], pe=False, module="filtersRNA"), ids=lambda c: c["name"])
def test_cli_filtersRNA_matches_reference_binary(cli, tmp_path, case):
    """`SOAPnuke filtersRNA` (seProcess + sRNA_findAdapter / sRNA_hasAdapter / sRNA_discard)."""
    run_both(cli, tmp_path, **case)


def test_cli_reproduces_golden_outputs(cli, tmp_path):
    golden = os.path.join(ROOT, "tests", "golden")
    for name in sorted(os.listdir(golden)):
        gd = os.path.join(golden, name)
        if not os.path.isdir(gd):
            continue
        meta = json.load(open(os.path.join(gd, "case.json")))
        w = tmp_path / name
        w.mkdir()
        args = ["-1", str(w / "r1.fq"), "-C", "c1.fq", "-o", str(w / "out"), "-T", str(meta["threads"])]
        open(w / "r1.fq", "wb").write(gzip.open(os.path.join(gd, "r1.fq.gz")).read())
        if meta["pe"]:
            open(w / "r2.fq", "wb").write(gzip.open(os.path.join(gd, "r2.fq.gz")).read())
            args += ["-2", str(w / "r2.fq"), "-D", "c2.fq"]
        if meta["patch_size"]:
            open(w / "cfg.txt", "w").write(f"patch={meta['patch_size']}\n")
            args += ["-c", str(w / "cfg.txt")]
        m = subprocess.run([cli, meta.get("module", "filter")] + args + meta["flags"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert m.returncode == 0, m.stderr.decode()
        for mate in (1, 2) if meta["pe"] else (1,):
            assert open(w / "out" / f"c{mate}.fq", "rb").read() == gzip.open(os.path.join(gd, f"c{mate}.fq.gz")).read(), f"{name}: clean fq{mate}"
        for f in glob.glob(os.path.join(gd, "*.txt")):
            assert filecmp.cmp(f, str(w / "out" / os.path.basename(f)), shallow=False), f"{name}: {os.path.basename(f)}"


def test_cli_error_convention(cli, tmp_path):
    """Errors print `Error:...` and exit(1), like the reference (e.g. unrecognized base)."""
    d = synth.gen_pairs(2000, L=100, seed=5, se=True)
    d["seq1"][100, 5] = ord("X")
    synth.write_fastq(str(tmp_path / "bad.fq"), d["seq1"], d["qual1"], d["len1"], 1)
    m = subprocess.run([cli, "filter", "-1", str(tmp_path / "bad.fq"), "-C", "c.fq", "-o", str(tmp_path / "o")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert m.returncode == 1 and b"Error:unrecognized sequence" in m.stderr
    m = subprocess.run([cli, "filter", "-C", "c.fq", "-o", str(tmp_path / "o")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert m.returncode == 1 and b"Error:input fastq1 is required" in m.stderr
