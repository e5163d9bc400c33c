"""GPU tier: the REFERENCE's own host program with the engine bound in. oracle/_ref/SOAPnuke_gpu is the reference
compiled from its sources with the five `virtual` keywords of tests/integration/build.sh and main() constructing
gpuPeProcess / gpuSeProcess (tests/integration/gpu_process.cpp): its reader threads, temp files, emission order,
update_stat and print_stat are the reference's, filter_*_fqs + stat_*_fqs go through include/snk_engine.h.
Clean FASTQ, trim files and all reports must equal the unmodified reference binary's, byte for byte."""
import concurrent.futures
import glob
import os
import subprocess
import threading

import pytest

import oracle_py as orc
import test_cli_gpu as tc
from helpers import ROOT, report_equal

pytestmark = pytest.mark.gpu
BOUND = os.path.join(ROOT, "oracle", "_ref", "SOAPnuke_gpu")

CASES = ["pe_cfg2_plain_T1", "pe_cfg2_plain_T4_multicycle", "pe_cfg2_gz_T3", "pe_polyg250", "pe_varlen_hardtrim",
         "pe_index_peinfo_seqtype0", "pe_index_seqtype1_fasta", "se_default", "se_adapter_T4", "se_fasta_gz",
         "trim_pe_T1", "trim_pe_peinfo_index_multicycle", "trim_se_fasta", "contam_pe_single", "gcontam_pe",
         "srna_trim_T1", "srna_hard_cfg"]

# The bound binary is the reference's host and sleeps in the same 5 s polling quanta (peprocess.cpp:3039): all cases are
# started together the first time one is needed (the engine serves concurrent processes), a test waits for its own.
_RUNS = {}
_LOCK = threading.Lock()


def _bound_run(name):
    case = next(c for c in tc._ALL_CASES if c["name"] == name)
    st = tc._reference_side(name)
    w = st["w"]
    m = subprocess.run([BOUND, case.get("module", "filter")] + st["base"] + ["-o", f"{w}/bound"] + case["flags"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    return st, m


def _bound_side(name):
    with _LOCK:
        if not _RUNS:
            pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(2, min(8, (os.cpu_count() or 2) // 2)))
            for n in CASES:
                _RUNS[n] = pool.submit(_bound_run, n)
    return _RUNS[name].result()


@pytest.mark.skipif(not (orc.have_reference() and os.path.exists(BOUND)), reason="reference binaries not available")
@pytest.mark.parametrize("name", CASES)
def test_reference_host_with_engine_binding(name):
    case = next(c for c in tc._ALL_CASES if c["name"] == name)
    pe = case["pe"]
    st, m = _bound_side(name)
    w, ext_out, r = st["w"], st["ext_out"], st["ref"]
    assert r.returncode == 0, r.stderr.decode()
    assert m.returncode == 0, m.stderr.decode()
    mates = (1, 2) if pe else (1,)
    for mate in mates:
        a, b = tc.read_maybe_gz(f"{w}/ref/c{mate}{ext_out}"), tc.read_maybe_gz(f"{w}/bound/c{mate}{ext_out}")
        assert a == b, f"{name}: clean fq{mate} differs ({len(a)} vs {len(b)} bytes)"
    if case.get("trim"):
        for mate in mates:
            a, b = tc.read_maybe_gz(f"{w}/ref/t{mate}.fq.gz"), tc.read_maybe_gz(f"{w}/bound/t{mate}.fq.gz")
            assert len(a) > 0 and a == b, f"{name}: trim fq{mate} differs ({len(a)} vs {len(b)} bytes)"
    reports = sorted(glob.glob(f"{w}/ref/*.txt"))
    assert len(reports) == (10 if pe else 6)
    for f in reports:
        assert report_equal(f, f"{w}/bound/{os.path.basename(f)}"), f"{name}: {os.path.basename(f)} differs"
