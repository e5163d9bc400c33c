"""CPU tier: the drop-in CLI's argument handling needs no GPU. Invalid invocations must fail the way the reference
does (`Error:...` on stderr, exit status 1, same message), the options this engine does not serve must be refused
explicitly, and a valid invocation without a CUDA device must fail loudly: there is no CPU fallback."""
import os
import subprocess

import pytest

import oracle_py as orc
from helpers import ROOT

CLI = os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke")


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    from soapnuke_b200 import build
    build.build_all()
    w = tmp_path_factory.mktemp("cliargs")
    (w / "a.fq").write_text("@r1/1\nACGT\n+\nIIII\n")
    (w / "b.fq").write_text("@r1/2\nACGT\n+\nIIII\n")
    return w


def run(binary, args, cwd, env=None):
    r = subprocess.run([binary] + args, cwd=str(cwd), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120, env=env)
    lines = r.stderr.decode().splitlines()
    return r.returncode, (lines[0] if lines else "")


def cfg(work, name, text):
    (work / name).write_text(text)
    return ["-c", name]


SAME_AS_REFERENCE = [
    ("no_fq1", ["filter", "-C", "c.fq", "-o", "o"], None),
    ("no_outdir", ["filter", "-1", "a.fq", "-C", "c.fq"], None),
    ("no_clean", ["filter", "-1", "a.fq", "-o", "o"], None),
    ("same_inputs", ["filter", "-1", "a.fq", "-2", "a.fq", "-C", "c.fq", "-D", "d.fq", "-o", "o"], None),
    ("trim_fields", ["filter", "-1", "a.fq", "-2", "b.fq", "-C", "c.fq", "-D", "d.fq", "-o", "o", "-t", "1,2,3"], None),
    ("seq_type", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "seqType=2\n"),
    ("qual_sys", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "qualSys=3\n"),
    ("unknown_key", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "nosuch=3\n"),
    ("adapter2_for_se", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o", "-r", "ACGT"], None),
    ("clean2_for_se", ["filter", "-1", "a.fq", "-C", "c.fq", "-D", "d.fq", "-o", "o"], None),
    ("se_bad_head_trim", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o", "-x", "20"], None),
    ("no_such_module", ["nosuchmodule", "-1", "a.fq"], None),
    ("srna_only_key_in_filter", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "adaRCtg=7\n"),
    ("filter_only_key_in_srna", ["filtersRNA", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "adaMis=1\n"),
]


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary oracle/_ref/SOAPnuke not built")
@pytest.mark.parametrize("case", SAME_AS_REFERENCE, ids=[c[0] for c in SAME_AS_REFERENCE])
def test_invalid_invocations_fail_like_the_reference(work, case):
    name, args, cfgtext = case
    if cfgtext:
        args = args + cfg(work, name + ".txt", cfgtext)
    mine = run(CLI, args, work)
    ref = run(orc.REF_BIN, args, work)
    assert mine[0] == 1 and ref[0] == 1
    assert mine[1] == ref[1] and mine[1].startswith("Error:")


REFUSED = [
    ("rmdup", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "rmdup\n", "does not implement"),
    ("output_split", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o", "-w", "100"], None, "does not implement"),
    ("streaming", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o", "-j"], None, "does not implement"),
    ("tile_range", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "tile=1101-1104\n", "tile ranges"),
    ("plain_trim_file", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "trimFq1=t1.fq\n", "must end with .gz"),
    ("srna_paired", ["filtersRNA", "-1", "a.fq", "-2", "b.fq", "-C", "c.fq", "-D", "d.fq", "-o", "o"], None, "paired input"),
    ("contam_ratio_count", ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], "contam1=ACGTACGT,GGGGCCCC\nctMatchR=0.2\n", "ctMatchR"),
    ("stlfr_module", ["filterStLFR", "-1", "a.fq"], None, "not served"),
]


@pytest.mark.parametrize("case", REFUSED, ids=[c[0] for c in REFUSED])
def test_unserved_options_are_refused_explicitly(work, case):
    name, args, cfgtext, needle = case
    if cfgtext:
        args = args + cfg(work, name + ".txt", cfgtext)
    rc, msg = run(CLI, args, work)
    assert rc == 1 and msg.startswith("Error:") and needle in msg, msg


def test_no_cuda_device_is_a_loud_failure(work):
    """The product has no CPU fallback: with no visible CUDA device a valid invocation stops with an error."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    rc, msg = run(CLI, ["filter", "-1", "a.fq", "-C", "c.fq", "-o", "o"], work, env=env)
    assert rc == 1 and msg.startswith("Error:") and "no CPU fallback" in msg, msg
