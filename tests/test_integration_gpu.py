"""GPU tier: the REFERENCE's own host program with the engine bound in. oracle/_ref/SOAPnuke_gpu is the reference
compiled from its sources with the five `virtual` keywords of tests/integration/build.sh and main() constructing
gpuPeProcess / gpuSeProcess (tests/integration/gpu_process.cpp): its reader threads, temp files, emission order,
update_stat and print_stat are the reference's, filter_*_fqs + stat_*_fqs go through include/snk_engine.h.
Clean FASTQ, trim files and all reports must equal the unmodified reference binary's, byte for byte."""
import os

import pytest

import oracle_py as orc
import test_cli_gpu as tc
from helpers import ROOT

pytestmark = pytest.mark.gpu
BOUND = os.path.join(ROOT, "oracle", "_ref", "SOAPnuke_gpu")

CASES = ["pe_cfg2_plain_T1", "pe_cfg2_plain_T4_multicycle", "pe_cfg2_gz_T3", "pe_polyg250", "pe_varlen_hardtrim",
         "pe_index_peinfo_seqtype0", "pe_index_seqtype1_fasta", "se_default", "se_adapter_T4", "se_fasta_gz",
         "trim_pe_T1", "trim_pe_peinfo_index_multicycle", "trim_se_fasta", "contam_pe_single", "gcontam_pe",
         "srna_trim_T1", "srna_hard_cfg"]


@pytest.mark.skipif(not (orc.have_reference() and os.path.exists(BOUND)), reason="reference binaries not available")
@pytest.mark.parametrize("name", CASES)
def test_reference_host_with_engine_binding(tmp_path, name):
    case = dict(next(c for c in tc._ALL_CASES if c["name"] == name))
    case.pop("env", None)                      # SNK_BATCH_READS is a knob of the drop-in CLI only
    tc.run_both(BOUND, tmp_path, out_name="bound", **case)
