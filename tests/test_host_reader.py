"""CPU tier: the reader of the host pipeline (ByteSource / nth_newline / count_newlines of soapnuke_b200/host/process.cpp) -
parallel page-cache copies with per-share newline counts, and the read-ahead that runs while the CUDA contexts are created -
returns exactly the file's bytes for any request size, thread count and read-ahead limit. Replaces the line loop of
sub_thread (peprocess.cpp:2198-2239) on the input side; needs no GPU (the engine library only has to load)."""
import os
import subprocess

import pytest

from helpers import ROOT, synth


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    from soapnuke_b200 import build
    build.build_engine()
    out = str(tmp_path_factory.mktemp("hosttest") / "bytesource_test")
    host = os.path.join(ROOT, "soapnuke_b200", "host")
    lib = os.path.join(ROOT, "soapnuke_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", "-pthread", "-I/usr/local/cuda/include", "-I" + host, "-o", out,
                           os.path.join(ROOT, "tests", "hosttest", "bytesource_test.cpp"), os.path.join(host, "cli_params.cpp"),
                           os.path.join(host, "gz_members.cpp"), os.path.join(host, "fast_deflate.cpp"),
                           "-L" + lib, "-lsnk_engine", "-lz", "-ldl", "-Wl,-rpath," + lib])
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_reader_returns_the_file(driver, tmp_path, seed):
    d = synth.gen_pairs(150000, L=150, seed=40 + seed, se=True)
    path = str(tmp_path / "r1.fq")
    synth.write_fastq(path, d["seq1"], d["qual1"], d["len1"], 1)
    p = subprocess.run([driver, path, str(seed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stdout.decode()[-500:] + p.stderr.decode()[-500:]
    assert b"nth_newline ok" in p.stdout
