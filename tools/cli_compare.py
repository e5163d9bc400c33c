#!/usr/bin/env python
"""Whole-program comparison on one box: `soapnuke_b200/bin/SOAPnuke filter` vs the unmodified
reference binary (oracle/_ref/SOAPnuke) on the same synthetic PE150 FASTQ files in /dev/shm.
Prints wall/user/sys for both and checks the outputs are byte-identical (decompressed).

    python tools/cli_compare.py [--pairs 4000000] [--gz] [--threads N]
"""
import argparse
import gzip
import hashlib
import json
import os
import resource
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from soapnuke_b200 import synth  # noqa: E402
import numpy as np  # noqa: E402

A1, A2 = synth.ADAPTER1.decode(), synth.ADAPTER2.decode()
FLAGS = ["-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7", "-X", "50", "-g", "10",
         "-y", "20,30", "-x", "20,10"]


def digest(path):
    h = hashlib.sha256()
    f = gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")
    while True:
        b = f.read(1 << 24)
        if not b:
            break
        h.update(b)
    return h.hexdigest()


def timed(cmd, env=None):
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    wall = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    if p.returncode:
        raise SystemExit(f"{cmd[0]} failed: {p.stderr.decode()[-400:]}")
    return dict(wall=wall, user=r1.ru_utime - r0.ru_utime, sys=r1.ru_stime - r0.ru_stime)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=4000000)
    ap.add_argument("--gz", action="store_true")
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--batch-sweep", default="", help="comma list of SNK_BATCH_READS values to time the B200 CLI with")
    ap.add_argument("--env-sweep", default="", help="';'-separated settings, each a ','-separated list of VAR=value, to time the B200 CLI with")
    ap.add_argument("--gz-members", type=int, default=0, help="with --gz: compress the input in members of this many MiB of text (0 = one member per file)")
    a = ap.parse_args()
    work = "/dev/shm/snk_cli_compare"
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    ext = ".fq.gz" if a.gz else ".fq"
    unique = min(a.pairs, 1 << 20)
    d = synth.gen_pairs(unique, L=150, seed=1002)
    for m in (1, 2):
        path = f"{work}/r{m}.fq"
        with open(path, "wb") as f:
            for k in range(0, a.pairs, unique):
                n = min(unique, a.pairs - k)
                tmp = f"{work}/part.fq"
                synth.write_fastq_fixed(tmp, d[f"seq{m}"][:n], d[f"qual{m}"][:n], 150, m, first=k)
                with open(tmp, "rb") as g:
                    shutil.copyfileobj(g, f, 1 << 24)
                os.unlink(tmp)
        if a.gz and a.gz_members:
            subprocess.check_call(["split", "-b", f"{a.gz_members}M", "-d", "-a", "5", path, path + ".part."])
            os.unlink(path)
            subprocess.check_call(f"ls {path}.part.* | xargs -P {os.cpu_count()} -n 4 gzip -2", shell=True)
            subprocess.check_call(f"cat {path}.part.*.gz > {path}.gz && rm {path}.part.*.gz", shell=True)
        elif a.gz:
            subprocess.check_call(["gzip", "-2", path])
    base = ["-1", f"{work}/r1{ext}", "-2", f"{work}/r2{ext}", "-C", "c1" + ext, "-D", "c2" + ext, "-T", str(a.threads)]
    out = {"pairs": a.pairs, "gz": a.gz, "threads": a.threads, "cores": os.cpu_count()}
    mine = timed([os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke"), "filter"] + base + ["-o", f"{work}/mine"] + FLAGS)
    out["b200"] = dict(mine, mreads_per_s=2 * a.pairs / mine["wall"] / 1e6)
    if not a.skip_reference:
        ref = timed([os.path.join(ROOT, "oracle", "_ref", "SOAPnuke"), "filter"] + base + ["-o", f"{work}/ref"] + FLAGS)
        out["reference"] = dict(ref, mreads_per_s=2 * a.pairs / ref["wall"] / 1e6)
        out["speedup_wall"] = ref["wall"] / mine["wall"]
        same = all(digest(f"{work}/mine/c{m}{ext}") == digest(f"{work}/ref/c{m}{ext}") for m in (1, 2))
        for f in sorted(os.listdir(f"{work}/ref")):
            if f.endswith(".txt"):
                same = same and open(f"{work}/ref/{f}", "rb").read() == open(f"{work}/mine/{f}", "rb").read()
        out["outputs_identical"] = same
    if a.batch_sweep:
        out["batch_sweep"] = {}
        for br in a.batch_sweep.split(","):
            env = dict(os.environ, SNK_BATCH_READS=br)
            shutil.rmtree(f"{work}/sweep", ignore_errors=True)
            t = timed([os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke"), "filter"] + base + ["-o", f"{work}/sweep"] + FLAGS, env=env)
            t["log"] = [l.strip() for l in open(f"{work}/sweep/log") if "stage seconds" in l]
            out["batch_sweep"][br] = t
        shutil.rmtree(f"{work}/sweep", ignore_errors=True)
    if a.env_sweep:
        out["env_sweep"] = {}
        for setting in a.env_sweep.split(";"):
            env = dict(os.environ)
            env.update(kv.split("=", 1) for kv in setting.split(",") if kv)
            shutil.rmtree(f"{work}/sweep", ignore_errors=True)
            t = timed([os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke"), "filter"] + base + ["-o", f"{work}/sweep"] + FLAGS, env=env)
            t["mreads_per_s"] = 2 * a.pairs / t["wall"] / 1e6
            t["log"] = [l.strip() for l in open(f"{work}/sweep/log") if "stage seconds" in l or "gzip input" in l]
            out["env_sweep"][setting] = t
        shutil.rmtree(f"{work}/sweep", ignore_errors=True)
    try:
        out["b200_log"] = [l.strip() for l in open(f"{work}/mine/log") if "stage seconds" in l]
    except Exception:
        pass
    print(json.dumps(out))
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
