#!/usr/bin/env python
"""TEST INFRASTRUCTURE (needs a GPU): random option sets through the drop-in CLI against the unmodified reference
binary - the host driver's side of the story (text path on the device, emission order, gzip members, trim files,
id handling), on top of what tools/ref_fuzz.py checks for the oracle.

    python tools/cli_fuzz.py FIRST_SEED LAST_SEED
"""
import concurrent.futures
import glob
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, _p)
import oracle_py as orc  # noqa: E402
import ref_fuzz  # noqa: E402
from helpers import report_equal  # noqa: E402
from soapnuke_b200 import synth  # noqa: E402

CLI = os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke")


def rd(path):
    return gzip.open(path).read() if path.endswith(".gz") else open(path, "rb").read()


def prepare(seed):
    g = ref_fuzz.gen(seed)
    rnd = random.Random(seed * 7 + 1)
    pe, n, d, idfn = g["pe"], g["n"] * rnd.choice([1, 4]), g["d"], g["idfn"]
    if n != g["n"]:                          # a larger batch of the same kind: tile the generated reads
        import numpy as np
        reps = n // g["n"]
        d = {k: (np.tile(v, (reps, 1)) if getattr(v, "ndim", 0) == 2 else (np.tile(v, reps) if hasattr(v, "ndim") else v)) for k, v in d.items()}
    gz_in, gz_out = rnd.random() < 0.4, rnd.random() < 0.4
    cfg = list(g["cfg"])
    if g["module"] == "filter" and rnd.random() < 0.3 and pe: cfg.append("pe_info")
    # (never both: with fasta output the reference skips the quality conversion but still subtracts the OUTPUT Phred
    # base in the clean statistics - negative table indices, undefined behaviour)
    r_out, r_fa = rnd.random(), rnd.random()
    if not any(c.startswith("qualSys") for c in cfg):          # (the generator may already have chosen Phred+64 in / either out)
        if r_out < 0.2: cfg.append("outQualSys=1")
        elif r_fa < 0.25: cfg.append("outFileType=fasta")
    trim = rnd.random() < 0.3
    if trim: cfg += ["trimFq1=t1.fq.gz"] + (["trimFq2=t2.fq.gz"] if pe else [])
    env = dict(os.environ)
    if rnd.random() < 0.5: env["SNK_BATCH_READS"] = str(rnd.choice([512, 3000, 20000]))
    w = tempfile.mkdtemp(prefix="clifz")
    ei, eo = (".fq.gz" if gz_in else ".fq"), (".fq.gz" if gz_out else ".fq")
    synth.write_fastq(f"{w}/r1{ei}", d["seq1"], d["qual1"], d["len1"], 1, gz=gz_in, ids=idfn(n, 1) if idfn else None)
    base = ["-1", f"{w}/r1{ei}", "-C", "c1" + eo, "-T", str(g["T"])]
    if pe:
        synth.write_fastq(f"{w}/r2{ei}", d["seq2"], d["qual2"], d["len2"], 2, gz=gz_in, ids=idfn(n, 2) if idfn else None)
        base += ["-2", f"{w}/r2{ei}", "-D", "c2" + eo]
    if cfg:
        open(f"{w}/cfg.txt", "w").write("".join(l + "\n" for l in cfg))
        base += ["-c", f"{w}/cfg.txt"]
    flags = list(g["flags"])
    if seed in ref_fuzz.ADALIST:             # adapter list files instead of the literal adapters
        for opt, lst, fn in (("-f", ref_fuzz.ADALIST[seed][0], "ada1.list"), ("-r", ref_fuzz.ADALIST[seed][1], "ada2.list")):
            if lst:
                open(f"{w}/{fn}", "w").write("".join(a + "\n" for a in lst))
                flags[flags.index(opt) + 1] = f"{w}/{fn}"
    r = orc.run_reference(base + ["-o", f"{w}/ref"] + flags, module=g["module"])
    return dict(seed=seed, w=w, base=base, flags=flags, module=g["module"], pe=pe, eo=eo, trim=trim, env=env, ref=r, cfg=cfg)


def main():
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(2, (os.cpu_count() or 4) // 2)) as ex:
        for c in ex.map(prepare, range(lo, hi)):
            tag = f"{c['seed']} {c['module']} {' '.join(c['flags'])} {c['cfg']} {c['env'].get('SNK_BATCH_READS', '')}"
            if c["ref"].returncode != 0:
                print("REF-CRASH", tag); continue
            m = subprocess.run([CLI, c["module"]] + c["base"] + ["-o", f"{c['w']}/mine"] + c["flags"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=c["env"], timeout=600)
            if m.returncode != 0:
                print("MINE-FAILED", tag, m.stderr.decode()[-200:]); continue
            bad = []
            names = [f"c{k}{c['eo']}" for k in ((1, 2) if c["pe"] else (1,))] + ([f"t{k}.fq.gz" for k in ((1, 2) if c["pe"] else (1,))] if c["trim"] else [])
            for nm in names:
                if rd(f"{c['w']}/ref/{nm}") != rd(f"{c['w']}/mine/{nm}"): bad.append(nm)
            kept = len(rd(f"{c['w']}/ref/{names[0]}"))
            for f in glob.glob(f"{c['w']}/ref/*.txt"):
                if not report_equal(f, f"{c['w']}/mine/" + os.path.basename(f)):
                    if "Basic_Statistics" in f and kept == 0: continue      # the reference prints uninitialised buffers
                    bad.append(os.path.basename(f))
            if bad: print("MISMATCH", tag, bad, c["w"])
            else: shutil.rmtree(c["w"], ignore_errors=True)
    print("done")


if __name__ == "__main__":
    main()
