"""Regenerates tests/golden/* by running the UNMODIFIED reference binary (oracle/_ref/SOAPnuke, built
from /root/reference by oracle/Makefile). Run in the build container only:

    python tests/golden/make_golden.py

Each case directory holds the input FASTQ (.gz), the CLI flags, and the reference's outputs
(clean FASTQ .gz + every report .txt). Tests compare the oracle / engine against these files.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from soapnuke_b200 import synth  # noqa: E402
from helpers import CFG2_FLAGS, CFG2_KW, A1  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "SOAPnuke")

CASES = {
    # name: (pe?, n, L, seed, gen kwargs, CLI flags, make_params kwargs, threads, config-file text)
    "pe_cfg2": (True, 1000, 100, 4242, dict(polyg_frac=0.1), CFG2_FLAGS, CFG2_KW, 2, "patch=5\n"),
    "pe_discard": (True, 600, 150, 4243, dict(), ["-f", A1, "-r", synth.ADAPTER2.decode()],
                   dict(adapter1=A1, adapter2=synth.ADAPTER2.decode()), 1, None),
    "se_default": (False, 1000, 100, 4244, dict(), [], dict(), 1, None),
    "se_trim_var": (False, 800, 120, 4245, dict(var_len=True), ["-f", A1, "-J", "-g", "8", "-4", "20"],
                    dict(adapter1=A1, ada_trim=True, polyG_tail=8, min_read_length=20), 3, "patch=3\n"),
    # filtersRNA module (generator gen_srna; params carry srna=True and the module defaults 18/49)
    "srna_trim": (False, 1000, 50, 4246, dict(var_len=True, srna=True),
                  ["-f", synth.SRNA_ADAPTER5.decode(), "-r", synth.SRNA_ADAPTER3.decode(), "-J", "-g", "6"],
                  dict(srna=True, adapter1=synth.SRNA_ADAPTER5.decode(), adapter2=synth.SRNA_ADAPTER3.decode(), ada_trim=True,
                       polyG_tail=6, min_read_length=18, max_read_length=49), 2, "patch=7\n"),
}


def main():
    only = set(sys.argv[1:])          # python make_golden.py [case ...]: regenerate only the named cases
    for name, (pe, n, L, seed, gkw, flags, pkw, T, cfg) in CASES.items():
        if only and name not in only:
            continue
        d = os.path.join(HERE, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        work = os.path.join("/tmp", "golden_" + name)
        shutil.rmtree(work, ignore_errors=True)
        os.makedirs(work)
        gkw = dict(gkw)
        module = "filtersRNA" if gkw.pop("srna", False) else "filter"
        data = synth.gen_srna(n, L=L, seed=seed, **gkw) if module == "filtersRNA" else synth.gen_pairs(n, L=L, seed=seed, se=not pe, **gkw)
        synth.write_fastq(os.path.join(work, "r1.fq"), data["seq1"], data["qual1"], data["len1"], 1)
        args = ["-1", os.path.join(work, "r1.fq"), "-C", "c1.fq", "-o", os.path.join(work, "out"), "-T", str(T)]
        if pe:
            synth.write_fastq(os.path.join(work, "r2.fq"), data["seq2"], data["qual2"], data["len2"], 2)
            args += ["-2", os.path.join(work, "r2.fq"), "-D", "c2.fq"]
        patch = None
        if cfg:
            with open(os.path.join(work, "cfg.txt"), "w") as f:
                f.write(cfg)
            args += ["-c", os.path.join(work, "cfg.txt")]
            patch = int(cfg.split("=")[1])
        subprocess.check_call([REF, module] + args + flags, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for m in (1, 2) if pe else (1,):
            for src, dst in ((os.path.join(work, f"r{m}.fq"), f"r{m}.fq.gz"), (os.path.join(work, "out", f"c{m}.fq"), f"c{m}.fq.gz")):
                with open(src, "rb") as fi, gzip.GzipFile(os.path.join(d, dst), "wb", compresslevel=9, mtime=0) as fo:
                    fo.write(fi.read())
        for f in sorted(os.listdir(os.path.join(work, "out"))):
            if f.endswith(".txt"):
                shutil.copy(os.path.join(work, "out", f), os.path.join(d, f))
        meta = dict(pe=pe, n=n, L=L, seed=seed, module=module, flags=flags, params=pkw, threads=T, patch_size=patch,
                    nprocs=os.cpu_count(), reference="SOAPnuke 2.1.9 @ 2d5b727, g++ -O3 -std=c++11 -include cstdint")
        with open(os.path.join(d, "case.json"), "w") as f:
            json.dump(meta, f, indent=1)
        print("golden", name, "written")


if __name__ == "__main__":
    main()
