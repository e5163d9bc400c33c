#!/bin/bash
# Session 7 (round 2), one GPU: the CLI tests after the read-ahead change, and its effect on a 4 Mi-pair run (parent holds a CUDA context, like bench.py).
OUT=gpurun_out; TAG=exp7; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_cli_gpu.py -m gpu -q > $OUT/${TAG}_pytest_cli.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_cli.log; tail -4 $OUT/${TAG}_pytest_cli.log
timeout 200 python - <<'PY' > gpurun_out/exp7_prefetch.txt 2>&1
import os, subprocess, sys, time, filecmp
sys.path.insert(0, os.getcwd())
from soapnuke_b200 import synth
w = "/dev/shm/snk_pf"; os.makedirs(w, exist_ok=True)
d = synth.gen_pairs(1 << 20, L=150, seed=1002)
for m in (1, 2):
    with open(f"{w}/r{m}.fq", "wb") as f:
        for k in range(4):
            synth.write_fastq_fixed(f"{w}/p.fq", d[f"seq{m}"], d[f"qual{m}"], 150, m, first=k << 20)
            f.write(open(f"{w}/p.fq", "rb").read())
A1, A2 = synth.ADAPTER1.decode(), synth.ADAPTER2.decode()
def cmd(out): return ["soapnuke_b200/bin/SOAPnuke", "filter", "-1", f"{w}/r1.fq", "-2", f"{w}/r2.fq", "-C", "c1.fq", "-D", "c2.fq", "-o", f"{w}/{out}", "-T", "16",
       "-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7", "-X", "50", "-g", "10", "-y", "20,30", "-x", "20,10"]
import torch
torch.zeros(1, device="cuda")
for label, env in (("read-ahead off", {"SNK_PREFETCH_MB": "0"}), ("read-ahead on", {}), ("read-ahead off", {"SNK_PREFETCH_MB": "0"}), ("read-ahead on", {})):
    for i in range(3):
        out = "out_off" if env else "out_on"
        t0 = time.time()
        p = subprocess.run(cmd(out), env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        t1 = time.time()
        log = [l.strip() for l in open(f"{w}/{out}/log") if "stage seconds" in l or "timeline" in l]
        print(f"{label}: rc {p.returncode} wall {t1 - t0:.3f}", p.stderr.decode()[-200:])
        for l in log: print("    ", l)
same = all(filecmp.cmp(f"{w}/out_on/{f}", f"{w}/out_off/{f}", shallow=False) for f in os.listdir(f"{w}/out_off") if f != "log")
print("outputs identical with and without read-ahead:", same)
PY
cat gpurun_out/exp7_prefetch.txt
