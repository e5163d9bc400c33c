#!/bin/bash
# Session 3 (round 2): reference-host binding tests, BASELINE configs 1, 2, 4 at stated scale on one GPU, start/exit cost diagnosis.
OUT=gpurun_out; TAG=exp3; mkdir -p $OUT
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 600 python -m pytest tests/test_integration_gpu.py -m gpu -q > $OUT/${TAG}_pytest_integration.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_integration.log; tail -5 $OUT/${TAG}_pytest_integration.log
timeout 900 python tools/run_configs.py --configs 1,2,4 --gpus 1 > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err; cut -c1-1600 $OUT/${TAG}_configs.jsonl; tail -3 $OUT/${TAG}_configs.err
# where do the seconds outside the CLI's own clock go?
python - <<'PY' > gpurun_out/exp3_timestamps.txt 2>&1
import os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
from soapnuke_b200 import synth
w = "/dev/shm/snk_ts"; os.makedirs(w, exist_ok=True)
d = synth.gen_pairs(1 << 20, L=150, seed=1002)
for m in (1, 2):
    with open(f"{w}/r{m}.fq", "wb") as f:
        for k in range(4):
            synth.write_fastq_fixed(f"{w}/p.fq", d[f"seq{m}"], d[f"qual{m}"], 150, m, first=k << 20)
            f.write(open(f"{w}/p.fq", "rb").read())
A1, A2 = synth.ADAPTER1.decode(), synth.ADAPTER2.decode()
cmd = ["soapnuke_b200/bin/SOAPnuke", "filter", "-1", f"{w}/r1.fq", "-2", f"{w}/r2.fq", "-C", "c1.fq", "-D", "c2.fq", "-o", f"{w}/out", "-T", "16",
       "-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7", "-X", "50", "-g", "10", "-y", "20,30", "-x", "20,10"]
import torch
for hold in (False, True):
    if hold:
        torch.zeros(1, device="cuda")          # the parent keeps a CUDA context open, like bench.py does
    for i in range(3):
        t0 = time.time()
        p = subprocess.run(cmd, env=dict(os.environ, SNK_TIMESTAMPS="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        t1 = time.time()
        ts = {l.split()[1]: float(l.split()[2]) for l in p.stderr.decode().splitlines() if l.startswith("snk-ts")}
        log = [l.strip() for l in open(f"{w}/out/log") if "seconds" in l]
        print(f"parent holds context={hold} run {i}: wall {t1 - t0:.3f}  spawn->main {ts['main'] - t0:.3f}  main->exit {ts['exit'] - ts['main']:.3f}  exit->reaped {t1 - ts['exit']:.3f}")
        print("   ", log)
PY
cat gpurun_out/exp3_timestamps.txt
