#!/bin/bash
# Session 8 (round 2), one GPU: writer variants on a 4 Mi-pair run (parent holds a CUDA context, like bench.py).
# (record of the experiment: SNK_WRITE_POPULATE and SNK_POOL_EXTRA existed only for this session and were removed afterwards - neither helped)
OUT=gpurun_out; mkdir -p $OUT
timeout 60 python __graft_entry__.py smoke > $OUT/exp8_smoke.log 2>&1; rc=$?; tail -1 $OUT/exp8_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 170 python - <<'PY' > gpurun_out/exp8_writer.txt 2>&1
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
variants = [("default", {}), ("populate", {"SNK_WRITE_POPULATE": "1"}), ("pool+8", {"SNK_POOL_EXTRA": "8"}), ("populate,pool+8", {"SNK_WRITE_POPULATE": "1", "SNK_POOL_EXTRA": "8"}),
            ("pwrite", {"SNK_WRITE_MMAP": "0"}), ("pwrite,pool+8", {"SNK_WRITE_MMAP": "0", "SNK_POOL_EXTRA": "8"}), ("populate,pool+16,w16", {"SNK_WRITE_POPULATE": "1", "SNK_POOL_EXTRA": "16", "SNK_WRITE_THREADS": "16"}),
            ("default", {})]
subprocess.run(cmd("ref_out"), env=dict(os.environ, SNK_WRITE_MMAP="0"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
for label, env in variants:
    walls = []
    for i in range(3):
        t0 = time.time()
        p = subprocess.run(cmd("out"), env=dict(os.environ, **env), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        walls.append(time.time() - t0)
    log = [l.strip() for l in open(f"{w}/out/log") if "stage seconds" in l or "timeline" in l]
    same = all(filecmp.cmp(f"{w}/out/{f}", f"{w}/ref_out/{f}", shallow=False) for f in os.listdir(f"{w}/ref_out") if f != "log")
    print(f"{label}: rc {p.returncode} walls {' '.join('%.3f' % x for x in walls)} identical {same}")
    for l in log: print("    ", l)
PY
cat gpurun_out/exp8_writer.txt
