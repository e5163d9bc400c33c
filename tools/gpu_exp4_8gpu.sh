#!/bin/bash
# Session 4 (round 2), 8 GPUs of one box: N>1 parity tests, SNK_GPUS=8 vs SNK_GPUS=1 byte equality, then BASELINE configs 3, 5 and 4
# at their stated scale through the CLI with SNK_GPUS=8 (streamed input, see tools/run_configs.py). The reference / kernel legs of
# these shapes were measured on 1-GPU boxes (profiles/r2_configs_2_4_5_1gpu.jsonl).
OUT=gpurun_out; TAG=exp4; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_box.txt; nproc >> $OUT/${TAG}_box.txt; free -g >> $OUT/${TAG}_box.txt; df -h /dev/shm >> $OUT/${TAG}_box.txt
timeout 120 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; rc=$?; tail -1 $OUT/${TAG}_smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed rc=$rc: stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi.log; tail -4 $OUT/${TAG}_pytest_multi.log
timeout 200 python - <<'PY' > $OUT/${TAG}_gpus8_vs_1.txt 2>&1
import filecmp, glob, os, subprocess, sys, time
sys.path.insert(0, os.getcwd())
from soapnuke_b200 import synth
w = "/dev/shm/snk_g8"; os.makedirs(w, exist_ok=True)
d = synth.gen_pairs(400000, L=150, seed=1003)
for m in (1, 2):
    synth.write_fastq_fixed(f"{w}/r{m}.fq", d[f"seq{m}"], d[f"qual{m}"], 150, m)
A1, A2 = synth.ADAPTER1.decode(), synth.ADAPTER2.decode()
base = ["soapnuke_b200/bin/SOAPnuke", "filter", "-1", f"{w}/r1.fq", "-2", f"{w}/r2.fq", "-C", "c1.fq", "-D", "c2.fq", "-T", "8",
        "-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7", "-X", "50", "-g", "10", "-y", "20,30", "-x", "20,10"]
for g in (1, 8):
    t0 = time.time()
    p = subprocess.run(base + ["-o", f"{w}/out{g}"], env=dict(os.environ, SNK_GPUS=str(g), SNK_BATCH_READS="8192"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    print(f"SNK_GPUS={g}: rc {p.returncode} wall {time.time() - t0:.2f}s", p.stderr.decode()[-300:])
    print("   ", [l.strip() for l in open(f"{w}/out{g}/log") if "seconds" in l])
same = all(filecmp.cmp(f"{w}/out1/{os.path.basename(f)}", f, shallow=False) for f in glob.glob(f"{w}/out8/*.txt") + glob.glob(f"{w}/out8/c?.fq"))
print("reports + clean FASTQ of SNK_GPUS=8 byte-identical to SNK_GPUS=1:", same, "files compared:", len(glob.glob(f"{w}/out8/*.txt")) + 2)
PY
cat $OUT/${TAG}_gpus8_vs_1.txt
timeout 600 python tools/run_configs.py --configs 3,5,4 --no-reference --no-kernel --timeout 240 > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err
cut -c1-1700 $OUT/${TAG}_configs.jsonl; tail -3 $OUT/${TAG}_configs.err
