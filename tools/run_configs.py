#!/usr/bin/env python
"""BASELINE.json configs at their stated scale, one JSON line per config.

For each config:
  * a synthetic base block (the generator of SURVEY 8d, seed 1000 + config number) is rendered to /dev/shm;
  * PARITY on the sampled sub-range: the drop-in CLI and the unmodified reference binary filter the block with the same
    flags and -T; clean FASTQ and every report must be byte-identical;
  * CPU BASELINE: the reference binary, -T <host cores>, timed on a few blocks (the whole config would take hours on the
    CPU); reads/s are extrapolated linearly, as SURVEY 8d prescribes;
  * STATED SCALE: the CLI filters the full number of reads. The text of configs 2-5 does not fit any disk (config 3 is
    ~410 GB), so the block is replayed through FIFOs (`cat block block ...`), outputs go to /dev/null, reports to disk;
    SNK_GPUS=<gpus> shards the batches over the GPUs of the box;
  * KERNEL: the resident-batch rate of the filter kernel for the config's shape with its roofline fraction.

    python tools/run_configs.py --configs 1,2,4 --gpus 1 [--scale 0.1] >> profiles/r2_configs.jsonl
"""
import argparse
import ctypes as C
import glob
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
from soapnuke_b200 import abi, synth  # noqa: E402

A1, A2 = synth.ADAPTER1.decode(), synth.ADAPTER2.decode()
SA3 = synth.SRNA_ADAPTER3.decode()
CLI = os.environ.get("SNK_CLI", os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke"))
REF = os.path.join(ROOT, "oracle", "_ref", "SOAPnuke")
WORK = "/dev/shm/snk_configs"

CFG2_FLAGS = ["-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7", "-X", "50", "-g", "10", "-y", "20,30", "-x", "20,10"]
CFG2_KW = dict(adapter1=A1, adapter2=A2, ada_trim=True, low_qual=5, low_qual_ratio=0.5, n_ratio=0.05, mean_quality=15, highA_ratio=0.7,
               polyX_num=50, polyG_tail=10, trim_bad_tail=(20, 30), trim_bad_head=(20, 10))
CONFIGS = {
    1: dict(name="SE 150bp 1M reads, default N+lowQ filter", pe=False, L=150, units=1_000_000, gpus=1, block=1 << 19, flags=[], pkw=dict(),
            gkw=dict(seed=1001)),
    2: dict(name="PE 2x150bp 50M pairs, adapter trim + all quality filters", pe=True, L=150, units=50_000_000, gpus=1, block=1 << 19, flags=CFG2_FLAGS,
            pkw=CFG2_KW, gkw=dict(seed=1002)),
    3: dict(name="PE 2x150bp 628M pairs, full filter pipeline", pe=True, L=150, units=628_000_000, gpus=8, block=1 << 19, flags=CFG2_FLAGS,
            pkw=CFG2_KW, gkw=dict(seed=1003)),
    4: dict(name="SE 50bp sRNA-like 500M reads, adapter-dominant", pe=False, L=50, units=500_000_000, gpus=8, block=1 << 20,
            flags=["-f", SA3, "-J", "-4", "15"], pkw=dict(adapter1=SA3, ada_trim=True, min_read_length=15),
            gkw=dict(seed=1004, adapter1=synth.SRNA_ADAPTER3, insert_range=(15, 35))),
    5: dict(name="PE 2x250bp 200M pairs, polyG-tail trim (30% tails)", pe=True, L=250, units=200_000_000, gpus=8, block=1 << 18,
            flags=["-f", A1, "-r", A2, "-J", "-g", "10"], pkw=dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10),
            gkw=dict(seed=1005, polyg_frac=0.3)),
}


def sh(cmd, env=None, timeout=3600):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=timeout)
    return p, time.perf_counter() - t0


def digest(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def log_lines(path):
    try:
        return [l.strip() for l in open(path) if "seconds" in l or "gzip input" in l]
    except OSError:
        return []


def kernel_rate(cfg, n):
    """Resident batches through snk_filter_*_device: CUDA events on the launching stream, 3 warm-up + 5 timed launches."""
    import torch
    lib = abi.load_engine()
    pe, L = cfg["pe"], cfg["L"]
    base = synth.gen_pairs(1 << 17, L=L, se=not pe, **cfg["gkw"])
    reps = max(1, n // base["n"])
    n = reps * base["n"]
    dev = torch.device("cuda:0")
    t = {}
    for k, v in base.items():
        if isinstance(v, np.ndarray):
            a = np.tile(v, (reps, 1)) if v.ndim == 2 else np.tile(v, reps)
            t[k] = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(dev)
    p = abi.make_params(is_pe=pe, threads=8, nprocs=1 << 20, **cfg["pkw"])
    h = C.c_void_p()
    assert lib.snk_engine_create(C.byref(p), 0, C.byref(h)) == 0, lib.snk_last_error()
    out1 = torch.empty(n, dtype=torch.int64, device=dev)
    out2 = torch.empty(n, dtype=torch.int64, device=dev)
    b1 = abi.Batch(t["seq1"].data_ptr(), t["qual1"].data_ptr(), t["len1"].data_ptr(), n, base["stride"])
    b2 = abi.Batch(t["seq2"].data_ptr(), t["qual2"].data_ptr(), t["len2"].data_ptr(), n, base["stride"]) if pe else None
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step(i):
        if pe:
            rc = lib.snk_filter_pe_device(h, C.byref(b1), C.byref(b2), out1.data_ptr(), out2.data_ptr(), i * n, s)
        else:
            rc = lib.snk_filter_se_device(h, C.byref(b1), out1.data_ptr(), i * n, s)
        assert rc == 0, lib.snk_last_error()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        step(3 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    lib.snk_engine_destroy(h)
    reads = n * (2 if pe else 1)
    peak, src = 6650.0, "fallback of B200_PROFILING.md"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    gbs = reads * (2 * L + 8) / ms / 1e6
    return dict(value=reads / ms / 1e3, unit="Mreads/s", reads_per_launch=reads, launch_ms=ms,
                roofline=dict(bound="hbm", achieved=gbs, peak=peak, unit="GB/s", frac=gbs / peak, peak_source=src,
                              algorithmic_bytes_per_read=2 * L + 8))


def run_config(num, a):
    cfg = CONFIGS[num]
    pe, L, block = cfg["pe"], cfg["L"], cfg["block"]
    gpus = a.gpus or cfg["gpus"]
    units = int(cfg["units"] * a.scale)
    cores = os.cpu_count()
    T = min(cores, 48)
    w = os.path.join(WORK, f"cfg{num}")
    shutil.rmtree(w, ignore_errors=True)
    os.makedirs(w)
    out = dict(config=num, workload=f"BASELINE configs[{num - 1}]: {cfg['name']}", units="pairs" if pe else "reads", units_stated=cfg["units"],
               units_run=units, scale=a.scale, n_gpus=gpus, host_cores=cores, flags=" ".join(cfg["flags"]), data="synthetic (SURVEY 8d generator)")
    d = synth.gen_pairs(block, L=L, se=not pe, **cfg["gkw"])
    mates = (1, 2) if pe else (1,)
    for m in mates:
        synth.write_fastq_fixed(f"{w}/blk{m}.fq", d[f"seq{m}"], d[f"qual{m}"], L, m)
    block_bytes = sum(os.path.getsize(f"{w}/blk{m}.fq") for m in mates)

    def args(prefix, outdir, threads):
        x = ["-1", f"{prefix}1.fq", "-C", "c1.fq", "-o", outdir, "-T", str(threads)]
        if pe:
            x += ["-2", f"{prefix}2.fq", "-D", "c2.fq"]
        return x + cfg["flags"]

    # ---- parity on the block: CLI vs reference, same flags, same -T
    env1 = dict(os.environ, SNK_GPUS="1")
    if not a.no_reference and os.path.exists(REF):
        pt = 8
        r, ref_wall_blk = sh([REF, "filter"] + args(f"{w}/blk", f"{w}/ref", pt))
        m, _ = sh([CLI, "filter"] + args(f"{w}/blk", f"{w}/mine", pt), env=env1)
        same = r.returncode == 0 and m.returncode == 0
        if same:
            for k in mates:
                same = same and digest(f"{w}/ref/c{k}.fq") == digest(f"{w}/mine/c{k}.fq")
            reports = sorted(glob.glob(f"{w}/ref/*.txt"))
            same = same and len(reports) == (10 if pe else 6)
            for f in reports:
                same = same and open(f, "rb").read() == open(f"{w}/mine/{os.path.basename(f)}", "rb").read()
        out["parity"] = dict(sample_units=block, threads=pt, outputs_identical=bool(same), what="clean FASTQ + all reports, CLI vs oracle/_ref/SOAPnuke")
        if not same:
            out["parity"]["stderr"] = (r.stderr.decode()[-300:] + " | " + m.stderr.decode()[-300:])
        shutil.rmtree(f"{w}/ref", ignore_errors=True)
        shutil.rmtree(f"{w}/mine", ignore_errors=True)
        # ---- CPU baseline: the reference on a few blocks, all host cores
        nblk = max(1, min(a.ref_blocks, units // block))
        for k in mates:
            with open(f"{w}/ref{k}.fq", "wb") as f:
                for _ in range(nblk):
                    with open(f"{w}/blk{k}.fq", "rb") as g:
                        shutil.copyfileobj(g, f, 1 << 24)
        r, wall = sh([REF, "filter"] + args(f"{w}/ref", f"{w}/refout", T))
        reads = nblk * block * len(mates)
        out["cpu_baseline"] = dict(value=reads / wall / 1e6, unit="Mreads/s", cores=T, kind="reference", wall_s=wall, ok=r.returncode == 0,
                                   sample=f"SOAPnuke 2.1.9 filter -T {T} on {nblk * block} {out['units']} ({nblk} blocks), plain FASTQ on /dev/shm, "
                                          f"whole-program wall time; extrapolates to {cfg['units'] * len(mates) / (reads / wall) / 3600:.2f} h for the stated scale")
        if num == 1:
            r, wall1 = sh([REF, "filter"] + args(f"{w}/ref", f"{w}/refout1", 1))
            out["cpu_baseline_T1"] = dict(value=reads / wall1 / 1e6, unit="Mreads/s", cores=1, wall_s=wall1, ok=r.returncode == 0)
        shutil.rmtree(f"{w}/refout", ignore_errors=True)
        shutil.rmtree(f"{w}/refout1", ignore_errors=True)
        for k in mates:
            os.unlink(f"{w}/ref{k}.fq")

    # ---- stated scale through the CLI
    nblk = max(1, (units + block - 1) // block)
    total_units = nblk * block
    od = f"{w}/out"
    os.makedirs(od)
    env = dict(os.environ, SNK_GPUS=str(gpus))
    if a.batch_reads:
        env["SNK_BATCH_READS"] = str(a.batch_reads)
    fits = nblk * block_bytes < a.shm_budget_gb * 1e9
    feeders = []
    if fits and nblk <= 64:                                       # small enough for real files (config 1)
        for k in mates:
            with open(f"{w}/in{k}.fq", "wb") as f:
                for _ in range(nblk):
                    with open(f"{w}/blk{k}.fq", "rb") as g:
                        shutil.copyfileobj(g, f, 1 << 24)
        how = "plain FASTQ files on /dev/shm -> clean FASTQ on /dev/shm"
    else:
        for k in mates:
            os.mkfifo(f"{w}/in{k}.fq")
            os.symlink("/dev/null", f"{od}/c{k}.fq")
            # the FIFO is opened by the feeder's own shell: opening it here would block until the CLI opens the other end
            listing = f"{w}/feed{k}.txt"
            open(listing, "w").write((f"{w}/blk{k}.fq\n") * nblk)
            feeders.append(subprocess.Popen(f"xargs cat < {listing} > {w}/in{k}.fq", shell=True, start_new_session=True))
        how = f"block of {block} {out['units']} replayed {nblk}x through FIFOs -> clean FASTQ to /dev/null, reports to disk"
    try:
        m, wall = sh([CLI, "filter"] + args(f"{w}/in", od, T), env=env, timeout=a.timeout)
    except subprocess.TimeoutExpired:
        m, wall = subprocess.CompletedProcess([], 124, b"", b"timed out"), float(a.timeout)
    for f in feeders:                                             # a feeder that never found its reader is still blocked in open()
        if f.poll() is None:
            try:
                os.killpg(f.pid, 15)
            except OSError:
                pass
        f.wait()
    reads = total_units * len(mates)
    out["e2e_file"] = dict(value=reads / wall / 1e6, unit="Mreads/s", wall_s=wall, reads=reads, ok=m.returncode == 0, how=how, threads=T,
                           text_gb=nblk * block_bytes / 1e9, log=log_lines(f"{od}/log"))
    if m.returncode != 0:
        out["e2e_file"]["stderr"] = m.stderr.decode()[-500:]
    else:
        try:                                                      # the run's own report: raw read count must be what was fed
            rep = open(glob.glob(f"{od}/Basic_Statistics_of_Sequencing_Quality.txt")[0]).read().splitlines()
            out["e2e_file"]["report_total_reads_line"] = next(l for l in rep if l.lower().startswith("total number of reads"))
        except Exception:
            pass
    if not a.no_kernel:
        out["kernel"] = kernel_rate(cfg, a.kernel_units)
    shutil.rmtree(w, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2")
    ap.add_argument("--gpus", type=int, default=0, help="override the config's GPU count")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the stated number of reads to run (1.0 = stated scale)")
    ap.add_argument("--ref-blocks", type=int, default=8, help="blocks in the reference's timed sample")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--no-kernel", action="store_true")
    ap.add_argument("--kernel-units", type=int, default=1 << 21)
    ap.add_argument("--batch-reads", type=int, default=0)
    ap.add_argument("--shm-budget-gb", type=float, default=40.0)
    ap.add_argument("--timeout", type=int, default=1500)
    a = ap.parse_args()
    os.makedirs(WORK, exist_ok=True)
    for num in [int(x) for x in a.configs.split(",")]:
        print(json.dumps(run_config(num, a)), flush=True)
    shutil.rmtree(WORK, ignore_errors=True)


if __name__ == "__main__":
    main()
