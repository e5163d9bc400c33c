#!/usr/bin/env python
"""bench.py — M reads/s of the FASTQ filter hot path on synthetic PE150 (BASELINE config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]

One "step" = one pass of the hot path (adapter match, trim, predicates, discard cascade, raw+clean
per-position histograms) over one batch of P read pairs per GPU with BASELINE config-2 flags
(`-f A1 -r A2 -J -l 5 -q 0.5 -n 0.05 -m 15 -p 0.7 -X 50 -g 10 -y 20,30 -x 20,10`).

 value     whole-job M reads/s with the batches already resident in HBM: K kernel launches per GPU and, for
           N > 1, the path's single collective (NCCL all-reduce of the statistics table) INSIDE the region timed
           with CUDA events on the launching stream, max over ranks.
 e2e       FILE TO FILE, what a user of the reference runs: `soapnuke_b200/bin/SOAPnuke filter` (one process per
           GPU) on the same plain FASTQ files in /dev/shm the reference arm reads, wall clock per run including
           process start, CUDA context, FASTQ parse on the device, clean FASTQ + report files written; at N = 1
           the outputs are compared byte for byte with the reference binary's.
 e2e_soa   pinned SoA host buffers -> 8-byte result records through snk_filter_pe_async (host<->device copies in the
           timed region), with the measured pinned host->device copy bandwidth of the box beside it (`h2d_probe`).
 e2e_text  raw FASTQ text in pinned memory -> clean FASTQ text in pinned memory (snk_filter_pe_text_async).
 roofline  algorithmic bytes (2L+8 per read) / average kernel duration vs the measured HBM copy
           bandwidth of MEASURED_PEAKS.json.
 stats_parity  outside the timed region: every rank filters a bounded slice of its own data, the tables are
           all-reduced, and rank 0 compares the result bit for bit with the oracle run over all ranks' slices.
 cpu_baseline  the unmodified reference binary (oracle/_ref/SOAPnuke, `filter -T <ncores>`) on this box's cores on the
           SAME files as `e2e` (rank 0, N=1 only): wall, user+sys and the number of 5 s poll quanta in the wall.

`--impl reference` times only that CPU reference (K+W runs) and prints the same JSON shape with the same `config`.
"""
import argparse
import ctypes as C
import filecmp
import json
import os
import resource
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from soapnuke_b200 import abi, synth  # noqa: E402

A1 = synth.ADAPTER1.decode()
A2 = synth.ADAPTER2.decode()
CFG2_FLAGS = ["-f", A1, "-r", A2, "-J", "-l", "5", "-q", "0.5", "-n", "0.05", "-m", "15", "-p", "0.7",
              "-X", "50", "-g", "10", "-y", "20,30", "-x", "20,10"]
CFG2_KW = dict(adapter1=A1, adapter2=A2, ada_trim=True, low_qual=5, low_qual_ratio=0.5, n_ratio=0.05,
               mean_quality=15, highA_ratio=0.7, polyX_num=50, polyG_tail=10, trim_bad_tail=(20, 30),
               trim_bad_head=(20, 10))
L = 150
METRIC = "Mreads/sec PE150 filter (adapter trim + all quality filters + raw/clean statistics)"
UNIT = "Mreads/s"
UNIQUE_PAIRS = 1 << 18          # generated once, tiled to the batch size
CLI = os.path.join(ROOT, "soapnuke_b200", "bin", "SOAPnuke")
PARITY_PAIRS = 32768            # per rank, stats_parity leg


def workload_config(pairs):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": "BASELINE configs[1]: PE 2x150bp, adapter trim + all quality filters (config-2 flags), one batch per GPU per step",
            "pairs_per_step_per_gpu": pairs, "read_len": L, "flags": " ".join(CFG2_FLAGS[4:]),
            "l2": "inputs larger than L2 (%.2f GB of rows per step per GPU)" % (4 * pairs * synth.stride_for(L) / 1e9)}


def measured_peak():
    """HBM copy bandwidth the roofline is reported against: the driver-written MEASURED_PEAKS.json when present (the
    kernel is timed alone, so a burst figure is preferred over a sustained one), else the profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        flat = {}

        def walk(prefix, obj):
            if isinstance(obj, dict):
                for k, v in obj.items():
                    walk(f"{prefix}.{k}" if prefix else str(k), v)
            elif isinstance(obj, (int, float)) and not isinstance(obj, bool):
                flat[prefix] = float(obj)
        walk("", peaks)
        hbm = {k: v for k, v in flat.items() if "hbm" in k.lower() and v > 100}      # GB/s figures, not fractions
        for test in (lambda k: k == "hbm_gbs", lambda k: "burst" in k.lower(), lambda k: True):
            for k in sorted(hbm):
                if test(k):
                    return hbm[k], f"measured (MEASURED_PEAKS.json {k})"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per read of the dominant kernel from the committed `ncu --set full` capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_host_batch(pairs, seed):
    """Synthetic PE150 batch of `pairs` pairs (config-2 mix), tiled from UNIQUE_PAIRS generated pairs."""
    base = synth.gen_pairs(min(pairs, UNIQUE_PAIRS), L=L, seed=seed)
    reps = (pairs + base["n"] - 1) // base["n"]
    out = {}
    for k, v in base.items():
        if isinstance(v, np.ndarray):
            t = np.tile(v, (reps, 1)) if v.ndim == 2 else np.tile(v, reps)
            out[k] = np.ascontiguousarray(t[:pairs])
        else:
            out[k] = v
    out["n"] = pairs
    return out


# --------------------------------------------------------------------------- the FASTQ files both arms run on
def work_dir(need_bytes):
    """A scratch directory with room for `need_bytes`: /dev/shm when it is large enough, else the default tmp."""
    for base in ("/dev/shm", None):
        try:
            if base is None or (os.path.isdir(base) and shutil.disk_usage(base).free > need_bytes + (1 << 30)):
                return tempfile.mkdtemp(prefix="snkbench_", dir=base)
        except Exception:
            continue
    return tempfile.mkdtemp(prefix="snkbench_")


def write_workload_files(work, pairs, seed=1002):
    """r1.fq / r2.fq: `pairs` synthetic PE150 pairs (config-2 mix): min(pairs, 2^20) generated pairs tiled with running ids."""
    unique = min(pairs, 1 << 20)
    d = synth.gen_pairs(unique, L=L, seed=seed)
    for m in (1, 2):
        with open(f"{work}/r{m}.fq", "wb") as f:
            for k in range(0, pairs, unique):
                n = min(unique, pairs - k)
                synth.write_fastq_fixed(f"{work}/part.fq", d[f"seq{m}"][:n], d[f"qual{m}"][:n], L, m, first=k)
                with open(f"{work}/part.fq", "rb") as g:
                    shutil.copyfileobj(g, f, 1 << 24)
                os.unlink(f"{work}/part.fq")
    return os.path.getsize(f"{work}/r1.fq") + os.path.getsize(f"{work}/r2.fq")


def run_timed(cmd, env=None, timeout=3600):
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=timeout)
    wall = time.perf_counter() - t0
    r1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    if p.returncode != 0:
        raise RuntimeError(f"{os.path.basename(cmd[0])} failed: " + p.stderr.decode()[-400:])
    return {"wall": wall, "cpu": (r1.ru_utime - r0.ru_utime) + (r1.ru_stime - r0.ru_stime)}


def file_args(work, out):
    return ["-1", f"{work}/r1.fq", "-2", f"{work}/r2.fq", "-C", "c1.fq", "-D", "c2.fq", "-o", out]


# --------------------------------------------------------------------------- reference arm
def reference_runs(work, out, runs, cores):
    """`SOAPnuke filter -T cores` (unmodified reference, or the oracle port when it did not build) on the files in `work`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    res = []
    for _ in range(runs):
        shutil.rmtree(out, ignore_errors=True)
        res.append(run_timed([orc.REF_BIN, "filter"] + file_args(work, out) + ["-T", str(cores)] + CFG2_FLAGS))
    return res


def have_reference():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "SOAPnuke"))


def describe_reference(pairs, cores, runs):
    wall = float(np.mean([r["wall"] for r in runs])); cpu = float(np.mean([r["cpu"] for r in runs]))
    return (f"SOAPnuke 2.1.9 filter -T {cores} on {pairs} synthetic PE150 pairs, plain FASTQ in/out on /dev/shm, whole-program wall time "
            f"{wall:.2f} s per run = {wall / 5.0:.1f} of its 5 s concat-poll quanta (peprocess.cpp:2769,3039), user+sys {cpu:.1f} CPU-s per run")


def time_oracle_port(pairs, steps, warmup, seed=1002):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    p = abi.make_params(is_pe=True, **CFG2_KW)
    d = synth.gen_pairs(min(pairs, 1 << 19), L=L, seed=seed)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        orc.filter_pe(p, d)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    return 2.0 * d["n"] / (sum(times) / len(times)) / 1e6, d["n"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = args.pairs
    cores = os.cpu_count() or 1
    if have_reference():
        work = work_dir(2 * pairs * 2 * (L + 20) * 2)
        try:
            write_workload_files(work, pairs)
            runs = reference_runs(work, f"{work}/out_ref", args.warmup + args.steps, cores)[args.warmup:]
        finally:
            shutil.rmtree(work, ignore_errors=True)
        wall = float(np.mean([r["wall"] for r in runs]))
        value = 2.0 * pairs / wall / 1e6
        base = {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": describe_reference(pairs, cores, runs),
                "wall_s_per_run": wall, "cpu_s_per_run": float(np.mean([r["cpu"] for r in runs])), "poll_quanta_per_run": wall / 5.0}
        ms = 1e3 * wall
    else:
        value, n = time_oracle_port(pairs, args.steps, args.warmup)
        base = {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"oracle C restatement, 1 thread, {n} synthetic PE150 pairs already parsed in memory"}
        ms = 1e3 * 2.0 * n / (value * 1e6)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(pairs),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from soapnuke_b200 import build
    from soapnuke_b200 import dist as snkdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the filter engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build_all()
    if world > 1:
        dist.barrier()
    lib = abi.load_engine()

    pairs = args.pairs
    params = abi.make_params(is_pe=True, threads=8, nprocs=1 << 20, **CFG2_KW)   # reference -T 8 partition: 400 k-pair blocks, 8 slots
    h = C.c_void_p()
    if lib.snk_engine_create(C.byref(params), local_rank, C.byref(h)):
        raise SystemExit("engine: " + lib.snk_last_error().decode())

    def check(rc):
        if rc:
            raise RuntimeError(lib.snk_last_error().decode())

    host = make_host_batch(pairs, seed=1002 + rank)
    stride = host["stride"]
    # pinned host copies (e2e_soa leg) and device-resident copies (value leg)
    pin = {}
    devt = {}
    for k in ("seq1", "qual1", "seq2", "qual2", "len1", "len2"):
        a = host[k]
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
        pin[k] = t.pin_memory()
        devt[k] = pin[k].to(dev, non_blocking=True)
    out_dev = [torch.empty(pairs, dtype=torch.int64, device=dev) for _ in range(2)]
    out_pin = [torch.empty(pairs, dtype=torch.int64).pin_memory() for _ in range(2)]
    torch.cuda.synchronize()

    db1 = abi.Batch(devt["seq1"].data_ptr(), devt["qual1"].data_ptr(), devt["len1"].data_ptr(), pairs, stride)
    db2 = abi.Batch(devt["seq2"].data_ptr(), devt["qual2"].data_ptr(), devt["len2"].data_ptr(), pairs, stride)
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    table = torch.empty(params.n_slots * abi.SLOT_WORDS, dtype=torch.int64, device=dev)

    def step_resident(i):
        check(lib.snk_filter_pe_device(h, C.byref(db1), C.byref(db2), out_dev[0].data_ptr(), out_dev[1].data_ptr(),
                                       C.c_uint64(i * pairs), sptr))

    def reduce_table():
        """The single collective of the path: the final statistics table (SUM; LAST_KEY words MAX)."""
        check(lib.snk_engine_stats_to_device(h, table.data_ptr(), sptr))
        if world > 1:
            snkdist.allreduce_stats(table, params.n_slots)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    # ---- value leg: device-resident batches + the final collective, CUDA events on the launching stream
    for i in range(args.warmup):
        step_resident(i)
    reduce_table()                      # warm-up of the collective (NCCL channel setup is not part of a step)
    barrier()
    launches0 = lib.snk_engine_launch_count(h)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    evk = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step_resident(args.warmup + i)
    evk.record(stream)
    reduce_table()
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    kernel_ms = ev0.elapsed_time(evk)
    collective_us = 1e3 * evk.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.snk_engine_launch_count(h) - launches0
    total_ms = max_over_ranks(total_ms)
    kernel_ms = max_over_ranks(kernel_ms)
    collective_us = max_over_ranks(collective_us)
    reads_per_step = 2 * pairs * world
    value = reads_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # ---- stats_parity leg (untimed): bounded slice per rank through the engine, all-reduced, vs the oracle on rank 0
    npar = min(PARITY_PAIRS, pairs)
    check(lib.snk_engine_stats_reset(h))
    pb1 = abi.Batch(devt["seq1"].data_ptr(), devt["qual1"].data_ptr(), devt["len1"].data_ptr(), npar, stride)
    pb2 = abi.Batch(devt["seq2"].data_ptr(), devt["qual2"].data_ptr(), devt["len2"].data_ptr(), npar, stride)
    check(lib.snk_filter_pe_device(h, C.byref(pb1), C.byref(pb2), out_dev[0].data_ptr(), out_dev[1].data_ptr(), C.c_uint64(rank * npar), sptr))
    reduce_table()
    torch.cuda.synchronize()
    gathered = {}
    for k in ("seq1", "qual1", "seq2", "qual2", "len1", "len2"):
        mine = devt[k][:npar].contiguous()
        if world > 1:                                   # NCCL has no int16: gather the bytes
            mb = mine.view(torch.uint8)
            parts = [torch.empty_like(mb) for _ in range(world)]
            dist.all_gather(parts, mb)
            gathered[k] = torch.cat([q.view(mine.dtype) for q in parts]).cpu().numpy()
        else:
            gathered[k] = mine.cpu().numpy()
    res_par = [out_dev[m][:npar].cpu().numpy().view(abi.RESULT_DTYPE).copy() for m in range(2)]
    stats_parity = None
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as orc
        dd = {k: (v.view(np.uint16) if k.startswith("len") else v) for k, v in gathered.items()}
        dd["n"] = npar * world; dd["stride"] = stride
        o1, o2, ost, oerr = orc.filter_pe(params, dd)
        got = table.cpu().numpy().view(np.uint64)
        stats_parity = bool(oerr == 0 and np.array_equal(got, ost) and np.array_equal(res_par[0], o1[:npar]) and np.array_equal(res_par[1], o2[:npar]))
        if not stats_parity:
            raise SystemExit("stats_parity FAILED: the all-reduced statistics table differs from the oracle's")
    check(lib.snk_engine_stats_reset(h))

    # ---- pinned host->device copy bandwidth of this box with all ranks copying at once (ceiling of e2e_soa)
    probe = torch.empty(1 << 29, dtype=torch.uint8).pin_memory()
    probe_dev = torch.empty(1 << 29, dtype=torch.uint8, device=dev)
    probe_dev.copy_(probe, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        probe_dev.copy_(probe, non_blocking=True)
    torch.cuda.synchronize()
    h2d_s = max_over_ranks(time.perf_counter() - t0)
    h2d_gbs = world * 4 * probe.numel() / h2d_s / 1e9
    del probe, probe_dev

    # ---- e2e_soa leg: pinned host buffers through the async host API, 3 lanes, sub-batches
    nl = lib.snk_engine_lanes(h)
    sub = args.sub_pairs
    chunks = [(a, min(pairs, a + sub)) for a in range(0, pairs, sub)]

    def e2e_step(i):
        for ci, (a, b) in enumerate(chunks):
            lane = ci % nl
            b1 = abi.Batch(pin["seq1"][a:b].data_ptr(), pin["qual1"][a:b].data_ptr(), pin["len1"][a:b].data_ptr(), b - a, stride)
            b2 = abi.Batch(pin["seq2"][a:b].data_ptr(), pin["qual2"][a:b].data_ptr(), pin["len2"][a:b].data_ptr(), b - a, stride)
            check(lib.snk_filter_pe_async(h, lane, C.byref(b1), C.byref(b2), out_pin[0][a:b].data_ptr(), out_pin[1][a:b].data_ptr(),
                                          C.c_uint64(i * pairs + a)))
        for lane in range(nl):
            check(lib.snk_engine_lane_sync(h, lane))

    e2e_warm = max(1, min(args.warmup, 2))
    for i in range(e2e_warm):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    soa_value = reads_per_step * args.steps / e2e_s / 1e6
    h2d = sum(pin[k].numel() * pin[k].element_size() for k in pin)
    d2h = sum(t.numel() * t.element_size() for t in out_pin)
    step_resident(args.warmup + args.steps)       # same data, resident: the two entry points must agree
    torch.cuda.synchronize()
    same = bool(torch.equal(out_dev[0].cpu(), out_pin[0]) and torch.equal(out_dev[1].cpu(), out_pin[1]))
    kept = int((out_pin[0].numpy().view(abi.RESULT_DTYPE)["category"] == 0).sum())

    # ---- e2e_text leg: raw FASTQ text in pinned host memory -> clean FASTQ text in pinned host memory
    # (line index, row packing, filter, record formatting all on the device)
    text_info = None
    if not args.no_text:
        work = tempfile.mkdtemp(prefix="snktxt_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        try:
            tsub = min(sub, host["n"])
            tin = []
            for m in (1, 2):
                synth.write_fastq_fixed(f"{work}/r{m}.fq", host[f"seq{m}"][:tsub], host[f"qual{m}"][:tsub], L, m)
                tin.append(torch.from_numpy(np.fromfile(f"{work}/r{m}.fq", dtype=np.uint8)).pin_memory())
        finally:
            shutil.rmtree(work, ignore_errors=True)
        tout = [[torch.empty(tin[m].numel() + 64, dtype=torch.uint8).pin_memory() for m in range(2)] for _ in range(nl)]
        toff = [[torch.empty(tsub + 1, dtype=torch.int32).pin_memory() for m in range(2)] for _ in range(nl)]
        fmt = abi.TextFormat(strip=1)
        meta = abi.TextMeta()
        nchunks = max(1, pairs // tsub)
        moved = [0, 0]

        def finish(lane):
            check(lib.snk_text_meta_sync(h, lane, C.byref(meta)))
            if meta.flags:
                raise SystemExit(f"text path flags {meta.flags}")
            check(lib.snk_text_fetch_async(h, lane, tout[lane][0].data_ptr(), tout[lane][1].data_ptr(), toff[lane][0].data_ptr(),
                                           toff[lane][1].data_ptr(), None, None))
            check(lib.snk_engine_lane_sync(h, lane))
            moved[1] = meta.out_bytes[0] + meta.out_bytes[1] + 2 * 4 * (tsub + 1)

        def text_step(i):
            fifo = []
            for ci in range(nchunks):
                lane = ci % nl
                if len(fifo) == nl:
                    finish(fifo.pop(0))
                check(lib.snk_filter_pe_text_async(h, lane, tin[0].data_ptr(), tin[0].numel(), tin[1].data_ptr(), tin[1].numel(),
                                                   tsub, stride, C.byref(fmt), C.c_uint64((i * nchunks + ci) * tsub)))
                fifo.append(lane)
            while fifo:
                finish(fifo.pop(0))

        tsteps = max(1, min(args.steps, 5))
        for i in range(e2e_warm):
            text_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(tsteps):
            text_step(i)
        torch.cuda.synchronize()
        text_s = max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            dist.barrier()
        text_info = {"value": 2 * tsub * nchunks * world * tsteps / text_s / 1e6, "unit": UNIT,
                     "h2d_bytes_per_step": int(nchunks * (tin[0].numel() + tin[1].numel())), "d2h_bytes_per_step": int(nchunks * moved[1]),
                     "sub_batch_pairs": tsub, "lanes": nl, "steps": tsteps,
                     "what": "raw FASTQ text (pinned host) -> clean FASTQ text (pinned host) through snk_filter_pe_text_async"}

    flags = C.c_uint32(0); bad = C.c_uint64(0)
    check(lib.snk_engine_error_flags(h, C.byref(flags), C.byref(bad)))
    if flags.value:
        raise SystemExit(f"engine raised error flags {flags.value} at read {bad.value}")
    lib.snk_engine_destroy(h)
    del devt, out_dev, pin, out_pin, table
    torch.cuda.empty_cache()

    soa_info = {"value": soa_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "sub_batch_pairs": sub, "lanes": nl, "results_match_resident": same,
                "h2d_probe": {"aggregate_gbs": h2d_gbs, "what": "pinned cudaMemcpyAsync host->device, all ranks at once, 2 GiB per rank",
                              "ceiling_mreads_s": h2d_gbs * 1e9 / (h2d / (2.0 * pairs)) / 1e6,
                              "frac_of_ceiling": soa_value / (h2d_gbs * 1e9 / (h2d / (2.0 * pairs)) / 1e6)}}
    # ---- e2e leg, file to file: the drop-in CLI, one process per GPU, on the files the reference arm reads
    fpairs = args.file_pairs or pairs
    file_bytes = 2 * fpairs * (2 * L + 25)
    work = None
    if rank == 0 and not args.no_file:
        work = work_dir(file_bytes * (1 + world) + (2 << 30))
        in_bytes = write_workload_files(work, fpairs)
    if world > 1:
        box = [work]
        dist.broadcast_object_list(box, src=0)
        work = box[0]
    env = dict(os.environ)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    env["CUDA_VISIBLE_DEVICES"] = visible.split(",")[local_rank] if visible else str(local_rank)
    env["SNK_GPUS"] = "1"
    out_mine = f"{work}/out_r{rank}"
    cores = os.cpu_count() or 1
    threads = max(2, cores // world)
    cli_cmd = [CLI, "filter"] + file_args(work, out_mine) + ["-T", str(min(threads, 16))] + CFG2_FLAGS
    file_info = None
    cpu_baseline = None
    try:
        if args.no_file:
            raise StopIteration
        fwarm = max(1, min(args.warmup, 2))
        for _ in range(fwarm):
            run_timed(cli_cmd, env)
        barrier()
        t0 = time.perf_counter()
        runs = [run_timed(cli_cmd, env) for _ in range(args.steps)]
        file_s = max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            dist.barrier()
        file_value = 2.0 * fpairs * world * args.steps / file_s / 1e6
        out_bytes = os.path.getsize(f"{out_mine}/c1.fq") + os.path.getsize(f"{out_mine}/c2.fq")
        stage_line = ""
        try:
            for ln in open(f"{out_mine}/log") if os.path.exists(f"{out_mine}/log") else []:
                if "stage seconds" in ln:
                    stage_line = ln.strip()
        except Exception:
            pass
        if rank == 0:
            file_info = {"value": file_value, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
                         "what": "soapnuke_b200/bin/SOAPnuke filter, plain FASTQ files on /dev/shm -> clean FASTQ + reports, one process per GPU, "
                                 "wall clock per run (process start and CUDA context included)",
                         "pairs_per_run": fpairs, "wall_s_per_run": float(np.mean([r["wall"] for r in runs])),
                         "cpu_s_per_run": float(np.mean([r["cpu"] for r in runs])), "host_threads": min(threads, 16),
                         "outputs_match_reference": None, "cli_stage_log": stage_line}
            if world == 1 and not args.no_cpu_baseline:
                try:
                    if have_reference():
                        rr = reference_runs(work, f"{work}/out_ref", 1, cores)
                        v = 2.0 * fpairs / rr[0]["wall"] / 1e6
                        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": describe_reference(fpairs, cores, rr),
                                        "wall_s_per_run": rr[0]["wall"], "cpu_s_per_run": rr[0]["cpu"], "poll_quanta_per_run": rr[0]["wall"] / 5.0}
                        same_files = all(filecmp.cmp(f"{work}/out_ref/c{m}.fq", f"{out_mine}/c{m}.fq", shallow=False) for m in (1, 2))
                        import glob
                        reports = sorted(glob.glob(f"{work}/out_ref/*.txt"))
                        same_reports = len(reports) == 10 and all(filecmp.cmp(f, f"{out_mine}/{os.path.basename(f)}", shallow=False) for f in reports)
                        file_info["outputs_match_reference"] = bool(same_files and same_reports)
                        if not (same_files and same_reports):
                            raise SystemExit("e2e FAILED: CLI outputs differ from the reference binary's on the bench files")
                    else:
                        v, n = time_oracle_port(fpairs, 1, 0)
                        cpu_baseline = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": f"oracle C restatement, 1 thread, {n} synthetic PE150 pairs already parsed in memory"}
                except SystemExit:
                    raise
                except Exception as ex:           # keep the GPU line even if the CPU leg fails
                    cpu_baseline = {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}
    except StopIteration:
        file_info = dict(soa_info, what="pinned SoA leg (--no-file)")
    finally:
        if world > 1:
            dist.barrier()
        if rank == 0 and work:
            shutil.rmtree(work, ignore_errors=True)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = 2 * pairs * (2 * L + 8)                       # per launch, per GPU
        launch_ms = kernel_ms / args.steps
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        tr = recorded_traffic()
        traffic = None
        if tr and tr.get("bytes_per_read"):
            traffic = tr["bytes_per_read"] * 2 * pairs
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(pairs),
            "e2e": file_info,
            "e2e_soa": soa_info,
            "e2e_text": text_info,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_read": 2 * L + 8,
                         "kernel": "snkcore::filter_kernel (PE150 instantiation)", "launch_ms": launch_ms},
            "collective": {"what": "NCCL all-reduce of the statistics table (SUM + tiny MAX), once per run, inside the timed region",
                           "us": collective_us, "bytes": int(params.n_slots * abi.SLOT_WORDS * 8), "ranks": world},
            "stats_parity": stats_parity,
            "stats_parity_what": f"{npar} pairs per rank through the engine, all-reduced table + rank 0's records == oracle over all ranks' slices",
            "engine": {"slots": params.n_slots, "slot_block": params.slot_block, "stride": stride},
            "kept_pairs_last_step": kept,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 22, help="read pairs per step per GPU")
    ap.add_argument("--sub-pairs", type=int, default=1 << 19, help="pairs per host sub-batch in the e2e_soa / e2e_text legs")
    ap.add_argument("--file-pairs", type=int, default=0, help="pairs in the FASTQ files of the e2e leg (default: --pairs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text", action="store_true", help="skip the FASTQ-text end-to-end leg")
    ap.add_argument("--no-file", action="store_true", help="skip the file-to-file CLI leg (profiling runs): `e2e` is then the pinned-SoA leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
