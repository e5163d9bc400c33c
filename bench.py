#!/usr/bin/env python
"""bench.py — M reads/s of the FASTQ filter hot path on synthetic PE150 (BASELINE config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]

One "step" = one pass of the hot path (adapter match, trim, predicates, discard cascade, raw+clean
per-position histograms) over one batch of P read pairs per GPU with BASELINE config-2 flags
(`-f A1 -r A2 -J -l 5 -q 0.5 -n 0.05 -m 15 -p 0.7 -X 50 -g 10 -y 20,30 -x 20,10`).

 value   whole-job M reads/s with the batch already resident in HBM (one kernel launch per step,
         timed with CUDA events on the launching stream, max over ranks).
 e2e     same metric through the host-buffer C-ABI entry points (snk_filter_pe_async on pinned host
         memory): host->device copies of every step's batch and device->host copies of the per-read
         result records are inside the timed region.
 roofline  algorithmic bytes (2L+8 per read) / average kernel duration vs the
         measured HBM copy bandwidth of MEASURED_PEAKS.json.
 cpu_baseline  the unmodified reference binary (oracle/_ref/SOAPnuke, `filter -T <ncores>`) timed on
         this box's cores on a bounded sample of the same workload (rank 0, N=1 only).

`--impl reference` times only that CPU reference and prints the same JSON shape.
"""
import argparse
import ctypes as C
import json
import os
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


# --------------------------------------------------------------------------- reference arm
def time_reference(sample_pairs, steps, warmup, seed=1002):
    """`SOAPnuke filter` (unmodified reference, all host cores) on a bounded sample; returns
    (Mreads/s, cores, seconds per step list, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    cores = os.cpu_count() or 1
    work = tempfile.mkdtemp(prefix="snkref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        unique = min(sample_pairs, 1 << 20)
        d = synth.gen_pairs(unique, L=L, seed=seed)
        for m in (1, 2):                      # the sample = `unique` generated pairs tiled to sample_pairs, ids running on
            with open(f"{work}/r{m}.fq", "wb") as f:
                for k in range(0, sample_pairs, unique):
                    n = min(unique, sample_pairs - k)
                    synth.write_fastq_fixed(f"{work}/part.fq", d[f"seq{m}"][:n], d[f"qual{m}"][:n], L, m, first=k)
                    with open(f"{work}/part.fq", "rb") as g:
                        shutil.copyfileobj(g, f, 1 << 24)
                    os.unlink(f"{work}/part.fq")
        if orc.have_reference():
            kind = "reference"
            times = []
            for s in range(warmup + steps):
                shutil.rmtree(f"{work}/out", ignore_errors=True)
                t0 = time.perf_counter()
                r = orc.run_reference(["-1", f"{work}/r1.fq", "-2", f"{work}/r2.fq", "-C", "c1.fq", "-D", "c2.fq",
                                       "-o", f"{work}/out", "-T", str(cores)] + CFG2_FLAGS, timeout=3600)
                dt = time.perf_counter() - t0
                if r.returncode != 0:
                    raise RuntimeError("reference run failed: " + r.stderr.decode()[-300:])
                if s >= warmup:
                    times.append(dt)
            sample = (f"SOAPnuke 2.1.9 filter -T {cores} on {sample_pairs} synthetic PE150 pairs, plain FASTQ in/out on /dev/shm, "
                      "whole-program wall time (includes FASTQ parse/format and its 5 s concat poll quantum)")
        else:
            kind = "port"
            cores = 1
            p = abi.make_params(is_pe=True, **CFG2_KW)
            d = synth.gen_pairs(sample_pairs, L=L, seed=seed) if unique < sample_pairs else d
            times = []
            for s in range(warmup + steps):
                t0 = time.perf_counter()
                orc.filter_pe(p, d)
                dt = time.perf_counter() - t0
                if s >= warmup:
                    times.append(dt)
            sample = f"oracle C restatement, 1 thread, {sample_pairs} synthetic PE150 pairs already parsed in memory"
        total = sum(times)
        return 2.0 * sample_pairs * len(times) / total / 1e6, cores, times, kind, sample
    finally:
        shutil.rmtree(work, ignore_errors=True)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, cores, times, kind, sample = time_reference(args.ref_pairs, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: PE 2x150bp, adapter trim + all quality filters (config-2 flags)",
                   "pairs_per_step": args.ref_pairs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from soapnuke_b200 import build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the filter engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    build.build_engine()
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
    # pinned host copies (e2e leg) and device-resident copies (value leg)
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

    def step_resident(i):
        check(lib.snk_filter_pe_device(h, C.byref(db1), C.byref(db2), out_dev[0].data_ptr(), out_dev[1].data_ptr(),
                                       C.c_uint64(i * pairs), sptr))

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

    # ---- value leg: device-resident batches, CUDA events on the launching stream
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    launches0 = lib.snk_engine_launch_count(h)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        step_resident(args.warmup + i)
    ev1.record(stream)
    # the single collective of the path: the final statistics table (sum; LAST_KEY words are max-reduced on host)
    if world > 1:
        from soapnuke_b200 import dist as snkdist
        st = torch.empty(params.n_slots * abi.SLOT_WORDS, dtype=torch.int64, device=dev)
        check(lib.snk_engine_stats_to_device(h, st.data_ptr(), sptr))
        snkdist.allreduce_stats(st, params.n_slots)
    barrier()
    kernel_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.snk_engine_launch_count(h) - launches0
    kernel_ms = max_over_ranks(kernel_ms)
    reads_per_step = 2 * pairs * world
    value = reads_per_step * args.steps / (kernel_ms * 1e-3) / 1e6

    # ---- e2e leg: pinned host buffers through the async host API, 3 lanes, sub-batches
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
    e2e_value = reads_per_step * args.steps / e2e_s / 1e6
    h2d = sum(pin[k].numel() * pin[k].element_size() for k in pin)
    d2h = sum(t.numel() * t.element_size() for t in out_pin)

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

        for i in range(e2e_warm):
            text_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            text_step(i)
        torch.cuda.synchronize()
        text_s = max_over_ranks(time.perf_counter() - t0)
        if world > 1:
            dist.barrier()
        text_info = {"value": 2 * tsub * nchunks * world * args.steps / text_s / 1e6, "unit": UNIT,
                     "h2d_bytes_per_step": int(nchunks * (tin[0].numel() + tin[1].numel())), "d2h_bytes_per_step": int(nchunks * moved[1]),
                     "sub_batch_pairs": tsub, "lanes": nl,
                     "what": "raw FASTQ text (pinned host) -> clean FASTQ text (pinned host) through snk_filter_pe_text_async"}

    flags = C.c_uint32(0); bad = C.c_uint64(0)
    check(lib.snk_engine_error_flags(h, C.byref(flags), C.byref(bad)))
    if flags.value:
        raise SystemExit(f"engine raised error flags {flags.value} at read {bad.value}")
    # sanity: the resident results of the last step equal the e2e results of the same data
    same = bool(torch.equal(out_dev[0].cpu(), out_pin[0]) and torch.equal(out_dev[1].cpu(), out_pin[1]))
    kept = int((out_pin[0].numpy().view(abi.RESULT_DTYPE)["category"] == 0).sum())

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
            "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: PE 2x150bp, adapter trim + all quality filters (config-2 flags), 1xB200 shape per GPU",
                       "pairs_per_step_per_gpu": pairs, "read_len": L, "stride": stride,
                       "l2": "inputs larger than L2 (%.2f GB per step per GPU)" % (4 * pairs * stride / 1e9),
                       "slots": params.n_slots, "slot_block": params.slot_block},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "sub_batch_pairs": sub, "lanes": nl, "results_match_resident": same},
            "e2e_text": text_info,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_read": 2 * L + 8,
                         "kernel": "snkcore::filter_kernel<10,2>", "launch_ms": launch_ms},
            "kept_pairs_last_step": kept,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                v, cores, times, kind, sample = time_reference(args.ref_pairs, 1, 0)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
            except Exception as ex:           # keep the GPU line even if the CPU leg fails
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    lib.snk_engine_destroy(h)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1 << 22, help="read pairs per step per GPU")
    ap.add_argument("--sub-pairs", type=int, default=1 << 19, help="pairs per host sub-batch in the e2e leg")
    ap.add_argument("--ref-pairs", type=int, default=4000000,
                    help="pairs in the CPU reference sample (8 M reads: the reference's 5 s concat poll quantum stays a small part)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text", action="store_true", help="skip the FASTQ-text end-to-end leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
