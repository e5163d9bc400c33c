#!/usr/bin/env python
"""Extracts the judged numbers from an `ncu --set full` capture (.ncu-rep, kept in gpurun_out/) into
small committed text files: key raw metrics, per-function instruction/stall attribution (needs the
in-tree .so for line info), and profiles/traffic.json (DRAM bytes per read, used by bench.py).

    python profiles/summarize.py gpurun_out/prof_r1_v3.ncu-rep profiles/r1_v3 --reads 2097152
"""
import argparse
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_prefix")
    ap.add_argument("--reads", type=float, required=True, help="reads processed by the profiled launch")
    ap.add_argument("--kernel", default="_ZN7snkcore13filter_kernelILi10ELi2E")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    out = []
    got = {}
    for i, h in enumerate(hdr):
        if h in KEYS or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            out.append(f"{h}\t{units[i]}\t{vals[i]}")
            got[h] = (units[i], vals[i])
    kname = [vals[i] for i, h in enumerate(hdr) if h == "Kernel Name"]
    text = [f"# {os.path.basename(a.rep)}  kernel: {kname[0] if kname else '?'}  reads in launch: {int(a.reads)}"] + out

    def to_bytes(u, v):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if "dram__bytes_read.sum" in got:
        tr = to_bytes(*got["dram__bytes_read.sum"]) + to_bytes(*got["dram__bytes_write.sum"])
        text.append(f"dram bytes per read\t{tr / a.reads:.1f}\t(algorithmic 2L+8 = 308 for PE150)")
        json.dump({"bytes_per_read": tr / a.reads, "source": os.path.basename(a.rep), "reads": a.reads},
                  open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))
    # per-instruction page -> attribute to functions through nvdisasm line info
    src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h2, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(h2)}
    tmp = "/tmp/snk_prof_cub"
    subprocess.run(f"rm -rf {tmp}; mkdir -p {tmp}; cd {tmp}; cuobjdump -xelf all {ROOT}/soapnuke_b200/lib/libsnk_engine.so >/dev/null 2>&1", shell=True)
    sass = subprocess.run(f"nvdisasm -g {tmp}/engine.sm_100a.cubin", shell=True, capture_output=True, text=True).stdout.splitlines()
    st = [i for i, l in enumerate(sass) if ".section" in l and ".text." + a.kernel in l]
    if st:
        en = [i for i, l in enumerate(sass) if ".section" in l and i > st[0]][0]
        cur, seq = None, []
        for ln in sass[st[0]:en]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+", ln):
                seq.append(cur)

        def fmap(path):
            fm, cur = {}, "?"
            for i, l in enumerate(open(path).read().splitlines(), 1):
                m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:SNK_HD_NOINLINE|SNK_HD_MEMBER|SNK_HD|__device__ __forceinline__|__device__|__global__|inline|__host__ __device__ inline)\s+[\w:<>\s\*&]*?\b(\w+)\s*\(", l)
                if m and not l.strip().startswith("//"):
                    cur = m.group(1)
                if l.startswith("filter_kernel("):          # the kernel's declarator sits on its own line behind __launch_bounds__
                    cur = "filter_kernel"
                fm[i] = cur
            return fm
        fmc = fmap(os.path.join(ROOT, "soapnuke_b200/csrc/filter_core.cuh"))
        ksrc = open(os.path.join(ROOT, "soapnuke_b200/csrc/filter_kernel.cuh")).read().splitlines()

        fmk = fmap(os.path.join(ROOT, "soapnuke_b200/csrc/filter_kernel.cuh"))

        def kphase(l):
            # inside filter_kernel: the nearest "// ---- phase" marker above the line; elsewhere: the enclosing function
            fn = fmk.get(l, "?")
            if fn != "filter_kernel":
                return fn
            for i in range(l - 1, 0, -1):
                if "// ---- " in ksrc[i - 1]:
                    return ksrc[i - 1].strip()[8:30]
                if "filter_kernel(" in ksrc[i - 1]:
                    break
            return "filter_kernel prologue / tile loop"
        g_inst, g_samp, g_thr = collections.Counter(), collections.Counter(), collections.Counter()
        if len(seq) == len(data):
            for fl, r in zip(seq, data):
                if fl is None:
                    g = "none"
                elif fl[0] == "filter_core.cuh":
                    g = "core:" + fmc.get(fl[1], "?")
                elif fl[0] == "filter_kernel.cuh":
                    g = "kernel:" + kphase(fl[1])
                else:
                    g = fl[0]
                g_inst[g] += int(r[ix["Instructions Executed"]])
                g_samp[g] += int(r[ix["# Samples"]])
                g_thr[g] += int(r[ix["Thread Instructions Executed"]])
            ti, ts = sum(g_inst.values()), sum(g_samp.values())
            text.append(f"\n# warp instructions per read: {ti / a.reads:.1f}   (function: warp-inst/read, stall-sample share, avg active threads)")
            for g, v in g_inst.most_common(24):
                text.append(f"{g:42s} {v / a.reads:7.1f}  {100 * g_samp[g] / ts:5.1f}%  {g_thr[g] / max(v, 1):5.1f}")
        else:
            text.append(f"# line attribution skipped: SASS length {len(seq)} != profile {len(data)} (library rebuilt since the capture?)")
    open(a.out_prefix + "_summary.txt", "w").write("\n".join(text) + "\n")
    print("\n".join(text))


if __name__ == "__main__":
    main()
