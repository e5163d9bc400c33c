"""Quick per-config kernel timing (resident batches) for the other BASELINE shapes."""
import sys, ctypes as C, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from soapnuke_b200 import abi, synth
from helpers import CFG2_KW, A1, A2
lib = abi.load_engine()
def run(name, pe, L, n, pkw, gkw):
    base = synth.gen_pairs(1 << 17, L=L, se=not pe, **gkw)
    reps = n // base["n"]
    dev = torch.device("cuda:0")
    t = {}
    for k, v in base.items():
        if isinstance(v, np.ndarray):
            a = np.tile(v, (reps, 1)) if v.ndim == 2 else np.tile(v, reps)
            t[k] = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(dev)
    p = abi.make_params(is_pe=pe, threads=8, nprocs=1 << 20, **pkw)
    h = C.c_void_p(); assert lib.snk_engine_create(C.byref(p), 0, C.byref(h)) == 0, lib.snk_last_error()
    out1 = torch.empty(n, dtype=torch.int64, device=dev); out2 = torch.empty(n, dtype=torch.int64, device=dev)
    b1 = abi.Batch(t["seq1"].data_ptr(), t["qual1"].data_ptr(), t["len1"].data_ptr(), n, base["stride"])
    if pe: b2 = abi.Batch(t["seq2"].data_ptr(), t["qual2"].data_ptr(), t["len2"].data_ptr(), n, base["stride"])
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def step(i):
        if pe: rc = lib.snk_filter_pe_device(h, C.byref(b1), C.byref(b2), out1.data_ptr(), out2.data_ptr(), i * n, s)
        else: rc = lib.snk_filter_se_device(h, C.byref(b1), out1.data_ptr(), i * n, s)
        assert rc == 0, lib.snk_last_error()
    for i in range(3): step(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for i in range(5): step(3 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    reads = n * (2 if pe else 1)
    print(f"{name:28s} L={L:4d} reads/launch={reads:9d}  {ms:8.3f} ms  {reads/ms/1e3:9.1f} Mreads/s  {reads*(2*L+8)/ms/1e6:8.1f} GB/s algorithmic")
    lib.snk_engine_destroy(h)
run("cfg1 SE150 default", False, 150, 1 << 22, dict(), dict(seed=1001))
run("cfg2 PE150 all filters", True, 150, 1 << 21, CFG2_KW, dict(seed=1002))
run("cfg2 PE150 discard-mode", True, 150, 1 << 21, dict(adapter1=A1, adapter2=A2), dict(seed=1002))
run("cfg4 SE50 adapter trim", False, 50, 1 << 23, dict(adapter1=synth.SRNA_ADAPTER3.decode(), ada_trim=True, min_read_length=15), dict(seed=1004, adapter1=synth.SRNA_ADAPTER3, insert_range=(15, 35)))
run("cfg5 PE250 polyG", True, 250, 1 << 20, dict(adapter1=A1, adapter2=A2, ada_trim=True, polyG_tail=10), dict(seed=1005, polyg_frac=0.3))
run("PE100 all filters", True, 100, 1 << 21, CFG2_KW, dict(seed=1006))
def run_srna(name, L, n, pkw):
    """filtersRNA module (BASELINE config 4 served natively)"""
    global synth
    base = synth.gen_srna(1 << 17, L=L, seed=1004)
    reps = n // base["n"]
    dev = torch.device("cuda:0")
    t = {}
    for k, v in base.items():
        if isinstance(v, np.ndarray):
            a = np.tile(v, (reps, 1)) if v.ndim == 2 else np.tile(v, reps)
            t[k] = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(dev)
    p = abi.make_params(is_pe=False, srna=True, adapter1=synth.SRNA_ADAPTER5, adapter2=synth.SRNA_ADAPTER3, threads=8, nprocs=1 << 20,
                        min_read_length=18, max_read_length=49, **pkw)
    h = C.c_void_p(); assert lib.snk_engine_create(C.byref(p), 0, C.byref(h)) == 0, lib.snk_last_error()
    out1 = torch.empty(n, dtype=torch.int64, device=dev)
    b1 = abi.Batch(t["seq1"].data_ptr(), t["qual1"].data_ptr(), t["len1"].data_ptr(), n, base["stride"])
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(3): assert lib.snk_filter_se_device(h, C.byref(b1), out1.data_ptr(), i * n, s) == 0
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): assert lib.snk_filter_se_device(h, C.byref(b1), out1.data_ptr(), (3 + i) * n, s) == 0
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:28s} L={L:4d} reads/launch={n:9d}  {ms:8.3f} ms  {n/ms/1e3:9.1f} Mreads/s  {n*(2*L+8)/ms/1e6:8.1f} GB/s algorithmic")
    lib.snk_engine_destroy(h)
run_srna("cfg4 SE50 filtersRNA trim", 50, 1 << 23, dict(ada_trim=True))
