"""GPU tier, N > 1: batches shard round-robin over the GPUs, no data-path collective, the per-GPU
statistics tables are summed once at the end (LAST_KEY words max-reduced). Skipped on boxes with fewer GPUs.

  * several engines in ONE process (what the CLI does with SNK_GPUS=n): batches k -> engine k % n, tables summed on
    the host, per-read records + the summed table bit-exact against the oracle run over the whole input
  * the drop-in CLI with SNK_GPUS=2 / 4 against the unmodified reference binary: clean FASTQ + all reports identical
    (plain and .gz, PE multi-cycle with the deferred-batch emission order, small batches so that every GPU gets many)
  * one process per GPU under torchrun + NCCL (what bench.py does): the all-reduced table equals the oracle's
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as orc
from helpers import A1, A2, CFG2_FLAGS, CFG2_KW, ROOT, Engine, abi, assert_same, oracle_run, synth

pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def need(n):
    return pytest.mark.skipif(gpu_count() < n, reason=f"needs {n} GPUs")


def sum_tables(parts, n_slots):
    """Host-side merge the CLI performs (process.cpp): counters add, LAST_KEY words take the max."""
    from soapnuke_b200.dist import last_key_positions
    idx = np.array(last_key_positions(n_slots))
    keys = np.max(np.stack([p[idx] for p in parts]), axis=0)
    tot = np.zeros_like(parts[0])
    for p in parts:
        q = p.copy()
        q[idx] = 0
        tot += q
    tot[idx] = keys
    return tot


@pytest.mark.parametrize("ngpu", [pytest.param(2, marks=need(2)), pytest.param(4, marks=need(4)), pytest.param(8, marks=need(8))])
def test_engines_on_several_devices_in_one_process(engine_lib, ngpu):
    """Regression for the per-device launch cache (engine.cu launch_one): every engine must raise its own kernel's
    dynamic shared-memory limit on ITS device. Round-robin batches, summed tables == oracle over the whole input."""
    n, batch = 90000, 7000
    d = synth.gen_pairs(n, L=150, seed=4242)
    p = abi.make_params(is_pe=True, threads=4, patch_size=1500, **CFG2_KW)
    o1, o2, ost, oerr = oracle_run(p, d)
    assert oerr == 0
    engines = [Engine(engine_lib, p, device=g) for g in range(ngpu)]
    try:
        r1 = np.zeros(n, dtype=abi.RESULT_DTYPE)
        r2 = np.zeros(n, dtype=abi.RESULT_DTYPE)
        for k, a in enumerate(range(0, n, batch)):
            b = min(n, a + batch)
            sub = {key: (np.ascontiguousarray(v[a:b]) if isinstance(v, np.ndarray) else v) for key, v in d.items()}
            x1, x2 = engines[k % ngpu].filter_host(sub, first=a)
            r1[a:b] = x1
            r2[a:b] = x2
        parts = [e.stats() for e in engines]
        for e in engines:
            assert e.error_flags()[0] == 0
    finally:
        for e in engines:
            e.close()
    assert all(pt.any() for pt in parts), "every engine must have processed batches"
    assert_same((r1, r2, sum_tables(parts, p.n_slots)), (o1, o2, ost), f"{ngpu} engines in one process")


MULTI_CASES = [
    dict(name="multi_pe_cfg2_plain_T4_multicycle", pe=True, n=60000, L=150, T=4, flags=CFG2_FLAGS, patch=25, env={"SNK_BATCH_READS": "4096"}),
    dict(name="multi_pe_cfg2_gz_T3", pe=True, n=40000, L=150, T=3, flags=CFG2_FLAGS, patch=40, gz_in=True, gz_out=True, env={"SNK_BATCH_READS": "3000"}),
    dict(name="multi_se_adapter_T4", pe=False, n=50000, L=100, T=4, flags=["-f", A1, "-J", "-g", "10"], patch=11, env={"SNK_BATCH_READS": "5000"}),
    dict(name="multi_pe_trimfiles_T2", pe=True, n=30000, L=150, T=2, flags=CFG2_FLAGS, patch=40, trim=True, env={"SNK_BATCH_READS": "3000"}),
]


import test_cli_gpu as tc  # noqa: E402  (the reference sides of all CLI cases run in one background pool)
if gpu_count() >= 2:
    tc.reg([dict(c) for c in MULTI_CASES])


@pytest.mark.skipif(not orc.have_reference(), reason="reference binary not available")
@pytest.mark.parametrize("ngpu", [pytest.param(2, marks=need(2)), pytest.param(4, marks=need(4))])
@pytest.mark.parametrize("case", MULTI_CASES, ids=lambda c: c["name"])
def test_cli_sharded_over_gpus_matches_reference_binary(tmp_path, case, ngpu):
    from soapnuke_b200 import build
    build.build_all()
    c = dict(case)
    env = dict(c.pop("env", {}))
    env["SNK_GPUS"] = str(ngpu)
    tc.run_both(tc.CLI, tmp_path, env=env, out_name=f"mine_g{ngpu}", **c)


TORCHRUN_WORKER = r"""
import ctypes as C, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SNK_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SNK_ROOT"], "oracle")); sys.path.insert(0, os.path.join(os.environ["SNK_ROOT"], "tests"))
from soapnuke_b200 import abi, synth, dist as snkdist
from helpers import CFG2_KW, Engine
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = abi.load_engine()
n, batch = 64000, 5000
d = synth.gen_pairs(n, L=150, seed=99)
p = abi.make_params(is_pe=True, threads=3, patch_size=2000, **CFG2_KW)
nb = (n + batch - 1) // batch
with Engine(lib, p, device=lr) as e:
    for k in snkdist.shard_batches(nb, rank, world):
        a, b = k * batch, min(n, (k + 1) * batch)
        sub = {key: (np.ascontiguousarray(v[a:b]) if isinstance(v, np.ndarray) else v) for key, v in d.items()}
        e.filter_host(sub, first=a)
    st = torch.empty(p.n_slots * abi.SLOT_WORDS, dtype=torch.int64, device="cuda")
    e.check(lib.snk_engine_stats_to_device(e.h, st.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    snkdist.allreduce_stats(st, p.n_slots)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.environ["SNK_OUT"], st.cpu().numpy().view(np.uint64))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("ngpu", [pytest.param(2, marks=need(2)), pytest.param(4, marks=need(4))])
def test_torchrun_nccl_allreduced_table_equals_oracle(tmp_path, ngpu):
    script = tmp_path / "worker.py"
    script.write_text(TORCHRUN_WORKER)
    env = dict(os.environ, SNK_ROOT=ROOT, SNK_OUT=str(tmp_path / "reduced.npy"))
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ngpu}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    reduced = np.load(tmp_path / "reduced.npy")
    d = synth.gen_pairs(64000, L=150, seed=99)
    p = abi.make_params(is_pe=True, threads=3, patch_size=2000, **CFG2_KW)
    _, _, whole, _ = oracle_run(p, d)
    assert np.array_equal(reduced, whole)
