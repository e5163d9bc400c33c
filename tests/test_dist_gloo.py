"""CPU tier of the N>1 path: batches shard round-robin across ranks (one process per GPU in
production, NCCL); the only exchange is the final statistics table. Here world_size 2 over gloo:
each rank filters its shard (the oracle stands in for the engine on CPU), the tables are reduced with
soapnuke_b200.dist.allreduce_stats (SUM for counters, MAX for the LAST_KEY words) and must equal the
single-process table bit for bit - including the report files written from it."""
import filecmp
import glob
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import CFG2_KW, ROOT, abi, synth


def _worker(rank, world, port, tmp, n, batch):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    from soapnuke_b200 import dist as snkdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = synth.gen_pairs(n, L=100, seed=77)
    p = abi.make_params(is_pe=True, threads=3, patch_size=30, **CFG2_KW)
    st = orc.new_stats(p)
    nb = (n + batch - 1) // batch
    for k in snkdist.shard_batches(nb, rank, world):
        a, b = k * batch, min(n, (k + 1) * batch)
        sub = {key: (np.ascontiguousarray(v[a:b]) if isinstance(v, np.ndarray) else v) for key, v in d.items()}
        orc.filter_pe(p, sub, stats=st, first_index=a)
    t = torch.from_numpy(st.view(np.int64).copy())
    snkdist.allreduce_stats(t, p.n_slots)
    if rank == 0:
        np.save(os.path.join(tmp, "reduced.npy"), t.numpy().view(np.uint64))
    dist.destroy_process_group()


def test_sharded_tables_allreduce_to_the_single_process_table(tmp_path, engine_lib):
    import ctypes as C
    import oracle_py as orc
    n, batch, world = 12000, 1700, 2
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path), n, batch), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    d = synth.gen_pairs(n, L=100, seed=77)
    p = abi.make_params(is_pe=True, threads=3, patch_size=30, **CFG2_KW)
    _, _, whole, _ = orc.filter_pe(p, d)
    assert np.array_equal(reduced, whole)
    for name, st in (("a", reduced), ("b", whole)):
        os.makedirs(tmp_path / name)
        assert engine_lib.snk_report_write_pe(C.byref(p), np.ascontiguousarray(st).ctypes.data, str(tmp_path / name).encode()) == 0
    for f in glob.glob(str(tmp_path / "b" / "*.txt")):
        assert filecmp.cmp(f, str(tmp_path / "a" / os.path.basename(f)), shallow=False)


def test_shard_batches_round_robin():
    from soapnuke_b200 import dist as snkdist
    got = sorted(sum((snkdist.shard_batches(11, r, 4) for r in range(4)), []))
    assert got == list(range(11))
    assert snkdist.shard_batches(11, 1, 4) == [1, 5, 9]
