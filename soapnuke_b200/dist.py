"""Multi-GPU plumbing: batches shard embarrassingly across ranks (one process per GPU); the only
exchange of the path is the final statistics table.

All words of the table are additive counters except the per-file LAST_KEY word
(max over reads of (global_index+1)<<16 | length), which is max-reduced.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import abi


def last_key_positions(n_slots):
    pos = []
    for s in range(n_slots):
        for f in range(abi.FILE_COUNT):
            pos.append(s * abi.SLOT_WORDS + abi.slot_file_off(f) + abi.FILE_GS_OFF + abi.GS_LAST_KEY)
    return pos


def allreduce_stats(table: torch.Tensor, n_slots: int) -> torch.Tensor:
    """In-place all-reduce of an int64 view of the statistics table (n_slots * SLOT_WORDS words).
    One SUM all-reduce for the counters plus one tiny MAX all-reduce for the LAST_KEY words."""
    assert table.dtype == torch.int64 and table.numel() == n_slots * abi.SLOT_WORDS
    idx = torch.tensor(last_key_positions(n_slots), dtype=torch.long, device=table.device)
    keys = table[idx].clone()
    table[idx] = 0
    dist.all_reduce(table, op=dist.ReduceOp.SUM)
    dist.all_reduce(keys, op=dist.ReduceOp.MAX)
    table[idx] = keys
    return table


def shard_batches(n_batches: int, rank: int, world: int):
    """Batch k of the input goes to rank k % world (round-robin keeps input order reconstructible)."""
    return [k for k in range(n_batches) if k % world == rank]
