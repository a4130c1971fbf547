"""Diagnostic (torchrun, one rank per GPU): time per phase of the one-file sharded compressor."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nnue_data_compress_b200 as nnp
from nnue_data_compress_b200.sharding import ORBIT_TABLE_ENTRIES, offsets_from_sizes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nnp.init(local)
nnp.use_torch_stream()
L = nnp.lib()
own = torch.empty(n * 40, dtype=torch.uint8, device=dev)
assert L.nnp_generate_bin_dev(ctypes.c_void_p(own.data_ptr()), n, 100, 42 + 1000 * rank) == 0
W = 65536
heads = [torch.empty(W * 40, dtype=torch.uint8, device=dev) for _ in range(world)]
tails = [torch.empty(40, dtype=torch.uint8, device=dev) for _ in range(world)]
dist.all_gather(heads, own[: W * 40].contiguous())
dist.all_gather(tails, own[-40:].contiguous())
buf = torch.cat(([tails[rank - 1]] if rank > 0 else []) + [own] + ([heads[rank + 1]] if rank < world - 1 else []))
lo = 1 if rank > 0 else 0
d_slice = torch.empty(n * 5 + (1 << 20), dtype=torch.uint8, device=dev)
calls = nnp.ShardCalls(d_slice, dev)
acc = {}


def lap(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0)
    return t


for it in range(8):
    if it == 3:
        acc.clear()
    dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    info = calls.begin(buf, buf.numel() // 40, lo, lo + n, rank == world - 1)
    t = lap("begin", t)
    mine = torch.tensor([int(info.payload_bytes)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    sizes = [int(x.item()) for x in out]
    t = lap("gather sizes", t)
    tab = calls.table()
    t = lap("table", t)
    parts = [torch.empty_like(tab) for _ in range(world)]
    dist.all_gather(parts, tab)
    tables = torch.cat(parts)
    t = lap("gather tables", t)
    carry, before, nxt, tot = calls.resolve(tables, sizes, world, rank)
    t = lap("resolve", t)
    calls.orbit(offsets_from_sizes(sizes)[rank], carry)
    t = lap("orbit", t)
    calls.emit(nxt)
    t = lap("emit", t)
if rank == 0:
    print("SHARD_TIME world", world, {k: round(v / 5 * 1e3, 3) for k, v in acc.items()}, "ms; sum", round(sum(acc.values()) / 5 * 1e3, 3))
dist.barrier()
dist.destroy_process_group()
