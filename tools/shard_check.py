"""Multi-process check of the sharded compressor (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_check.py [positions_per_rank] [plies]

Every rank generates its own records, the ranks exchange one halo record and an overlap window
(what a file reader would read directly), compress them as ONE file with compress_sharded over
NCCL, and rank 0 compares the assembled slices with a single-GPU compression of all records."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nnue_data_compress_b200 as nnp
from nnue_data_compress_b200.sharding import compress_sharded


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    plies = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    overlap = 65536
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nnp.init(local)
    L = nnp.lib()
    nnp.use_torch_stream()
    own = torch.empty(n * 40, dtype=torch.uint8, device=dev)
    assert L.nnp_generate_bin_dev(ctypes.c_void_p(own.data_ptr()), n, plies, 4242 + rank) == 0
    heads = [torch.empty(overlap * 40, dtype=torch.uint8, device=dev) for _ in range(world)]
    tails = [torch.empty(40, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(heads, own[: overlap * 40].contiguous())
    dist.all_gather(tails, own[-40:].contiguous())
    parts = ([tails[rank - 1]] if rank > 0 else []) + [own] + ([heads[rank + 1]] if rank < world - 1 else [])
    buf = torch.cat(parts)
    lo = 1 if rank > 0 else 0
    use_table = os.environ.get("SHARD_TABLE", "1") == "1"
    d_slice = torch.empty(max(buf.numel() // 8 + (1 << 20) if plies > 20 else buf.numel() + (1 << 20), 8), dtype=torch.uint8, device=dev)
    calls = nnp.ShardCalls(d_slice, dev)
    info = calls.begin(buf, buf.numel() // 40, lo, lo + n, rank == world - 1)
    extra = dict(table=calls.table, resolve=calls.resolve) if use_table else {}
    got, off, total = compress_sharded(info.payload_bytes, calls.orbit, calls.emit, device=dev, **extra)
    piece = d_slice[:got]
    # assemble on rank 0 and compare with one single-GPU run over all records
    sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([piece.numel(), off], dtype=torch.int64, device=dev))
    cap = max(int(s[0]) for s in sizes)
    padded = torch.zeros(cap, dtype=torch.uint8, device=dev)
    padded[: piece.numel()] = piece
    pieces = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(pieces, padded)
    owns = [torch.empty(n * 40, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(owns, own)
    ok = True
    if rank == 0:
        whole = torch.cat(owns)
        cap2 = whole.numel() // 8 + (1 << 20) if plies > 20 else whole.numel() + (1 << 20)
        ref = torch.empty(cap2, dtype=torch.uint8, device=dev)
        sz = ctypes.c_size_t(0)
        rc = L.nnp_bin_to_binpack_dev(ctypes.c_void_p(whole.data_ptr()), whole.numel(), ctypes.c_void_p(ref.data_ptr()), cap2,
                                      ctypes.byref(sz))
        assert rc == 0, rc
        asm = torch.zeros(total, dtype=torch.uint8, device=dev)
        for r in range(world):
            ln, o = int(sizes[r][0]), int(sizes[r][1])
            asm[o:o + ln] = pieces[r][:ln]
        ok = sz.value == total and torch.equal(asm, ref[: sz.value])
        print(f"SHARD_CHECK world={world} positions={n * world} plies={plies} table={use_table} file_bytes={total} "
              f"{'IDENTICAL' if ok else 'MISMATCH'} to the single-GPU run", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
