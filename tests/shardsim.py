"""TEST INFRASTRUCTURE: a CPU stand-in for the three sharded-compression entry points of the C ABI
(nnp_shard_compress_begin_dev / _orbit / _emit_dev), built on the oracle, so that the rank
orchestration of nnue_data_compress_b200.sharding can be exercised with gloo on CPU."""
from refutil import BIN_TO_BINPACK, oracle, oracle_convert

NO_CARRY = (1 << 64) - 1
THRESHOLD = 1 << 20


class OracleShard:
    def __init__(self, records: bytes, own_lo: int, own_hi: int, reaches_eof: bool):
        n = len(records) // 40
        lib = oracle()
        recs = [records[40 * i:40 * i + 40] for i in range(n)]
        head = [i == 0 or lib.orc_is_continuation(recs[i - 1], recs[i]) != 1 for i in range(n)]

        def find(start):
            for i in range(start, n):
                if head[i]:
                    return i
            return n

        first, end = find(own_lo), find(own_hi)
        if end >= n and not reaches_eof:
            raise RuntimeError("window")
        first = min(first, end)
        self.first, self.end = first, end
        self.payload = b""
        self.head_off = []
        i = first
        while i < end:
            j = i + 1
            while j < end and not head[j]:
                j += 1
            rc, chunk = oracle_convert(BIN_TO_BINPACK, b"".join(recs[i:j]))
            assert rc == 0 and chunk[:4] == b"BINP" and len(chunk) - 8 == int.from_bytes(chunk[4:8], "little")
            self.head_off.append(len(self.payload))
            self.payload += chunk[8:]
            i = j
        self.payload_bytes = len(self.payload)
        self.starts = []
        self.base = 0

    def orbit(self, base, carry):
        self.base = base
        self.starts = []
        last = None if carry == NO_CARRY else carry
        for off in self.head_off:
            g = base + off
            if last is None or g - last >= THRESHOLD:
                self.starts.append(off)
                last = g
        first = base + self.starts[0] if self.starts else NO_CARRY
        carry_out = base + self.starts[-1] if self.starts else carry
        return len(self.starts), first, carry_out

    def emit(self, next_start):
        out = bytearray()
        cuts = self.starts + [self.payload_bytes]
        out += self.payload[:cuts[0]]
        for k in range(len(self.starts)):
            size = cuts[k + 1] - cuts[k]
            if k == len(self.starts) - 1:
                size = next_start - (self.base + cuts[k])
            out += b"BINP" + size.to_bytes(4, "little") + self.payload[cuts[k]:cuts[k + 1]]
        return bytes(out)

    # -- the table form of the carry chain (nnp_shard_compress_table_dev / _resolve_dev) ----------

    def table(self):
        import torch

        from nnue_data_compress_b200.sharding import ORBIT_TABLE_ENTRIES

        nxt = []
        for h, off in enumerate(self.head_off):
            j = h + 1
            while j < len(self.head_off) and self.head_off[j] - off < THRESHOLD:
                j += 1
            nxt.append(j)
        rows = []
        for e in range(ORBIT_TABLE_ENTRIES):
            if e < len(self.head_off) and (e == 0 or self.head_off[e - 1] < THRESHOLD):
                cur, last, count = e, 0, 0
                while cur < len(self.head_off):
                    last = self.head_off[cur]
                    count += 1
                    cur = nxt[cur]
                rows += [self.head_off[e], last, count]
            else:
                rows += [-1, 0, 0]
        return torch.tensor(rows, dtype=torch.int64)

    @staticmethod
    def resolve(tables, sizes, world, rank):
        from nnue_data_compress_b200.sharding import ORBIT_TABLE_ENTRIES

        t = tables.tolist()
        base, carry, chunks, next_start = 0, None, 0, None
        mine = None
        for r in range(world):
            if r == rank:
                mine = (NO_CARRY if carry is None else carry, chunks)
            if sizes[r]:
                target = 0 if carry is None else max(carry + THRESHOLD - base, 0)
                rows = t[r * ORBIT_TABLE_ENTRIES * 3:(r + 1) * ORBIT_TABLE_ENTRIES * 3]
                for e in range(ORBIT_TABLE_ENTRIES):
                    first = rows[3 * e]
                    if first < 0:
                        break
                    if first >= target:
                        if r > rank and next_start is None:
                            next_start = base + first
                        carry = base + rows[3 * e + 1]
                        chunks += rows[3 * e + 2]
                        break
            base += sizes[r]
        return mine[0], mine[1], base if next_start is None else next_start, chunks
