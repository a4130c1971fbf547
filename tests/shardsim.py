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
