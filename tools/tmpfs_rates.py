"""Diagnostic (GPU box host): write / read rate of a fresh tmpfs file with 1..16 threads (os.pwrite / os.preadv
release the GIL): the storage ceiling of the file-to-file leg."""
import os
import sys
import threading
import time

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2 << 30
path = "/dev/shm/nnp_rate_test.bin"
buf = bytearray(os.urandom(1 << 20)) * 64  # 64 MiB


def run(fn, ways):
    per = size // ways
    th = [threading.Thread(target=fn, args=(i * per, per)) for i in range(ways)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    return size / (time.perf_counter() - t0) / 1e9


for ways in (1, 2, 4, 8, 16):
    if os.path.exists(path):
        os.remove(path)
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)

    def wr(off, n):
        done = 0
        while done < n:
            done += os.pwrite(fd, memoryview(buf)[: min(len(buf), n - done)], off + done)

    def rd(off, n):
        done = 0
        b = bytearray(64 << 20)
        while done < n:
            done += os.preadv(fd, [memoryview(b)[: min(len(b), n - done)]], off + done)

    w = run(wr, ways)
    w2 = run(wr, ways)  # rewrite of existing pages
    r = run(rd, ways)
    os.close(fd)
    print(f"{ways:2d} threads: fresh write {w:6.2f} GB/s, rewrite {w2:6.2f} GB/s, read {r:6.2f} GB/s")
os.remove(path)
print("cpus", os.cpu_count())
