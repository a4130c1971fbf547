#!/usr/bin/env python
"""bench.py -- throughput of the conversion hot path (.bin -> .binpack and .binpack -> .bin).

    python bench.py --gpus N --steps K --warmup W              # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

One step = one pass of the hot path over one batch: `positions` synthetic positions
(random legal-move games, ~100 plies per chain, generated on the device) are compressed
.bin -> .binpack and the result is decompressed .binpack -> .bin. Both directions count, so
the metric is (2 * positions) / step time in Mpos/s; per-direction numbers are reported too.

value  : inputs resident in HBM, device pointers through the C ABI (*_dev entry points),
         timed with CUDA events on the stream the kernels run on, max over ranks.
e2e    : the same step through the host-buffer entry points (pinned host memory in, pinned
         host memory out; H2D and D2H inside the timed region).
roofline, cpu_baseline: see DESIGN.md "Measurement".

N > 1 (torchrun, one rank per GPU), weak scaling, `positions` per GPU fixed: the ranks write ONE
.binpack, byte-identical to a single reference run over all records (chains belong to the rank of
their head, the chunk-flush rule is replayed from the all-gathered orbit tables), and then decode
that ONE file by contiguous chunk ranges (nnp_shard_decompress_dev + one all-gather of the position
counts). No payload crosses NVLink inside the timed region. `decompress_strong` reports BASELINE
configs[2] beside it: the same 100 M-position file decoded by 1/2/4/8 ranks (strong scaling).

Auxiliary keys at N = 1, all outside the timed regions of `value` and `e2e`: `parity_checked` (the
compiled reference run on the first records of this run's own input, compared with the GPU output),
`plain` (BASELINE configs[3]), `sweep` (configs[4] chain lengths 1 / 8 / 64 / 400, plus chain length 1 as training
pipelines produce it: the positions of 100-ply games in random order), `halfkp`, `e2e_file`.
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpos/s bin->binpack & binpack->bin"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "nnue_data_compression")
REF_GEN = os.path.join(ROOT, "oracle", "_ref", "gen_ref")


_T0 = time.time()


def log(msg):
    """progress on stderr (the JSON line is the only thing on stdout)"""
    print(f"[bench +{time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--positions", type=int, default=100_000_000, help="positions per GPU per step")
    ap.add_argument("--plies", type=int, default=100, help="maximum plies per chain of the synthetic games")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-sample", type=int, default=4_000_000, help="positions of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-halfkp", action="store_true", help="skip the auxiliary .binpack -> HalfKP rows measurement")
    ap.add_argument("--no-plain", action="store_true", help="skip the auxiliary .plain directions (BASELINE configs[3])")
    ap.add_argument("--no-sweep", action="store_true", help="skip the auxiliary chain-length sweep (BASELINE configs[4])")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling decode of the 100M file (configs[2])")
    ap.add_argument("--no-file", action="store_true", help="skip the file-to-file leg (tmpfs)")
    ap.add_argument("--plain-positions", type=int, default=10_000_000)
    ap.add_argument("--sweep-positions", type=int, default=0, help="0 = --positions")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU during the timed region: NVML from a thread
    every 5 ms (the timed region of the default run is shorter than one nvidia-smi period), with
    `nvidia-smi --query-gpu` as the fallback when pynvml is missing."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.smax = [], set(), None
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _nvml_loop(self, nvml, handle):
        bits = {
            "hw_slowdown": getattr(nvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        get_reasons = getattr(nvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)))
                r = int(get_reasons(handle))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001 - a failed sample is just a missing sample
                pass
            time.sleep(0.005)

    def _smi_loop(self):
        fields = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={fields}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(parts[0]))
                self.smax = float(parts[1])
                for name, v in zip(self.NAMES, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                time.sleep(0.05)

    def start(self):
        try:
            import pynvml as nvml

            nvml.nvmlInit()
            handle = nvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nvml, handle), daemon=True)
        except Exception:  # noqa: BLE001
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=10)
        sm = sorted(self.sm)
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": self.smax,
            "samples": len(sm),
            "source": self.source,
            "window": "warm-up + timed steps (same workload)",
            "reasons": sorted(self.reasons),
        }


# ------------------------------------------------------------------------------------------------
# the reference's CPU implementation (cpu_baseline leg and --impl reference)


def _ref_shard_job(workdir, idx, positions, plies, seed):
    """Generates one shard with the reference's own chess library (not timed)."""
    path = os.path.join(workdir, f"shard{idx}.bin")
    subprocess.run([REF_GEN, path, str(positions), str(plies), str(seed + idx)], check=True)
    return path


def _ref_round_trip(paths):
    """Runs bin->binpack then binpack->bin with the reference binary, one process per shard, all
    processes in parallel. Returns (compress_s, decompress_s) wall-clock."""
    def run_all(cmds):
        t0 = time.perf_counter()
        procs = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("reference binary failed")
        return time.perf_counter() - t0

    tc = run_all([[REF_BIN, p, p[:-4] + ".binpack"] for p in paths])
    td = run_all([[REF_BIN, p[:-4] + ".binpack", p[:-4] + ".rt.bin"] for p in paths])
    return tc, td


def reference_available():
    return os.access(REF_BIN, os.X_OK) and os.access(REF_GEN, os.X_OK)


def cpu_reference_measure(positions_total, plies, seed, procs, steps, warmup):
    """Times the unmodified reference (oracle/_ref, built from /root/reference by oracle/Makefile)
    on `procs` host processes, each on its own shard of positions_total / procs positions."""
    workdir = tempfile.mkdtemp(prefix="nnp_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        per = max(positions_total // procs, 1000)
        gens = [threading.Thread(target=_ref_shard_job, args=(workdir, i, per, plies, seed)) for i in range(procs)]
        for g in gens:
            g.start()
        for g in gens:
            g.join()
        paths = [os.path.join(workdir, f"shard{i}.bin") for i in range(procs)]
        for _ in range(warmup):
            _ref_round_trip(paths)
        tcs, tds = [], []
        for _ in range(steps):
            tc, td = _ref_round_trip(paths)
            tcs.append(tc)
            tds.append(td)
        n = per * procs
        tc, td = sum(tcs) / len(tcs), sum(tds) / len(tds)
        return {
            "positions": n,
            "procs": procs,
            "compress_mpos_s": n / tc / 1e6,
            "decompress_mpos_s": n / td / 1e6,
            "value": 2 * n / (tc + td) / 1e6,
            "ms_per_step": (tc + td) * 1e3,
        }
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not reference_available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (run oracle/Makefile where /root/reference exists)"}))
        return 0
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    per_proc = 250_000  # ~0.7 s compress + ~0.3 s decompress per process and step
    r = cpu_reference_measure(per_proc * procs, args.plies, args.seed, procs, max(args.steps, 1), max(args.warmup, 0))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": r["value"],
        "unit": "Mpos/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {
            "workload": f"bin->binpack then binpack->bin, random legal-move games, <= {args.plies} plies per chain",
            "positions_per_step": r["positions"],
            "sample": f"{procs} processes x {per_proc} positions (bounded sample of the 100M-position workload)",
        },
        "compress_mpos_s": r["compress_mpos_s"],
        "decompress_mpos_s": r["decompress_mpos_s"],
        "cpu_baseline": {
            "value": r["value"], "unit": "Mpos/s", "cores": procs, "kind": "reference",
            "sample": f"{procs} x {per_proc} positions, reference binary (make release flags, -O2), one process per shard",
        },
        "e2e": {"value": r["value"], "unit": "Mpos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm


def chunk_starts(data):
    """Byte offsets of the BINP chunk headers of an in-memory .binpack (compress_file.cpp:500-521)."""
    out, pos = [], 0
    while pos + 8 <= len(data) and data[pos:pos + 4] == b"BINP":
        out.append(pos)
        pos += 8 + int.from_bytes(data[pos + 4:pos + 8], "little")
    return out


def run_ref(src, dst):
    t0 = time.perf_counter()
    subprocess.run([REF_BIN, src, dst], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0


def scratch_dir():
    return tempfile.mkdtemp(prefix="nnp_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def parity_and_cpu_baseline(d_bin, d_pack, pack_bytes, d_out, sample, plies):
    """The compiled reference (oracle/_ref), one process, on the FIRST `sample` records of this run's
    own input: its timing is the cpu_baseline, its files are the parity check. A reference run over a
    prefix of the records writes the same chunks as the run over all of them except for its last chunk
    (cut at the end of the prefix), so every chunk in front of that one must equal the GPU's .binpack
    byte for byte, and decoding those chunks with the reference must give the first records of the
    GPU's .bin output."""
    work = scratch_dir()
    try:
        p_bin, p_pack, p_rt = (os.path.join(work, x) for x in ("in.bin", "out.binpack", "rt.bin"))
        with open(p_bin, "wb") as f:
            f.write(d_bin[: sample * 40].cpu().numpy().tobytes())
        tc = run_ref(p_bin, p_pack)
        td = run_ref(p_pack, p_rt)
        with open(p_pack, "rb") as f:
            ref_pack = f.read()
        starts = chunk_starts(ref_pack)
        checked = {"positions": 0, "identical": None, "note": "sample too small for a complete chunk"}
        if len(starts) >= 2:
            cut = starts[-1]
            ours = d_pack[:cut].cpu().numpy().tobytes()
            same_pack = cut <= pack_bytes and ours == ref_pack[:cut]
            p_pre, p_pre_rt = os.path.join(work, "pre.binpack"), os.path.join(work, "pre.bin")
            with open(p_pre, "wb") as f:
                f.write(ref_pack[:cut])
            run_ref(p_pre, p_pre_rt)
            with open(p_pre_rt, "rb") as f:
                ref_bin = f.read()
            same_bin = d_out[: len(ref_bin)].cpu().numpy().tobytes() == ref_bin
            checked = {
                "positions": len(ref_bin) // 40,
                "binpack_bytes": cut,
                "chunks": len(starts) - 1,
                "identical": bool(same_pack and same_bin),
                "bin_to_binpack_identical": bool(same_pack),
                "binpack_to_bin_identical": bool(same_bin),
                "how": f"reference binary on the first {sample} records of this run's input; its complete chunks against the "
                       "GPU .binpack prefix, their reference decode against the GPU .bin prefix",
            }
        cpu = {
            "value": 2 * sample / (tc + td) / 1e6, "unit": "Mpos/s", "cores": 1, "kind": "reference",
            "sample": f"first {sample} records of this run's own input ({plies} plies per chain), reference binary one process: "
                      f"bin->binpack {sample / tc / 1e6:.3f} Mpos/s, binpack->bin {sample / td / 1e6:.3f} Mpos/s",
        }
        return cpu, checked
    finally:
        shutil.rmtree(work, ignore_errors=True)



def plain_leg(torch, L, check, ptr, timed, d_bin, n, dev, peak):
    """BASELINE configs[3]: the .plain directions on device buffers, `n` positions (the first records of
    the bench input). Algorithmic bytes of a direction = |input| + |output|; times are
    CUDA events around the whole call, best of 3. The compiled reference converts a 1 M-position
    sample of the same data on one host core: its Mpos/s, and its files against the GPU's."""
    out = {"positions": n}
    size = ctypes.c_size_t(0)

    def call(fn, d_src, n_src, d_dst, cap):
        check(fn(ptr(d_src), n_src, ptr(d_dst) if d_dst is not None else None, cap, ctypes.byref(size)), fn.__name__)
        return size.value

    def pack_of(records):
        nb = records * 40
        cap = nb // 8 + (1 << 20)
        d = torch.empty(cap, dtype=torch.uint8, device=dev)
        rc = L.nnp_bin_to_binpack_dev(ptr(d_bin), nb, ptr(d), cap, ctypes.byref(size))
        if rc == -8:
            cap = size.value + 4096
            d = torch.empty(cap, dtype=torch.uint8, device=dev)
            rc = L.nnp_bin_to_binpack_dev(ptr(d_bin), nb, ptr(d), cap, ctypes.byref(size))
        check(rc, "plain leg: bin->binpack")
        return d, cap, size.value

    def entry(ms, nbytes):
        return {"ms": ms, "mpos_s": n / (ms * 1e-3) / 1e6, "algorithmic_bytes": nbytes,
                "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / peak}

    d_pk, cap, pk = pack_of(n)
    txt = call(L.nnp_binpack_to_plain_dev, d_pk, pk, None, 0)
    d_txt = torch.empty(txt + 64, dtype=torch.uint8, device=dev)
    out["binpack_to_plain"] = entry(timed(lambda: call(L.nnp_binpack_to_plain_dev, d_pk, pk, d_txt, txt + 64)), pk + txt)
    d_pk2 = torch.empty(cap + (1 << 20), dtype=torch.uint8, device=dev)
    ms = timed(lambda: call(L.nnp_plain_to_binpack_dev, d_txt, txt, d_pk2, cap + (1 << 20)))
    out["plain_to_binpack"] = entry(ms, txt + size.value)
    out["plain_bytes"], out["binpack_bytes"] = txt, pk
    d_rec = torch.empty(n * 40, dtype=torch.uint8, device=dev)
    out["plain_to_bin"] = entry(timed(lambda: call(L.nnp_plain_to_bin_dev, d_txt, txt, d_rec, n * 40)), txt + n * 40)
    btxt = call(L.nnp_bin_to_plain_dev, d_bin, n * 40, None, 0)
    if btxt + 64 > d_txt.numel():
        d_txt = torch.empty(btxt + 64, dtype=torch.uint8, device=dev)
    out["bin_to_plain"] = entry(timed(lambda: call(L.nnp_bin_to_plain_dev, d_bin, n * 40, d_txt, d_txt.numel())), btxt + n * 40)
    del d_rec, d_pk2

    if reference_available():
        m = min(1_000_000, n)
        work = scratch_dir()
        try:
            d_s, _, ps = pack_of(m)
            p_pack, p_plain, p_pack2 = (os.path.join(work, x) for x in ("s.binpack", "s.plain", "s2.binpack"))
            with open(p_pack, "wb") as f:
                f.write(d_s[:ps].cpu().numpy().tobytes())
            t1 = run_ref(p_pack, p_plain)
            t2 = run_ref(p_plain, p_pack2)
            with open(p_plain, "rb") as f:
                ref_plain = f.read()
            with open(p_pack2, "rb") as f:
                ref_pack2 = f.read()
            got = call(L.nnp_binpack_to_plain_dev, d_s, ps, d_txt, d_txt.numel())
            same1 = d_txt[:got].cpu().numpy().tobytes() == ref_plain
            d_in = torch.frombuffer(bytearray(ref_plain), dtype=torch.uint8).to(dev)
            got2 = call(L.nnp_plain_to_binpack_dev, d_in, d_in.numel(), d_s, d_s.numel())
            same2 = d_s[:got2].cpu().numpy().tobytes() == ref_pack2
            out["reference"] = {"positions": m, "cores": 1, "binpack_to_plain_mpos_s": m / t1 / 1e6,
                                "plain_to_binpack_mpos_s": m / t2 / 1e6, "binpack_to_plain_identical": bool(same1),
                                "plain_to_binpack_identical": bool(same2)}
        finally:
            shutil.rmtree(work, ignore_errors=True)
    return out


def file_leg(torch, L, check, d_bin, d_pack, pack_bytes, n_pos):
    """File to file from tmpfs through the slab pipeline (csrc/files.cu: reader, H2D, kernels, D2H and
    writer overlap on double buffers): wall-clock seconds of the whole call, best of three (the first call
    also page-locks the staging buffers, which the library keeps). Output compared with the device path."""
    import numpy as np

    work = scratch_dir()
    try:
        p_in, p_pack, p_back = (os.path.join(work, x) for x in ("in.bin", "out.binpack", "back.bin"))
        slab = 256 << 20
        with open(p_in, "wb") as f:
            for off in range(0, n_pos * 40, slab):
                f.write(d_bin[off:off + slab].cpu().numpy().tobytes())
        pos = ctypes.c_uint64(0)
        tc = td = None
        for _ in range(3):
            t0 = time.perf_counter()
            check(L.nnp_bin_to_binpack_file(p_in.encode(), p_pack.encode(), 0, 0, ctypes.byref(pos)), "file bin->binpack")
            t1 = time.perf_counter()
            assert pos.value == n_pos, (pos.value, n_pos)
            check(L.nnp_binpack_to_bin_file(p_pack.encode(), p_back.encode(), 0, 0, ctypes.byref(pos)), "file binpack->bin")
            t2 = time.perf_counter()
            assert pos.value == n_pos, (pos.value, n_pos)
            tc = t1 - t0 if tc is None else min(tc, t1 - t0)
            td = t2 - t1 if td is None else min(td, t2 - t1)
        same = os.path.getsize(p_pack) == pack_bytes and os.path.getsize(p_back) == n_pos * 40
        if same:
            with open(p_pack, "rb") as f:
                same = np.array_equal(np.frombuffer(f.read(), dtype=np.uint8), d_pack[:pack_bytes].cpu().numpy())
        # what one thread reads from the same tmpfs file, for scale
        t0 = time.perf_counter()
        with open(p_in, "rb", buffering=0) as f:
            buf = bytearray(slab)
            while f.readinto(buf):
                pass
        t_read = time.perf_counter() - t0
        return {
            "storage": "tmpfs " + os.path.dirname(work), "devices": int(L.nnp_device_count()),
            "bin_to_binpack": {"s": tc, "mpos_s": n_pos / tc / 1e6, "input_gbs": n_pos * 40 / tc / 1e9},
            "binpack_to_bin": {"s": td, "mpos_s": n_pos / td / 1e6, "output_gbs": n_pos * 40 / td / 1e9},
            "value": 2 * n_pos / (tc + td) / 1e6, "unit": "Mpos/s",
            "binpack_identical_to_device_path": bool(same),
            "one_thread_read_of_the_input_gbs": n_pos * 40 / t_read / 1e9,
        }
    finally:
        shutil.rmtree(work, ignore_errors=True)


def sweep_leg(torch, L, check, ptr, n, seed, dev, peak, d_bin, d_out):
    """BASELINE configs[4]: chain lengths 1, 8, 64 and 400 plies, `n` positions each, device-resident
    round trip (best of 3 per direction, the library's CUDA events), generated into the bench's own
    buffers."""
    size = ctypes.c_size_t(0)
    t_total, t_dom = ctypes.c_float(0), ctypes.c_float(0)
    rows = []
    for plies in (1, "shuffled", 8, 64, 400):
        shuffled = plies == "shuffled"
        if shuffled:
            # chain length 1 as training pipelines produce it: the positions of 100-ply games in random order
            # (the generator's own chain-length-1 file is the start position over and over); some records
            # happen to continue their predecessor, which is what keeps this file on the general route
            check(L.nnp_generate_bin_dev(ptr(d_bin), n, 100, seed + 7), "sweep generate")
            g = torch.Generator(device=dev)
            g.manual_seed(seed)
            rows40 = d_bin[: n * 40].view(torch.int64).view(n, 5)
            rows40.copy_(rows40[torch.randperm(n, device=dev, generator=g)])
            torch.cuda.synchronize()
            plies = 1
        else:
            check(L.nnp_generate_bin_dev(ptr(d_bin), n, plies, seed + plies), "sweep generate")
        cap = n * 40 // 8 + (1 << 20)
        d = torch.empty(cap, dtype=torch.uint8, device=dev)
        rc = L.nnp_bin_to_binpack_dev(ptr(d_bin), n * 40, ptr(d), cap, ctypes.byref(size))
        if rc == -8:
            cap = size.value + 4096
            del d
            d = torch.empty(cap, dtype=torch.uint8, device=dev)
            rc = L.nnp_bin_to_binpack_dev(ptr(d_bin), n * 40, ptr(d), cap, ctypes.byref(size))
        check(rc, "sweep bin->binpack")
        pk = size.value
        c_best = d_best = ck = dk = None
        for _ in range(3):
            check(L.nnp_bin_to_binpack_dev(ptr(d_bin), n * 40, ptr(d), cap, ctypes.byref(size)), "sweep bin->binpack")
            L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
            if c_best is None or t_total.value < c_best:
                c_best, ck = t_total.value, t_dom.value
            c_name = L.nnp_last_dominant_kernel().decode()
            check(L.nnp_binpack_to_bin_dev(ptr(d), pk, ptr(d_out), d_out.numel(), ctypes.byref(size)), "sweep binpack->bin")
            L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
            if d_best is None or t_total.value < d_best:
                d_best, dk = t_total.value, t_dom.value
            d_name = L.nnp_last_dominant_kernel().decode()
            assert size.value == n * 40
        alg = n * 40 + pk
        rows.append({
            "plies": plies, "input": "shuffled game positions" if shuffled else "games", "positions": n, "binpack_bytes": pk, "bytes_per_position": alg / n,
            "compress_ms": c_best, "decompress_ms": d_best, "round_trip_mpos_s": 2 * n / ((c_best + d_best) * 1e-3) / 1e6,
            "compress_frac": alg / (c_best * 1e-3) / 1e9 / peak, "decompress_frac": alg / (d_best * 1e-3) / 1e9 / peak,
            "compress_kernel": c_name, "compress_kernel_ms": ck, "decompress_kernel": d_name, "decompress_kernel_ms": dk,
        })
        del d
    return rows


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import nnue_data_compress_b200 as nnp
    from nnue_data_compress_b200.sharding import compress_sharded, decompress_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the conversion path has no CPU implementation")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL announces its version on stdout when NCCL_DEBUG asks for it; stdout carries the one JSON
        # line only, so file descriptor 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    nnp.init(local_rank)
    L = nnp.lib()
    stream = torch.cuda.current_stream()
    nnp.use_torch_stream()  # kernels on torch's current stream: ordered with torch ops, bracketed by torch events
    L.nnp_last_dominant_kernel.restype = ctypes.c_char_p

    def check(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {L.nnp_strerror(rc).decode()} ({L.nnp_last_cuda_error().decode()})")

    def ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_flag(ok):
        if world == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))

    n_pos = args.positions
    bin_bytes = n_pos * 40
    log(f"rank {rank}: generating {n_pos} positions on the device")
    d_bin = torch.empty(bin_bytes, dtype=torch.uint8, device=dev)
    check(L.nnp_generate_bin_dev(ptr(d_bin), n_pos, args.plies, args.seed + 1000 * rank), "generate")

    def compress_sized(d_records, n_bytes, guess):
        """bin->binpack into a buffer of a guessed size, grown once if the library asks for more"""
        need = ctypes.c_size_t(0)
        d = torch.empty(guess, dtype=torch.uint8, device=dev)
        rc = L.nnp_bin_to_binpack_dev(ptr(d_records), n_bytes, ptr(d), guess, ctypes.byref(need))
        if rc == -8:  # NNP_ERR_CAPACITY: *out_bytes holds the size required
            guess = int(need.value) + 4096
            d = torch.empty(guess, dtype=torch.uint8, device=dev)
            rc = L.nnp_bin_to_binpack_dev(ptr(d_records), n_bytes, ptr(d), guess, ctypes.byref(need))
        check(rc, "sizing pass")
        return d, guess, int(need.value)

    log("generated; sizing pass")
    # the generic capacity bound assumes every record is a chain head (34 B/pos); one sizing pass
    # gives the real size so that the benchmark buffers are not 3.4 GB of slack
    d_pack, cap_guess, pack_bytes = compress_sized(d_bin, bin_bytes, bin_bytes // 8 + (1 << 20))
    k1_kernel = L.nnp_last_dominant_kernel().decode()
    d_out = torch.empty(bin_bytes + bin_bytes // 16 + (64 << 20) if world > 1 else bin_bytes, dtype=torch.uint8, device=dev)
    out_cap = d_out.numel()

    t_total = ctypes.c_float(0)
    t_dom = ctypes.c_float(0)

    # N > 1: the ranks write ONE .binpack, byte-identical to a single reference run over all records
    # (SURVEY.md 8e), and decode that ONE file by chunk ranges. Setup, untimed: every rank gets the last
    # record of the rank before it (halo) and the first records of the rank behind it (overlap window)
    # -- what a file reader would read directly -- so that every chain is owned by exactly one rank;
    # and after a first compression the slices are assembled into the file on every rank (the file a
    # multi-GPU reader would have mapped).
    one_file = None
    d_file = None
    slice_info = {}
    dec_info = {}
    if world > 1:
        window = min(65536, n_pos)
        heads = [torch.empty(window * 40, dtype=torch.uint8, device=dev) for _ in range(world)]
        tails = [torch.empty(40, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(heads, d_bin[: window * 40].contiguous())
        dist.all_gather(tails, d_bin[-40:].contiguous())
        parts = ([tails[rank - 1]] if rank > 0 else []) + [d_bin] + ([heads[rank + 1]] if rank < world - 1 else [])
        d_buf = torch.cat(parts)
        del heads, tails, parts
        own_lo = 1 if rank > 0 else 0
        d_slice = torch.empty(cap_guess, dtype=torch.uint8, device=dev)
        calls = nnp.ShardCalls(d_slice, dev)

        def one_file():
            calls.begin(d_buf, d_buf.numel() // 40, own_lo, own_lo + n_pos, rank == world - 1)
            L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
            dom = t_dom.value
            got, off, total = compress_sharded(calls.last_payload_bytes, calls.orbit, calls.emit, device=dev, table=calls.table,
                                               resolve=calls.resolve, status=calls.last_status)
            slice_info.update(bytes=got, offset=off, file_bytes=total)
            return dom

        one_file()
        torch.cuda.synchronize()
        mine = torch.tensor([slice_info["offset"], slice_info["bytes"]], dtype=torch.int64, device=dev)
        allc = torch.empty(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine)
        layout = allc.tolist()
        widest = max(layout[1::2])
        padded = torch.zeros(widest, dtype=torch.uint8, device=dev)
        padded[: slice_info["bytes"]] = d_slice[: slice_info["bytes"]]
        gathered = torch.empty(world * widest, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, padded)
        d_file = torch.empty(slice_info["file_bytes"], dtype=torch.uint8, device=dev)
        for r in range(world):
            off, nb = layout[2 * r], layout[2 * r + 1]
            d_file[off:off + nb] = gathered[r * widest:r * widest + nb]
        del gathered, padded
        log(f"one file of {d_file.numel()} bytes assembled on every rank")

    def shard_decode(d_src, n_src, w, r):
        got = ctypes.c_size_t(0)
        rng = nnp.ChunkRange()
        check(L.nnp_shard_decompress_dev(ptr(d_src), n_src, w, r, ptr(d_out), out_cap, ctypes.byref(got), ctypes.byref(rng)),
              "sharded binpack->bin")
        L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
        return got.value

    def step_device():
        n1 = ctypes.c_size_t(0)
        if one_file is None:
            check(L.nnp_bin_to_binpack_dev(ptr(d_bin), bin_bytes, ptr(d_pack), cap_guess, ctypes.byref(n1)), "bin->binpack")
            L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
            c_ms, c_dom = t_total.value, t_dom.value
            n2 = ctypes.c_size_t(0)
            check(L.nnp_binpack_to_bin_dev(ptr(d_pack), n1.value, ptr(d_out), out_cap, ctypes.byref(n2)), "binpack->bin")
            L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
            assert n2.value == bin_bytes, (n2.value, bin_bytes)
            return c_ms, c_dom, t_total.value, t_dom.value
        ea, eb, ec = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ea.record(stream)
        c_dom = one_file()
        eb.record(stream)
        got, off, total = decompress_sharded(lambda w, r: shard_decode(d_file, d_file.numel(), w, r), device=dev)
        ec.record(stream)
        dec_info.update(bytes=got, offset=off, file_bytes=total)
        d_dom = t_dom.value
        torch.cuda.synchronize()
        assert total == world * bin_bytes, (total, world * bin_bytes)
        return ea.elapsed_time(eb), c_dom, eb.elapsed_time(ec), d_dom

    log(f"binpack is {pack_bytes} bytes; warm-up")
    # the clock sampler runs from the first warm-up step to the end of the timed region: the same
    # workload throughout, and the timed region alone (tens of ms) is too short for NVML's latency
    visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    phys = local_rank
    if visible:
        ids = [v for v in visible.split(",") if v.strip()]
        if local_rank < len(ids) and ids[local_rank].strip().isdigit():
            phys = int(ids[local_rank])
    sampler = ClockSampler(phys)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    log("timed region")
    launches0 = L.nnp_kernel_launches()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    c_ms = c_dom = d_ms = d_dom = 0.0
    for _ in range(args.steps):
        a, b, c, d = step_device()
        c_ms += a
        c_dom += b
        d_ms += c
        d_dom += d
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = int(L.nnp_kernel_launches() - launches0)
    elapsed_ms = ev0.elapsed_time(ev1)
    K = max(args.steps, 1)
    c_ms, c_dom, d_ms, d_dom = c_ms / K, c_dom / K, d_ms / K, d_dom / K
    dec_kernel = L.nnp_last_dominant_kernel().decode()
    log(f"device-resident: {elapsed_ms / K:.2f} ms/step")

    # ---- N > 1, untimed: the ranks' pieces of the decoded file against ONE single-GPU decode of it
    one_file_check = None
    if world > 1:
        piece = d_out[: dec_info["bytes"]]
        sums = torch.zeros(1, dtype=torch.int64, device=dev)
        sums[0] = piece.view(torch.int64).sum()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        ok = True
        how = "not checked: the whole .bin does not fit next to the bench buffers"
        if rank == 0 and world * bin_bytes <= (48 << 30):
            whole = torch.empty(world * bin_bytes, dtype=torch.uint8, device=dev)
            n2 = ctypes.c_size_t(0)
            check(L.nnp_binpack_to_bin_dev(ptr(d_file), d_file.numel(), ptr(whole), whole.numel(), ctypes.byref(n2)), "whole file")
            ok = (n2.value == world * bin_bytes and bool(torch.equal(whole[: dec_info["bytes"]], piece))
                  and int(whole.view(torch.int64).sum().item()) == int(sums.item()))
            how = ("rank 0 decodes the whole file on one GPU: its own piece byte for byte, all pieces by the sum of their "
                   "64-bit words")
            del whole
        one_file_check = {"identical": ok, "how": how}

    # ---- BASELINE configs[2]: the 100M-position file decoded by 1/2/4/8 ranks (strong scaling)
    strong = None
    if not args.no_strong:
        if world > 1:
            nb = torch.tensor([pack_bytes], dtype=torch.int64, device=dev)
            dist.broadcast(nb, 0)
            d_small = torch.empty(int(nb.item()), dtype=torch.uint8, device=dev)
            if rank == 0:
                d_small.copy_(d_pack[:pack_bytes])
            dist.broadcast(d_small, 0)
        else:
            d_small = d_pack[:pack_bytes]
        res = {}

        def strong_step():
            got, off, total = decompress_sharded(lambda w, r: shard_decode(d_small, d_small.numel(), w, r), device=dev)
            res.update(bytes=got, offset=off, total=total)

        for _ in range(max(args.warmup, 1)):
            strong_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(K):
            strong_step()
        s1.record(stream)
        barrier()
        s_ms = allmax(s0.elapsed_time(s1) / K)
        ok = res["total"] == bin_bytes
        if world > 1:  # every rank checks its piece against its own single-GPU decode of the file
            piece = d_out[: res["bytes"]].clone()
            n2 = ctypes.c_size_t(0)
            check(L.nnp_binpack_to_bin_dev(ptr(d_small), d_small.numel(), ptr(d_out), out_cap, ctypes.byref(n2)), "strong check")
            ok = ok and n2.value == bin_bytes and bool(torch.equal(d_out[res["offset"]:res["offset"] + res["bytes"]], piece))
            del piece
        ok = allmin_flag(ok)
        strong = {
            "workload": f"binpack->bin of the ONE {n_pos}-position file (rank 0's), contiguous chunk ranges over {world} GPU(s), "
                        "one all-gather of the position counts",
            "positions": n_pos, "n_gpus": world, "ms": s_ms, "mpos_s": n_pos / (s_ms * 1e-3) / 1e6,
            "identical_to_single_gpu": ok, "scaling": "strong",
        }
        if world > 1:
            del d_small

    # ---- e2e: host buffers through the public C ABI, H2D and D2H inside the timed region
    e2e = None
    pinned_ok = 0
    if not args.no_e2e:
        # pinned host buffers (cudaHostAlloc through torch); nnp_host_alloc() hands out the same kind
        try:
            t_bin = torch.empty(bin_bytes, dtype=torch.uint8, pin_memory=True)
            t_pack = torch.empty(cap_guess, dtype=torch.uint8, pin_memory=True)
            t_out = torch.empty(bin_bytes, dtype=torch.uint8, pin_memory=True)
            pinned_ok = 1
        except RuntimeError as exc:  # not enough lockable host memory for N ranks x 8.5 GB
            log(f"cannot pin host memory for the e2e leg: {exc}")
            pinned_ok = 0
        pinned_ok = int(allmin_flag(pinned_ok))
    if not args.no_e2e and pinned_ok:
        t_bin.copy_(d_bin)  # untimed setup: the step's input starts in host memory
        torch.cuda.synchronize()
        h_bin, h_pack, h_out = t_bin.data_ptr(), t_pack.data_ptr(), t_out.data_ptr()

        def step_host():
            n1 = ctypes.c_size_t(0)
            check(L.nnp_bin_to_binpack(ctypes.c_void_p(h_bin), bin_bytes, ctypes.c_void_p(h_pack), cap_guess,
                                       ctypes.byref(n1)), "host bin->binpack")
            n2 = ctypes.c_size_t(0)
            check(L.nnp_binpack_to_bin(ctypes.c_void_p(h_pack), n1.value, ctypes.c_void_p(h_out), bin_bytes,
                                       ctypes.byref(n2)), "host binpack->bin")
            return n1.value, n2.value

        e_steps = max(1, min(args.steps, 3))
        log("e2e warm-up")
        step_host()
        log("e2e timed")
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e_steps):
            n1, n2 = step_host()
        e1.record(stream)
        barrier()
        e_ms = e0.elapsed_time(e1) / e_steps
        e2e = {"ms": e_ms, "h2d": bin_bytes + n1, "d2h": n1 + n2}
        del t_bin, t_pack, t_out

    # ---- cpu_baseline + parity_checked (N = 1): the compiled reference on the first records of this run's own
    # input, after all timed regions and before the auxiliary legs reuse the buffers
    cpu_baseline = parity = None
    if not args.no_cpu_baseline and world == 1 and reference_available():
        log("cpu baseline + parity")
        cpu_baseline, parity = parity_and_cpu_baseline(d_bin, d_pack, pack_bytes, d_out, min(args.cpu_sample, n_pos), args.plies)
    elif not args.no_cpu_baseline and world == 1:
        cpu_baseline = {"value": None, "unit": "Mpos/s", "cores": 0, "kind": "reference",
                        "sample": "oracle/_ref missing on this box"}

    # ---- auxiliary (N = 1, outside every timed region above)
    halfkp = plain = sweep = e2e_file = None
    peak, peak_src = 6650.0, "fallback"
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peaks = json.load(f)
        if "hbm_gbs" in peaks:
            peak, peak_src = float(peaks["hbm_gbs"]), "measured"

    def timed(fn, reps=3):
        """best of `reps`, CUDA events on the stream the library's kernels run on"""
        best = None
        for _ in range(reps):
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x0.record(stream)
            fn()
            x1.record(stream)
            torch.cuda.synchronize()
            ms = x0.elapsed_time(x1)
            best = ms if best is None else min(best, ms)
        return best

    if world == 1 and not args.no_halfkp:
        try:
            log("halfkp rows")
            white = torch.empty((n_pos, 32), dtype=torch.int32, device=dev)
            black = torch.empty((n_pos, 32), dtype=torch.int32, device=dev)
            meta = torch.empty((n_pos, 8), dtype=torch.uint8, device=dev)
            cnt = ctypes.c_size_t(0)
            for _ in range(3):
                check(L.nnp_binpack_to_halfkp_dev(ptr(d_pack), pack_bytes, ptr(white), ptr(black), ptr(meta), n_pos,
                                                  ctypes.byref(cnt)), "binpack->halfkp")
                L.nnp_last_timing(ctypes.byref(t_total), ctypes.byref(t_dom))
            assert cnt.value == n_pos, (cnt.value, n_pos)
            halfkp = {
                "mpos_s": n_pos / (t_total.value * 1e-3) / 1e6,
                "ms": t_total.value,
                "kernel": "k_emit_chains_halfkp_verify",
                "kernel_ms": t_dom.value,
                "bytes_per_position": 264,
                "output_gbs": 264 * n_pos / (t_dom.value * 1e-3) / 1e9,
                "frac_of_hbm_peak": (264 * n_pos + pack_bytes) / (t_dom.value * 1e-3) / 1e9 / peak,
                "note": "device-resident .binpack -> two int32[32] index rows + 8 B of targets per position; third call timed "
                        "by the library's CUDA events; not part of `value`",
            }
            del white, black, meta
        except Exception as e:  # auxiliary: never fails the bench line
            halfkp = {"error": str(e)[:200]}

    if world == 1 and not args.no_plain:
        try:
            log("plain directions")
            plain = plain_leg(torch, L, check, ptr, timed, d_bin, min(args.plain_positions, n_pos), dev, peak)
        except Exception as e:  # noqa: BLE001
            plain = {"error": str(e)[:300]}

    if world == 1 and not args.no_file:
        try:
            log("file to file")
            e2e_file = file_leg(torch, L, check, d_bin, d_pack, pack_bytes, n_pos)
        except Exception as e:  # noqa: BLE001
            e2e_file = {"error": str(e)[:300]}

    if world == 1 and not args.no_sweep:
        try:
            log("chain-length sweep")
            del d_pack
            sweep = sweep_leg(torch, L, check, ptr, min(args.sweep_positions or n_pos, n_pos), args.seed, dev, peak, d_bin, d_out)
            # the sweep generated into the bench's own buffers: nothing below may use d_bin / d_out / d_pack
        except Exception as e:  # noqa: BLE001
            sweep = {"error": str(e)[:300]}

    # ---- the line
    elapsed_ms = allmax(elapsed_ms)
    ms_per_step = elapsed_ms / K
    total_pos = n_pos * world
    value = 2 * total_pos / (ms_per_step * 1e-3) / 1e6
    c_ms_m, d_ms_m = allmax(c_ms), allmax(d_ms)
    e2e_ms = allmax(e2e["ms"]) if e2e else None

    shard_offsets = None
    if world > 1:
        # where every rank's slice of the one .binpack and its piece of the one .bin belong
        mine = torch.tensor([slice_info["offset"], slice_info["bytes"], slice_info["file_bytes"], dec_info["offset"],
                             dec_info["bytes"]], dtype=torch.int64, device=dev)
        allc = torch.empty(5 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine)
        flat = allc.tolist()
        shard_offsets = [flat[5 * r:5 * r + 5] for r in range(world)]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel of each direction. SURVEY.md 8(d): algorithmic bytes =
    # |input| + |output| of the direction = 40 B + binpack bytes per position; one launch of the
    # kernel processes all positions of the rank's step. kernel_ms is measured live by the library
    # with CUDA events recorded around that launch on the stream it runs on.
    alg_bytes = bin_bytes + pack_bytes
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)

    def dram_traffic(kernel):
        # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel,
        # scaled from the capture's position count to this run's (both recorded in profiles/traffic.json)
        t = traffic.get(kernel)
        if not t or t.get("plies") != args.plies:
            return None  # no capture of this kernel at this chain length
        return int((t["dram_read_bytes"] + t["dram_write_bytes"]) * (n_pos / t["positions"]))

    def roof(kernel, ms, share_of):
        ach = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else None
        return {"kernel": kernel, "kernel_ms": ms, "achieved": ach, "frac": ach / peak if ach else None,
                "traffic": dram_traffic(kernel), "share_of_direction": ms / share_of if share_of > 0 else None}

    comp = roof(k1_kernel, c_dom, c_ms_m)
    deco = roof(dec_kernel, d_dom, d_ms_m)
    top = comp if c_dom >= d_dom else deco  # the kernel that dominated the step
    roofline = {
        "bound": "hbm",
        "kernel": top["kernel"],
        "achieved": top["achieved"],
        "peak": peak,
        "peak_source": peak_src + (" MEASURED_PEAKS.json hbm_gbs" if peak_src == "measured" else " B200_PROFILING.md 6.65 TB/s"),
        "unit": "GB/s",
        "frac": top["frac"],
        "traffic": top["traffic"],
        "algorithmic_bytes_per_launch": alg_bytes,
        "kernel_ms": top["kernel_ms"],
        "share_of_step": top["kernel_ms"] / ms_per_step,
        "note": "integer-issue bound (see the ncu captures under profiles/), not HBM bound",
        "compress": comp,
        "decompress": deco,
    }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": "Mpos/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u64",
        "data": "synthetic",
        "config": {
            "workload": f"bin->binpack then binpack->bin of {n_pos} positions per GPU, random legal-move games, "
                        f"<= {args.plies} plies per chain (BASELINE configs[1]+[2])",
            "positions_per_gpu": n_pos,
            "bin_bytes_per_gpu": bin_bytes,
            "binpack_bytes_per_gpu": pack_bytes,
            "value_counts": "every position twice per step (once per direction); the reference arm counts the same way",
            "l2": f"inputs ({bin_bytes / 1e9:.1f} GB .bin, {pack_bytes / 1e9:.2f} GB .binpack per GPU) are larger than the 126 MB L2",
            "sharding": ("compress: ONE .binpack over all ranks, byte-identical to a single run (chains owned by the rank of "
                         "their head via halo + overlap window; chunk-flush rule replayed from all-gathered orbit tables); "
                         "decompress: that ONE file by contiguous chunk ranges per rank, one all-gather of the position "
                         "counts; no payload crosses NVLink in the timed region") if world > 1 else "single GPU",
        },
        "compress_mpos_s": total_pos / (c_ms_m * 1e-3) / 1e6,
        "decompress_mpos_s": total_pos / (d_ms_m * 1e-3) / 1e6,
        "compress_ms": c_ms_m,
        "decompress_ms": d_ms_m,
        "clocks": clocks,
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if e2e:
        line["e2e"] = {
            "value": 2 * total_pos / (e2e_ms * 1e-3) / 1e6,
            "unit": "Mpos/s",
            "ms_per_step": e2e_ms,
            "h2d_bytes_per_step": e2e["h2d"],
            "d2h_bytes_per_step": e2e["d2h"],
        }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if parity:
        line["parity_checked"] = parity
    if strong:
        line["decompress_strong"] = strong
    if one_file_check:
        line["one_file_check"] = one_file_check
    for key, val in (("halfkp", halfkp), ("plain", plain), ("sweep", sweep), ("e2e_file", e2e_file)):
        if val:
            line[key] = val
    if shard_offsets is not None:
        line["config"]["binpack_slice_offset_bytes_filebytes__bin_piece_offset_bytes"] = shard_offsets
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
