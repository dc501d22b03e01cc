#!/usr/bin/env python
"""bench.py — the headline measurement: ALP decode throughput of a decimal-heavy f64 column on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path on the host cores

A "step" is one pass of the decode hot path over one column: BASELINE.json configs[1], 2^30 synthetic decimal-heavy
f64 values (<= 3 decimals; SURVEY.md §8d generator, seed 42) per GPU, already compressed and resident in HBM when the
timed region starts.  N > 1 (torchrun, one rank per GPU) shards the global column by whole row-groups: rank r owns
values [r*2^30, (r+1)*2^30); there is no data-path collective, so scaling is weak.

Printed JSON (rank 0, one line):
  value      decoded GB/s (uncompressed f64 bytes produced per second), whole job, device-timed with CUDA events
  roofline   algorithmic HBM bytes (compressed bytes read + decoded bytes written, summed from the column's own
             per-vector metadata) / measured kernel time, against the measured copy bandwidth of MEASURED_PEAKS.json
  e2e        the same metric through the host-buffer API (alpb200_decompress_host): pinned host column in,
             pinned host values out, copies inside the timed region
  cpu_baseline  the reference's CPU decode (oracle/_ref, all host threads) on a bounded slice of the same column
Beside the contract's keys (reported, not the headline):
  encode        device-timed alpb200_encode_f64 of the same column (vector-order layout, the default) with its roofline
                fraction, the completion-order layout beside it, and the row-group init
  scan_sum      fused decode + SUM (alpb200_decode_sum_f64), bound by the compressed read
  e2e_scan      SUM over the pinned host column through alpb200_sum_host_f64 (only compressed bytes cross PCIe)
  e2e_compress  pinned host values in, pinned host column out through alpb200_compress_host_f64
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "alp_decode_f64_GBps"
UNIT = "GB/s"
KIND = 2  # decimal-heavy f64
DEFAULT_VALUES = 1 << 30
WORKLOAD = "alp_decode 2^30 synthetic decimal-heavy f64 (<=3 decimals) per GPU, 1024-value vectors"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = []
        reasons = set()
        sm_max = None
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                sm_max = float(r[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": sm_max, "reasons": sorted(reasons), "samples": len(sm)}


def pinned(shape, dtype):
    import torch

    t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name) if np.dtype(dtype).kind == "f" else torch.uint8, pin_memory=True)
    return t


def pinned_host_column(h):
    """Copy of a HostColumn whose arrays are page-locked (what a host engine would hand to the codec)."""
    import torch

    from alp_b200 import _abi

    out = _abi.HostColumn(h.n_vectors, h.value_bytes, 128, 1)
    keep = []

    def pin(arr):
        t = torch.empty(max(arr.nbytes, 16) + 128, dtype=torch.uint8, pin_memory=True)
        keep.append(t)
        off = (-t.data_ptr()) % 128
        view = t.numpy()[off : off + arr.nbytes].view(arr.dtype).reshape(arr.shape)
        view[...] = arr
        return view

    out.meta = pin(h.meta)
    out.packed = pin(h.packed[: max(h.packed_bytes, 128)])
    out.exc_val = pin(h.exc_val[: max(h.n_exceptions, 1)])
    out.exc_pos = pin(h.exc_pos[: max(h.n_exceptions, 1)])
    out.totals = h.totals.copy()
    out._keep = keep
    return out


def cpu_decode_baseline(col_host, n_threads, budget_s, checker):
    """Time the CPU checker's decode of a host column, repeated until `budget_s` seconds have passed."""
    n_values = col_host.n_vectors * 1024
    out = np.empty(n_values, dtype=np.float64)
    checker.decode_column(col_host, n_threads=n_threads, out=out)  # warm-up (page faults, caches)
    reps, t0 = 0, time.perf_counter()
    while True:
        checker.decode_column(col_host, n_threads=n_threads, out=out)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or reps >= 200:
            break
    return n_values * 8.0 * reps / dt / 1e9, reps, dt


class JsonChannel:
    """The process's real stdout, reserved for the ONE JSON line: file descriptor 1 is pointed at stderr for everything
    else (NCCL prints a version banner on stdout when the box sets NCCL_DEBUG, libraries print progress, ...)."""

    def __init__(self):
        sys.stdout.flush()
        self._out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)

    def emit(self, obj):
        self._out.write(json.dumps(obj) + "\n")
        self._out.flush()


def run_reference(args):
    """--impl reference: the reference's own CPU decode (falp + patch_exceptions) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    channel = JsonChannel()
    from oracle import pyoracle

    checker = pyoracle.best()
    threads = host_threads()
    n_sample = min(args.values, 1 << 27)  # bounded sample of the same workload
    x = pyoracle.generate(n_sample, KIND)
    col = checker.encode_column(x, n_threads=threads, packed_capacity=n_sample * 4, exc_capacity=n_sample // 8)
    out = np.empty(n_sample, dtype=np.float64)
    for _ in range(max(1, args.warmup)):
        checker.decode_column(col, n_threads=threads, out=out)
    assert out.tobytes() == x.tobytes()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        checker.decode_column(col, n_threads=threads, out=out)
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = n_sample * 8.0 / (ms * 1e-3) / 1e9
    t0 = time.perf_counter()
    checker.decode_column(col, n_threads=1, out=out)  # the reference's own methodology is single-core
    one_thread = n_sample * 8.0 / (time.perf_counter() - t0) / 1e9
    sample = "2^%d values of the same column, all vectors, decode = falp + patch_exceptions" % int(np.log2(n_sample))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_values": n_sample, "host_threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": checker.kind, "sample": sample, "build": checker.build_info,
                         "single_thread_value": one_thread},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    channel.emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="alp_b200", choices=["alp_b200", "reference"])
    ap.add_argument("--values", type=int, default=int(os.environ.get("ALPB200_BENCH_VALUES", DEFAULT_VALUES)), help="values per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "alp_b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    channel = JsonChannel()
    import torch
    import torch.distributed as dist

    import alp_b200
    from alp_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = (args.values // _abi.VECTOR_SIZE) * _abi.VECTOR_SIZE
    n_vec = n // _abi.VECTOR_SIZE
    launches = 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed set-up: this rank's shard of the global column, compressed on the device ----
    x = alp_b200.generate(n, KIND, dev, first_index=rank * n)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    big = alp_b200.DeviceColumn(n_vec, 8, dev)  # worst-case capacities; allocated before anything is timed
    ws = torch.empty(max(256, alp_b200.lib.alpb200_encode_workspace_bytes(n_vec)), dtype=torch.uint8, device=dev)
    states = alp_b200.rowgroup_init(x)
    alp_b200.encode(x, states, col=big, workspace=ws)  # warm-up pass (also first touch of the output buffers)
    alp_b200.encode(x, states, col=big, workspace=ws, ordered=False)
    torch.cuda.synchronize()
    ev[3].record()
    alp_b200.encode(x, states, col=big, workspace=ws, ordered=False)  # completion-order layout (reported beside the default)
    ev[4].record()
    del states  # its blocks go back to torch's caching allocator: no cudaMalloc inside the timed init below
    ev[0].record()
    states = alp_b200.rowgroup_init(x)
    ev[1].record()
    alp_b200.encode(x, states, col=big, workspace=ws)  # the default, vector-order layout: this is the column decoded below
    ev[2].record()
    packed_bytes, n_exc = big.read_totals()
    init_ms, encode_ms, encode_unordered_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[3].elapsed_time(ev[4])
    # trim the worst-case container to what was used
    col = alp_b200.DeviceColumn(n_vec, 8, dev, max(packed_bytes, 128), max(n_exc, 1))
    col.meta.copy_(big.meta)
    col.packed[:packed_bytes].copy_(big.packed[:packed_bytes])
    col.exc_val[:n_exc].copy_(big.exc_val[:n_exc])
    col.exc_pos[:n_exc].copy_(big.exc_pos[:n_exc])
    col.totals.copy_(big.totals)
    col.max_block_bytes = big.max_block_bytes
    del big
    torch.cuda.empty_cache()
    meta = col.meta.cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1)
    out = torch.empty(n, dtype=torch.float64, device=dev)

    # correctness of the very thing that is timed: decode == original, bit for bit
    alp_b200.decode(col, out=out)
    torch.cuda.synchronize()
    verified = bool(torch.equal(out.view(torch.int64), x.view(torch.int64)))
    assert verified, "decoded column differs from the original"

    # algorithmic bytes of one decode launch: 128*bw + 13-byte header + 10 bytes per exception read,
    # 8192 bytes written, per vector — summed from the column's own metadata
    hdr = 13
    read_bytes = int(meta["bw"].astype(np.int64).sum()) * 128 + hdr * n_vec + 10 * int(meta["exc_cnt"].astype(np.int64).sum())
    algo_bytes = read_bytes + n * 8

    # ---- timed region: K decode launches, data resident in HBM (3 GB compressed in, 8 GB out: far beyond L2) ----
    for _ in range(args.warmup):
        alp_b200.decode(col, out=out)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        time.sleep(0.3)
        start.record()
        for _ in range(args.steps):
            alp_b200.decode(col, out=out)
            launches += 1
        stop.record()
        barrier()
        time.sleep(0.2)
    ms_local = start.elapsed_time(stop) / args.steps
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * 8.0 / (ms * 1e-3) / 1e9

    # encode side (reported, not the headline): device-timed, input resident
    enc_gbps = n * 8.0 / (encode_ms * 1e-3) / 1e9

    # fused decode + SUM scan (reported, not the headline): nothing is written back, so the bound is the compressed read
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    for _ in range(3):
        alp_b200.decode_sum(col, out=acc)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        alp_b200.decode_sum(col, out=acc)
    s1.record()
    torch.cuda.synchronize()
    scan_ms = s0.elapsed_time(s1) / args.steps

    # ---- e2e: host column in, host values out, through alpb200_decompress_host ----
    e2e = e2e_scan = e2e_compress = None
    if not args.no_e2e:
        hcol = pinned_host_column(col.to_host())
        hout_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
        hout = hout_t.numpy()
        codec = alp_b200.HostCodec(n_vec, 8, local)
        codec.decompress(hcol, out=hout)  # warm-up + check
        assert hout.view(np.int64)[:: 4099].tobytes() == x.cpu().numpy().view(np.int64)[:: 4099].tobytes()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            codec.decompress(hcol, out=hout)
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        h2d = hcol.meta.nbytes + hcol.packed_bytes + hcol.n_exceptions * 10
        e2e = {
            "value": world * n * 8.0 / e2e_s / 1e9,
            "unit": UNIT,
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(n * 8),
            "ms_per_step": e2e_s * 1e3,
            "steps": args.e2e_steps,
            "api": "alpb200_decompress_host_f64 (pinned host buffers, 16-chunk pipeline over 3 streams)",
        }
        # the scan query end to end: host column in, one double out (only the compressed bytes cross PCIe)
        want_sum = float(x.sum().item())
        got_sum = codec.sum(hcol)
        assert abs(got_sum - want_sum) <= 1e-9 * abs(want_sum), (got_sum, want_sum)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            codec.sum(hcol)
        scan_s = (time.perf_counter() - t0) / args.e2e_steps
        ts = torch.tensor([scan_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        scan_s = float(ts.item())
        e2e_scan = {
            "value": world * n * 8.0 / scan_s / 1e9,
            "unit": "GB/s of decoded-equivalent f64 scanned",
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 8,
            "ms_per_step": scan_s * 1e3,
            "api": "alpb200_sum_host_f64 (SUM over the host column; reference: bench_end_to_end alp_func + aggr_plus)",
        }
        # the other direction end to end: pinned host values in, pinned host column out (chunk pipeline, H2D-bound)
        ref_packed = hcol.packed[: hcol.packed_bytes].tobytes()
        hcol.packed[: hcol.packed_bytes] = 0
        codec.compress(hout, col=hcol)  # warm-up + check: the same column comes back
        assert hcol.packed[: hcol.packed_bytes].tobytes() == ref_packed
        del ref_packed
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            codec.compress(hout, col=hcol)
        comp_s = (time.perf_counter() - t0) / args.e2e_steps
        tc = torch.tensor([comp_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        comp_s = float(tc.item())
        e2e_compress = {
            "value": world * n * 8.0 / comp_s / 1e9,
            "unit": "GB/s of f64 input compressed",
            "h2d_bytes_per_step": int(n * 8),
            "d2h_bytes_per_step": int(h2d),
            "ms_per_step": comp_s * 1e3,
            "api": "alpb200_compress_host_f64 (pinned host buffers; row-group init + encode per chunk, 16-chunk pipeline over 3 streams)",
        }
        codec.close()

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's decode on a bounded slice of this column ----
    cpu = None
    if rank == 0 and world == 1:
        from oracle import pyoracle

        checker = pyoracle.best()
        n_slice_vec = min(n_vec, (1 << 27) // 1024)  # same bounded sample as `--impl reference`
        sl = col.to_host(0, n_slice_vec)
        threads = host_threads()
        gbps, reps, dt = cpu_decode_baseline(sl, threads, args.cpu_seconds, checker)
        one_gbps, _, _ = cpu_decode_baseline(sl, 1, 1.0, checker)
        cpu = {
            "value": gbps,
            "unit": UNIT,
            "cores": threads,
            "kind": checker.kind,
            "single_thread_value": one_gbps,
            "sample": "first 2^%d values of the same column decoded %d times in %.1f s (falp + patch_exceptions, %s)" % (int(np.log2(n_slice_vec * 1024)), reps, dt, checker.build_info),
        }

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = algo_bytes / (ms_local * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get("decode_f64_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "values_per_gpu": n,
                "vectors_per_gpu": n_vec,
                "bits_per_value": 8.0 * (read_bytes) / n,
                "l2": "inputs larger than L2 (%.2f GB compressed in, %.2f GB out per step)" % (read_bytes / 1e9, n * 8 / 1e9),
                "parallelism": "row-group shards, %d rank(s), no data-path collective" % world,
                "verified_bit_exact": verified,
            },
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "kernel": "decode_kernel<double,8>",
            },
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks.summary(),
            "e2e_scan": e2e_scan,
            "e2e_compress": e2e_compress,
            "encode": {
                "GBps": enc_gbps,
                "ms": encode_ms,
                "roofline_frac": algo_bytes / (encode_ms * 1e-3) / 1e9 / peak,
                "rowgroup_init_ms": init_ms,
                "bits_per_value": 8.0 * read_bytes / n,
                "unordered_layout": {"GBps": n * 8.0 / (encode_unordered_ms * 1e-3) / 1e9, "ms": encode_unordered_ms,
                                     "roofline_frac": algo_bytes / (encode_unordered_ms * 1e-3) / 1e9 / peak},
            },
            "scan_sum": {
                "ms": scan_ms,
                "GBps_decoded_equivalent": n * 8.0 / (scan_ms * 1e-3) / 1e9,
                "read_GBps": read_bytes / (scan_ms * 1e-3) / 1e9,
                "roofline_frac": read_bytes / (scan_ms * 1e-3) / 1e9 / peak,
            },
        }
        channel.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
