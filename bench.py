#!/usr/bin/env python
"""bench.py — the headline measurement: ALP decode throughput of a decimal-heavy f64 column on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path on the host cores

A "step" is one pass of the decode hot path over one column: BASELINE.json configs[1], 2^30 synthetic decimal-heavy
f64 values (<= 3 decimals; SURVEY.md §8d generator, seed 42) per GPU, already compressed and resident in HBM when the
timed region starts.  N > 1 (torchrun, one rank per GPU) shards the global column by whole row-groups: rank r owns
values [r*2^30, (r+1)*2^30); there is no data-path collective, so scaling is weak.

Printed JSON (rank 0, one line) — the contract's keys:
  value      decoded GB/s (uncompressed f64 bytes produced per second), whole job, device-timed with CUDA events
  roofline   algorithmic HBM bytes (compressed bytes read + decoded bytes written, summed from the column's own
             per-vector metadata) / measured kernel time, against the measured copy bandwidth of MEASURED_PEAKS.json
  e2e        the same metric through the host-buffer API (alpb200_decompress_host): pinned host column in, pinned host
             values out, copies inside the timed region; `link` = pinned-copy bandwidths measured in the same run at the
             same N (all ranks at once), `link_frac` = the time the two copies need at those rates / the measured time
  cpu_baseline  the reference's CPU decode (oracle/_ref, all host threads) on a bounded slice of the same column
and beside them (reported, not the headline):
  configs    one block per BASELINE config 2 / 3 / 4 (decimal f64, high-precision f64 = ALP_RD, mixed f32): decode,
             encode (vector order and completion order), row-group init, init + encode, fused decode + SUM — each
             device-timed with its roofline fraction and a bit-exact check — and, at N = 1, the reference's CPU decode /
             encode / scan on a bounded slice of the same column (`cpu`)
  e2e_scan / e2e_compress   SUM over the pinned host column (alpb200_sum_host_f64) and host values in -> host column out
             (alpb200_compress_host_f64), with the reference's CPU rate for the same job beside them
  scatter    N > 1: rank 0 hands its compressed column out as row-group shards over NCCL (outside the timed region)
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
# BASELINE.json configs measured beside the headline: kind -> (label, value bytes, values relative to --values)
SIDE_CONFIGS = {
    2: ("config 2: decimal-heavy f64 (ALP)", 8, 1.0),
    3: ("config 3: high-precision f64 (ALP_RD)", 8, 1.0),
    4: ("config 4: mixed decimal / exception f32 (ALP)", 4, 0.25),
}
CPU_SAMPLE_VALUES = 1 << 27  # bounded sample of a column for the CPU legs


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


def workload_config(n_values):
    """The `config` object: what the workload IS — identical in both arms (the driver compares them)."""
    return {
        "workload": WORKLOAD,
        "values_per_gpu": int(n_values),
        "vectors_per_gpu": int(n_values // 1024),
        "generator": "splitmix64 seed 42, k mod 10^6 / 10^d, d = (i / 102400) mod 4 (SURVEY.md section 8d)",
        "l2": "inputs larger than L2 / LLC (about 2.75 B/value compressed in, 8 B/value out per step)",
    }


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


def algorithmic_read_bytes(meta, value_bytes):
    """Compressed bytes a decode reads, SURVEY.md §8d, summed from the column's own per-vector records: ALP vectors
    128*bw + header (bw, e, f, count, base = 13 / 9 B) + (S+2) B per exception; ALP_RD vectors 128*(right_bw + left_bw)
    + 4-byte header + 4 B per exception, plus the 16-byte dictionary once per row-group."""
    alp = meta["scheme"] == 2
    bw = meta["bw"].astype(np.int64)
    exc = meta["exc_cnt"].astype(np.int64)
    alp_bytes = int((128 * bw[alp] + (5 + value_bytes) + (value_bytes + 2) * exc[alp]).sum())
    rd_bytes = int((128 * (bw[~alp] + meta["e"][~alp].astype(np.int64)) + 4 + 4 * exc[~alp]).sum())
    rd_groups = int((~alp)[::100].sum())
    return alp_bytes + rd_bytes + 16 * rd_groups


def time_cpu(fn, budget_s, max_reps=50):
    """Seconds per call of fn(): one warm-up, then repeated until `budget_s` has passed (at least once)."""
    fn()
    reps, t0 = 0, time.perf_counter()
    while True:
        fn()
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or reps >= max_reps:
            return dt / reps, reps


def ref_bench_binary():
    """oracle/_ref/ref_bench_v{4,3}: the reference compiled as an executable (oracle/ref_bench_main.cpp), or None."""
    from oracle import pyoracle

    names = ["ref_bench_v4", "ref_bench_v3"] if pyoracle._cpu_has_avx512() else ["ref_bench_v3"]
    for name in names:
        path = os.path.join(ROOT, "oracle", "_ref", name)
        if os.path.exists(path) and os.access(path, os.X_OK):
            return path
    return None


def cpu_legs(kind, n_values, threads, budget_s, steps=0, warmup=1, x_host=None):
    """The reference's CPU path on the first n_values values of a synthetic column, all host threads: decode (falp +
    patch_exceptions, or unffor x2 + rd decode), encode with given states (encode + analyze_ffor + ffor / rd encode + 2x ffor,
    test/test_alp_sample.cpp:141-145,164-166), row-group init, and the scan query (alp_func + aggr_plus, q1.cpp:63-100).
    Runs oracle/_ref/ref_bench_* (the reference as an executable: its thread-local encoder scratch is slow inside a dlopen'ed
    library); without it, the same drivers in-process through oracle/pyoracle.py.  GB/s = bytes of uncompressed values per second."""
    vb = SIDE_CONFIGS[kind][1]
    n = int(n_values) // 1024 * 1024
    nbytes = n * vb
    exe = ref_bench_binary()
    if exe is not None:
        cmd = [exe, "--kind", str(kind), "--values", str(n), "--threads", str(threads), "--seconds", "%.3f" % budget_s]
        if steps > 0:
            cmd += ["--steps", str(steps), "--warmup", str(warmup)]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
        if res.returncode != 0:
            raise RuntimeError("ref_bench failed (%d): %s" % (res.returncode, res.stderr[-500:]))
        r = json.loads(res.stdout.strip().splitlines()[-1])
        t_dec, t_enc, t_init, t_scan = r["decode_s"], r["encode_s"], r["init_s"], r["scan_s"]
        leg = {"kind": "reference", "build": r["build"], "round_trip_bit_exact": bool(r["round_trip_bit_exact"]), "decode_reps": r["decode_reps"],
               "decode_single_thread_GBps": vb / r["decode_single_thread_s_per_value"] / 1e9, "scan_sum": r["scan_sum"], "runner": os.path.relpath(exe, ROOT)}
    else:
        from alp_b200 import _abi
        from oracle import pyoracle

        checker = pyoracle.best()
        if x_host is None:
            x_host = pyoracle.generate(n, kind, n_threads=threads)
        x_host = x_host[:n]
        units = 67 if vb == 8 else 35
        col = _abi.HostColumn(n // 1024, vb, (n // 1024 + threads * 100) * units * 128, n + threads * 102400)
        states = checker.bench_init(x_host, n_threads=threads)
        t_init, _ = time_cpu(lambda: checker.bench_init(x_host, n_threads=threads, states=states), budget_s / 4)
        t_enc, _ = time_cpu(lambda: checker.bench_encode(x_host, n_threads=threads, states=states, col=col), budget_s / 4)
        out = np.empty(n, dtype=x_host.dtype)
        if steps > 0:
            for _ in range(max(1, warmup)):
                checker.decode_column(col, n_threads=threads, out=out)
            t0 = time.perf_counter()
            for _ in range(steps):
                checker.decode_column(col, n_threads=threads, out=out)
            t_dec, reps = (time.perf_counter() - t0) / steps, steps
        else:
            t_dec, reps = time_cpu(lambda: checker.decode_column(col, n_threads=threads, out=out), budget_s / 4, 200)
        ok = out.tobytes() == x_host.tobytes()
        got = [0.0]

        def scan():
            got[0] = checker.sum_column(col, n_threads=threads)

        t_scan, _ = time_cpu(scan, budget_s / 4)
        n1 = min(col.n_vectors, 4096)
        t_dec1, _ = time_cpu(lambda: checker.decode_column(col, n=n1, n_threads=1, out=out), 0.3)
        leg = {"kind": checker.kind, "build": checker.build_info, "round_trip_bit_exact": bool(ok), "decode_reps": reps,
               "decode_single_thread_GBps": n1 * 1024 * vb / t_dec1 / 1e9, "scan_sum": got[0], "runner": "in-process (oracle/pyoracle.py)"}
    leg.update({
        "cores": threads,
        "sample": "first 2^%.2f values of the same column" % np.log2(max(n, 1)),
        "decode_ms": t_dec * 1e3,
        "decode_GBps": nbytes / t_dec / 1e9,
        "encode_GBps": nbytes / t_enc / 1e9,
        "rowgroup_init_ms": t_init * 1e3,
        "init_plus_encode_GBps": nbytes / (t_enc + t_init) / 1e9,
        "scan_sum_GBps": nbytes / t_scan / 1e9,
    })
    return leg


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
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores — the headline is its
    decode (falp + patch_exceptions) of the SAME 2^30-value column; encode / scan and configs 3 / 4 ride along."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    channel = JsonChannel()
    threads = host_threads()
    n = args.values // 1024 * 1024
    # the headline leg: K timed decode passes over the WHOLE column after W warm-up passes (encode / init / scan ride along)
    head = cpu_legs(KIND, n, threads, args.cpu_seconds, steps=args.steps, warmup=max(1, args.warmup))
    verified = head["round_trip_bit_exact"]
    assert verified, "the reference's decode of its own column differs from the original"
    ms = head["decode_ms"]
    value = n * 8.0 / (ms * 1e-3) / 1e9
    side = {str(KIND): dict(head, label=SIDE_CONFIGS[KIND][0])}
    for kind, (label, vb, rel) in SIDE_CONFIGS.items():
        if kind != KIND:
            side[str(kind)] = dict(cpu_legs(kind, min(n * rel, CPU_SAMPLE_VALUES), threads, args.cpu_seconds), label=label)
    sample = "the whole 2^%.2f-value column, all vectors, decode = falp + patch_exceptions" % np.log2(max(n, 1))
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
        "config": workload_config(n),
        "verified_bit_exact": bool(verified),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": head["kind"], "sample": sample, "build": head["build"], "runner": head["runner"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "configs": side,
    }
    channel.emit(line)


# ------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------------
def cuda_ms(fn, reps):
    """Average device time of fn() over `reps` launches (CUDA events on torch's current stream, which is the stream
    the library launches on)."""
    import torch

    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def compress_on_device(x, dev, reps=3):
    """Row-group init + encode of a device column, device-timed -> (trimmed DeviceColumn, timings dict)."""
    import torch

    import alp_b200

    vb = x.element_size()
    n_vec = x.numel() // 1024
    big = alp_b200.DeviceColumn(n_vec, vb, dev)  # worst-case capacities; allocated before anything is timed
    ws = torch.empty(max(256, alp_b200.lib.alpb200_encode_workspace_bytes(n_vec)), dtype=torch.uint8, device=dev)
    states = alp_b200.rowgroup_init(x)
    alp_b200.encode(x, states, col=big, workspace=ws)  # warm-up (also first touch of the output buffers)
    alp_b200.encode(x, states, col=big, workspace=ws, ordered=False)
    n_rg = states.shape[0]
    init_ws = torch.empty(max(256, alp_b200.lib.alpb200_init_workspace_bytes(x.numel())), dtype=torch.uint8, device=dev)
    t = {}
    t["encode_unordered_ms"] = cuda_ms(lambda: alp_b200.encode(x, states, col=big, workspace=ws, ordered=False), reps)
    t["rowgroup_init_ms"] = cuda_ms(lambda: alp_b200.rowgroup_init(x, states=states, workspace=init_ws), reps)
    t["encode_ms"] = cuda_ms(lambda: alp_b200.encode(x, states, col=big, workspace=ws), reps)  # vector order: the column used below

    def both():
        alp_b200.rowgroup_init(x, states=states, workspace=init_ws)
        alp_b200.encode(x, states, col=big, workspace=ws)

    t["init_plus_encode_ms"] = cuda_ms(both, reps)
    t["row_groups"] = int(n_rg)
    packed_bytes, n_exc = big.read_totals()
    col = alp_b200.DeviceColumn(n_vec, vb, dev, max(packed_bytes, 128), max(n_exc, 1))  # trimmed to what was used
    col.meta.copy_(big.meta)
    col.packed[:packed_bytes].copy_(big.packed[:packed_bytes])
    col.exc_val[:n_exc].copy_(big.exc_val[:n_exc])
    col.exc_pos[:n_exc].copy_(big.exc_pos[:n_exc])
    col.totals.copy_(big.totals)
    col.max_block_bytes = big.max_block_bytes
    del big, ws, states, init_ws
    torch.cuda.empty_cache()
    return col, t


def config_block(kind, n, dev, rank, peak, steps, with_cpu, cpu_seconds):
    """One BASELINE config on this rank's GPU: device-timed encode / init / decode / scan with roofline fractions and a
    bit-exact check; at N = 1 the reference's CPU numbers for a bounded slice of the same column beside them."""
    import torch

    import alp_b200
    from alp_b200 import _abi

    label, vb, _ = SIDE_CONFIGS[kind]
    x = alp_b200.generate(n, kind, dev, first_index=rank * n)
    col, t = compress_on_device(x, dev)
    meta = col.meta.cpu().numpy().view(_abi.VEC_META_DTYPE).reshape(-1)
    read_bytes = algorithmic_read_bytes(meta, vb)
    algo = read_bytes + n * vb
    out = torch.empty_like(x)
    alp_b200.decode(col, out=out)
    ibits = torch.int64 if vb == 8 else torch.int32
    verified = bool(torch.equal(out.view(ibits), x.view(ibits)))
    assert verified, "config %d: decoded column differs from the original" % kind
    for _ in range(3):
        alp_b200.decode(col, out=out)
    dec_ms = cuda_ms(lambda: alp_b200.decode(col, out=out), steps)
    acc = torch.zeros(1, dtype=torch.float64, device=dev)
    for _ in range(3):
        alp_b200.decode_sum(col, out=acc)
    scan_ms = cuda_ms(lambda: alp_b200.decode_sum(col, out=acc), steps)
    acc.zero_()
    alp_b200.decode_sum(col, out=acc)
    want = float(x.sum(dtype=torch.float64).item())
    got = float(acc.item())
    mm = torch.empty(3, dtype=torch.float64, device=dev)
    for _ in range(3):
        alp_b200.decode_minmax(col, out=mm)
    mm_ms = cuda_ms(lambda: alp_b200.decode_minmax(col, out=mm), steps)
    mm_min, mm_max, mm_cnt = alp_b200.minmax_result(mm)
    mm_ok = mm_min == float(x.min().item()) and mm_max == float(x.max().item()) and mm_cnt == int((x == x).sum().item())
    assert mm_ok, "config %d: fused MIN / MAX / COUNT differs from torch's over the original column" % kind
    # predicate filter: `value < c` with c = the mean of MIN and MAX (selects a good share of every config's values)
    f_c = 0.5 * (mm_min + mm_max)
    f_bitmap = torch.empty(32 * (n // 1024), dtype=torch.int32, device=dev)
    f_sel = torch.zeros(1, dtype=torch.int64, device=dev)
    for _ in range(3):
        alp_b200.decode_filter(col, "<", f_c, bitmap=f_bitmap, selected=f_sel)
    f_ms = cuda_ms(lambda: alp_b200.decode_filter(col, "<", f_c, bitmap=f_bitmap, selected=f_sel), steps)
    f_want = int((x.double() < f_c).sum().item())
    f_ok = int(f_sel.item()) == f_want
    assert f_ok, "config %d: fused filter selected %d values, torch %d" % (kind, int(f_sel.item()), f_want)
    del f_bitmap
    dec_block = None
    if vb == 4:  # float columns: the decimal-sum variant (integers added exactly, one conversion per thread; include/alp_b200.h)
        for _ in range(3):
            alp_b200.decode_sum(col, out=acc, flags=alp_b200.SUM_DECIMAL)
        dscan_ms = cuda_ms(lambda: alp_b200.decode_sum(col, out=acc, flags=alp_b200.SUM_DECIMAL), steps)
        acc.zero_()
        alp_b200.decode_sum(col, out=acc, flags=alp_b200.SUM_DECIMAL)
        dgot = float(acc.item())
        sum_abs = float(x.abs().sum(dtype=torch.float64).item())
        dec_block = {"ms": dscan_ms, "read_GBps": read_bytes / (dscan_ms * 1e-3) / 1e9, "roofline_frac": read_bytes / (dscan_ms * 1e-3) / 1e9 / peak,
                     "GBps_decoded_equivalent": n * vb / (dscan_ms * 1e-3) / 1e9,
                     "abs_diff_over_sum_abs": abs(dgot - want) / max(sum_abs, 1e-300), "bound": 2.0 ** -23,
                     "semantics": "ALPB200_SUM_DECIMAL: sum of the decimals the floats stand for (|diff| <= 2^-23 * sum|x|)"}
        assert dec_block["abs_diff_over_sum_abs"] <= 2.0 ** -23, "decimal SUM outside its stated bound"

    def rate(ms, nbytes):
        return nbytes / (ms * 1e-3) / 1e9

    block = {
        "label": label,
        "values": int(n),
        "dtype": "f64" if vb == 8 else "f32",
        "schemes": sorted({2: "ALP", 1: "ALP_RD"}[int(s)] for s in set(meta["scheme"].tolist())),
        "bits_per_value": 8.0 * read_bytes / n,
        "exceptions_per_vector": float(meta["exc_cnt"].astype(np.int64).sum()) / max(1, meta.shape[0]),
        "algorithmic_bytes_per_launch": int(algo),
        "verified_bit_exact": verified,
        "decode": {"ms": dec_ms, "GBps": rate(dec_ms, n * vb), "roofline_frac": rate(dec_ms, algo) / peak},
        "encode": {"ms": t["encode_ms"], "GBps": rate(t["encode_ms"], n * vb), "roofline_frac": rate(t["encode_ms"], algo) / peak, "layout": "vector order"},
        "encode_unordered": {"ms": t["encode_unordered_ms"], "GBps": rate(t["encode_unordered_ms"], n * vb),
                             "roofline_frac": rate(t["encode_unordered_ms"], algo) / peak, "layout": "completion order"},
        "rowgroup_init_ms": t["rowgroup_init_ms"],
        "init_plus_encode": {"ms": t["init_plus_encode_ms"], "GBps": rate(t["init_plus_encode_ms"], n * vb),
                             "roofline_frac": rate(t["init_plus_encode_ms"], algo) / peak},
        "scan_sum": {"ms": scan_ms, "GBps_decoded_equivalent": rate(scan_ms, n * vb), "read_GBps": rate(scan_ms, read_bytes),
                     "roofline_frac": rate(scan_ms, read_bytes) / peak, "rel_err_vs_torch_sum": abs(got - want) / max(abs(want), 1e-300)},
    }
    block["scan_minmax"] = {"ms": mm_ms, "GBps_decoded_equivalent": rate(mm_ms, n * vb), "read_GBps": rate(mm_ms, read_bytes),
                            "roofline_frac": rate(mm_ms, read_bytes) / peak, "matches_torch_min_max_count": bool(mm_ok),
                            "api": "alpb200_decode_minmax_* (decode + patch in shared memory, reduce; nothing written back)"}
    block["scan_filter"] = {"ms": f_ms, "GBps_decoded_equivalent": rate(f_ms, n * vb), "read_GBps": rate(f_ms, read_bytes),
                            "roofline_frac": rate(f_ms, read_bytes + n // 8) / peak, "selected_fraction": f_want / n,
                            "count_matches_torch": bool(f_ok),
                            "api": "alpb200_decode_filter_* (value < constant -> 1 bit per value; decode + patch in shared memory)"}
    if dec_block is not None:
        block["scan_sum_decimal"] = dec_block
    if with_cpu:
        ns = min(n, CPU_SAMPLE_VALUES)
        block["cpu"] = cpu_legs(kind, ns, host_threads(), cpu_seconds, x_host=None if ref_bench_binary() else x[:ns].cpu().numpy())
        c = block["cpu"]
        block["vs_cpu"] = {
            "decode": block["decode"]["GBps"] / c["decode_GBps"],
            "encode": block["encode"]["GBps"] / c["encode_GBps"],
            "init_plus_encode": block["init_plus_encode"]["GBps"] / c["init_plus_encode_GBps"],
            "scan_sum": block["scan_sum"]["GBps_decoded_equivalent"] / c["scan_sum_GBps"],
        }
    return block, x, col, meta, read_bytes


def link_probe(dev, world, barrier, nbytes=1 << 30, reps=3):
    """Pinned host <-> device copy bandwidth of this rank's PCIe link, measured with every rank copying at the same
    time (the e2e numbers are bounded by it): each direction alone, then both directions at once on two streams."""
    import torch

    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.zero_()
    h_out.zero_()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d, d2h):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return nbytes * reps / float(t.item()) / 1e9  # per rank, at the pace of the slowest rank

    run(True, True)  # warm-up
    res = {"h2d_GBps": run(True, False), "d2h_GBps": run(False, True)}
    both = run(True, True)
    res["duplex_GBps_each_way"] = both
    res["ranks_copying_at_once"] = world
    res["bytes_per_copy"] = nbytes
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="alp_b200", choices=["alp_b200", "reference"])
    ap.add_argument("--values", type=int, default=int(os.environ.get("ALPB200_BENCH_VALUES", DEFAULT_VALUES)), help="values per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="budget of the CPU legs, per config")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the config 3 / 4 blocks")
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
    peak, peak_src = measured_peak()
    with_cpu = rank == 0 and world == 1  # the CPU legs run beside the GPU numbers at N = 1 only

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the side configs first (their buffers are gone before the headline column is built) ----
    configs = {}
    if not args.no_side:
        for kind in (3, 4):
            nk = int(n * SIDE_CONFIGS[kind][2]) // 1024 * 1024
            block, xk, colk, _, _ = config_block(kind, nk, dev, rank, peak, args.steps, with_cpu, args.cpu_seconds)
            configs[str(kind)] = block
            del xk, colk
            torch.cuda.empty_cache()

    # ---- untimed set-up of the headline: this rank's shard of the global column, compressed on the device ----
    block2, x, col, meta, read_bytes = config_block(KIND, n, dev, rank, peak, args.steps, with_cpu, args.cpu_seconds)
    configs[str(KIND)] = block2
    verified = block2["verified_bit_exact"]
    algo_bytes = read_bytes + n * 8
    out = torch.empty(n, dtype=torch.float64, device=dev)

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

    def max_over_ranks(seconds):
        ts = torch.tensor([seconds], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        return float(ts.item())

    # ---- e2e: host column in, host values out, through alpb200_decompress_host ----
    e2e = e2e_scan = e2e_compress = None
    if not args.no_e2e:
        link = link_probe(dev, world, barrier)
        hcol = pinned_host_column(col.to_host())
        hout_t = torch.empty(n, dtype=torch.float64, pin_memory=True)
        hout = hout_t.numpy()
        codec = alp_b200.HostCodec(n_vec, 8, local)
        codec.decompress(hcol, out=hout)  # warm-up + check of the whole result (compared on the device)
        out.copy_(hout_t, non_blocking=True)
        e2e_ok = bool(torch.equal(out.view(torch.int64), x.view(torch.int64)))
        assert e2e_ok, "alpb200_decompress_host: result differs from the original"
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            codec.decompress(hcol, out=hout)
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        h2d = hcol.meta.nbytes + hcol.packed_bytes + hcol.n_exceptions * 10
        d2h = n * 8

        e2e = {
            "value": world * n * 8.0 / e2e_s / 1e9,
            "unit": UNIT,
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h),
            "ms_per_step": e2e_s * 1e3,
            "steps": args.e2e_steps,
            "verified_bit_exact": e2e_ok,
            "api": "alpb200_decompress_host_f64 (pinned host buffers, chunk pipeline over 3 streams)",
            "link": link,
            "link_bound_ms": max(h2d / link["h2d_GBps"], d2h / link["d2h_GBps"]) / 1e6,
            "link_frac": max(h2d / link["h2d_GBps"], d2h / link["d2h_GBps"]) / 1e9 / e2e_s,
            "bound": "PCIe D2H of the decoded values (8 B/value); link_frac = time the slower copy alone needs at the measured pinned-copy rate / measured time",
        }
        # the scan query end to end: host column in, one double out (only the compressed bytes cross PCIe)
        want_sum = float(x.sum().item())
        got_sum = codec.sum(hcol)
        assert abs(got_sum - want_sum) <= 1e-9 * abs(want_sum), (got_sum, want_sum)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            codec.sum(hcol)
        scan_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e_scan = {
            "value": world * n * 8.0 / scan_s / 1e9,
            "unit": "GB/s of decoded-equivalent f64 scanned",
            "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 8,
            "ms_per_step": scan_s * 1e3,
            "api": "alpb200_sum_host_f64 (SUM over the host column; reference: bench_end_to_end alp_func + aggr_plus)",
            "link_frac": h2d / link["h2d_GBps"] / 1e9 / scan_s,
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
        comp_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e_compress = {
            "value": world * n * 8.0 / comp_s / 1e9,
            "unit": "GB/s of f64 input compressed",
            "h2d_bytes_per_step": int(n * 8),
            "d2h_bytes_per_step": int(h2d),
            "ms_per_step": comp_s * 1e3,
            "api": "alpb200_compress_host_f64 (pinned host buffers; row-group init + encode per chunk, chunk pipeline over 3 streams)",
            "link_frac": n * 8 / link["h2d_GBps"] / 1e9 / comp_s,
        }
        if "cpu" in block2:
            e2e_scan["cpu_reference_GBps"] = block2["cpu"]["scan_sum_GBps"]
            e2e_scan["vs_cpu"] = e2e_scan["value"] / block2["cpu"]["scan_sum_GBps"]
            e2e_compress["cpu_reference_GBps"] = block2["cpu"]["init_plus_encode_GBps"]
            e2e_compress["vs_cpu"] = e2e_compress["value"] / block2["cpu"]["init_plus_encode_GBps"]
            e2e["cpu_reference_GBps"] = block2["cpu"]["decode_GBps"]
        codec.close()
        del hcol, hout, hout_t

    # ---- N > 1: rank 0 hands ITS compressed column out as row-group shards over NCCL (BASELINE config 5 says
    # "scattered"); a set-up step, outside the timed decode, reported on its own ----
    scatter = None
    if world > 1:
        from alp_b200 import shard

        tensors = shard.column_tensors(col) if rank == 0 else None
        # the first scatter also opens NCCL's point-to-point connections (hundreds of ms): it is reported as `first_ms`,
        # the steady-state figure is the second one
        barrier()
        t0 = time.perf_counter()
        mine, (first, count) = shard.scatter_column(tensors, src=0, value_bytes=8, device=dev)
        torch.cuda.synchronize()
        first_s = max_over_ranks(time.perf_counter() - t0)
        del mine
        barrier()
        t0 = time.perf_counter()
        mine, (first, count) = shard.scatter_column(tensors, src=0, value_bytes=8, device=dev)
        torch.cuda.synchronize()
        sc_s = max_over_ranks(time.perf_counter() - t0)
        shard_col = shard.tensors_to_device_column(mine, 8, dev)
        got = alp_b200.decode(shard_col)
        want = alp_b200.generate(count * 1024, KIND, dev, first_index=first * 1024)  # rank 0's column starts at value 0
        ok = torch.tensor([int(torch.equal(got.view(torch.int64), want.view(torch.int64)))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        sent = read_bytes * (world - 1) / world if rank == 0 else 0
        scatter = {"ms": sc_s * 1e3, "first_ms": first_s * 1e3, "bytes_sent_by_rank0": int(sent), "GBps": (sent / sc_s / 1e9) if rank == 0 else None,
                   "shards_decode_bit_exact": bool(ok.item()), "api": "alp_b200.shard.scatter_column (NCCL point-to-point, whole row-groups)"}
        del shard_col, got, want, mine

    if rank == 0:
        achieved = algo_bytes / (ms_local * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get("decode_f64_dram_bytes_per_launch")
        except Exception:
            pass
        cpu = None
        if "cpu" in block2:
            c = block2["cpu"]
            cpu = {"value": c["decode_GBps"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "single_thread_value": c["decode_single_thread_GBps"],
                   "sample": c["sample"] + ", decoded %d times (falp + patch_exceptions; %s)" % (c["decode_reps"], c["build"]), "runner": c["runner"]}
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
            "config": workload_config(n),
            "verified_bit_exact": verified,
            "parallelism": "row-group shards, %d rank(s), no data-path collective" % world,
            "bits_per_value": 8.0 * read_bytes / n,
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
            "configs": configs,
            "scatter": scatter,
        }
        channel.emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
