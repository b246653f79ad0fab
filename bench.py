#!/usr/bin/env python
"""bench.py — pileup+call throughput of the B200 engine on the synthetic shapes of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                    (the reference algorithm's CPU path: the oracle, all host threads)

A "step" is one pass of the hot path (K_tile + K_exact + record download) over one batch of synthetic reads.  At N = 1 the
workload is BASELINE.json configs[1]: germline-threshold, chr20 shape (63,025,520 loci), 30x, 150 bp.  At N > 1 every rank
holds one chr20-shaped contig of an N-contig genome (LociPartitioning assigns contiguous contig ranges to ranks; no
data-path collective) and rank 0 gathers the variant records over NCCL — weak scaling.

`value` is measured with the packed reads already resident in HBM; `e2e` goes through guac_reads_pack +
guac_germline_threshold from (pinned) HOST buffers every step, results copied back.  `cpu_baseline` is the oracle (a C++
restatement of the reference's Scala algorithm — there is no JVM on the box) on a bounded window of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from guacamole_b200 import abi, synth  # noqa: E402

CHR20 = synth.CHR20_LENGTH
READ_LEN = 150


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="germline", choices=["germline", "somatic"])
    ap.add_argument("--contig-length", type=int, default=CHR20)
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--cpu-window", type=int, default=4_000_000, help="loci of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--seed", type=int, default=20261020)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in a background thread every 5 ms (the timed
    region of the resident leg lasts milliseconds), nvidia-smi as the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self.thread = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample(self):
        n = self.nvml
        try:
            self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            self.reason_bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.005)

    def start(self):
        if self.nvml:
            self._sample()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.nvml:
            return self._smi_once()
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=1)
        self._sample()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.reason_bits & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            f = [x.strip() for x in out]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi (after the timed region)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}


def measured_traffic(kernel, n_loci, depth):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r1_traffic.json), scaled
    linearly in loci when the capture was taken on a slice; None if there is no capture for this shape."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(kernel)
    if not t or abs(t["depth"] - depth) > 1e-9:
        return None
    return t["traffic_bytes"] * (n_loci / t["workload_loci"])


def algorithmic_bytes(view, n_loci):
    """per read start 4 + ref_len 4 + 4*c + ceil(L/4) + 1 flag bytes; per locus 0.25 B of reference."""
    n = int(view.n_reads)
    n_ops = int(np.ctypeslib.as_array(view.cigar_off, shape=(n + 1,))[n])
    bases = int(np.ctypeslib.as_array(view.seq_off, shape=(n + 1,))[n])
    return n * 9 + 4 * n_ops + (bases + 3) // 4 + 0.25 * n_loci


def pinned_copy(view):
    """Copies the generator's columns into page-locked host memory (torch pinned tensors) and returns a guac_read_batch
    over them, so that the e2e leg's host->device copies start from pinned memory."""
    import torch
    n = int(view.n_reads)
    so = np.ctypeslib.as_array(view.seq_off, shape=(n + 1,))
    co = np.ctypeslib.as_array(view.cigar_off, shape=(n + 1,))
    mo = np.ctypeslib.as_array(view.md_off, shape=(n + 1,))
    keep = []

    def col(ptr, count, dtype, ctype):
        count = max(int(count), 1)
        src = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,))
        t = torch.empty(count, dtype=dtype, pin_memory=True)
        t.numpy()[:] = src.view(t.numpy().dtype)
        keep.append(t)
        return C.cast(t.data_ptr(), C.POINTER(ctype))

    b = abi.ReadBatchC()
    b.n_reads, b.n_contigs = view.n_reads, view.n_contigs
    b.contig_length = col(view.contig_length, view.n_contigs, torch.int64, C.c_int64)
    b.contig = col(view.contig, n, torch.int32, C.c_int32)
    b.start = col(view.start, n, torch.int64, C.c_int64)
    b.cigar_off = C.cast(col(view.cigar_off, n + 1, torch.int64, C.c_int64), C.POINTER(C.c_uint64))
    b.cigar = C.cast(col(view.cigar, co[n], torch.int32, C.c_int32), C.POINTER(C.c_uint32))
    b.seq_off = C.cast(col(view.seq_off, n + 1, torch.int64, C.c_int64), C.POINTER(C.c_uint64))
    b.seq = col(view.seq, so[n], torch.uint8, C.c_uint8)
    b.qual = col(view.qual, so[n], torch.uint8, C.c_uint8)
    b.mapq = col(view.mapq, n, torch.uint8, C.c_uint8)
    b.flags = col(view.flags, n, torch.uint8, C.c_uint8)
    b.sample = col(view.sample, n, torch.int32, C.c_int32)
    b.md_off = C.cast(col(view.md_off, n + 1, torch.int64, C.c_int64), C.POINTER(C.c_uint64))
    # c_char_p fields convert to Python bytes on access: take the raw pointer value from the struct instead
    md_addr = C.c_void_p.from_address(C.addressof(view) + abi.ReadBatchC.md.offset).value
    md_pinned = col(C.cast(md_addr, C.POINTER(C.c_uint8)), mo[n], torch.uint8, C.c_uint8)
    C.c_void_p.from_address(C.addressof(b) + abi.ReadBatchC.md.offset).value = C.cast(md_pinned, C.c_void_p).value
    return b, keep


def cpu_leg(args, steps, warmup, n_threads):
    """The reference algorithm on host cores: oracle over a bounded window of the same workload, Spark-style loci tasks."""
    import oracle_binding as orc
    somatic = args.workload == "somatic"
    window = min(args.cpu_window // (3 if somatic else 1), args.contig_length)
    samples = [(1, args.depth * 2), (0, args.depth)] if somatic else [(0, args.depth)]
    sbs = [synth.generate([("20", args.contig_length)], depth=d, read_length=READ_LEN, seed=args.seed, sample=s,
                          window=(0, 0, window)) for s, d in samples]
    ranges = orc.partition_loci_uniformly(n_threads, [(0, 0, window - READ_LEN - 40)])
    arr = orc.ranges_array(ranges)
    L = orc.lib()
    n_loci = sum(r[2] - r[1] for r in ranges)
    times = []
    n_rec = 0
    for i in range(warmup + steps):
        h = C.c_void_p()
        t0 = time.perf_counter()
        if somatic:
            prm = orc.somatic_params(odds=20, min_mapq=1)
            rc = L.orc_somatic_standard(C.byref(sbs[0].c), C.byref(sbs[1].c), None, arr, C.c_size_t(len(ranges)), C.byref(prm),
                                        n_threads, C.byref(h))
        else:
            prm = orc.threshold_params(8)
            rc = L.orc_germline_threshold(C.byref(sbs[0].c), None, arr, C.c_size_t(len(ranges)), C.byref(prm), n_threads, C.byref(h))
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError("oracle failed: " + L.orc_last_error().decode())
        n_rec = L.orc_result_n(h)
        L.orc_result_free(h)
        if i >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    reads = sum(sb.n_reads for sb in sbs)
    return {"loci_per_s": n_loci / sec, "reads_per_s": reads / sec, "sec_per_step": sec, "loci": n_loci,
            "reads": reads, "records": int(n_rec), "threads": n_threads, "window": window}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_threads = os.cpu_count() or 1
    if args.workload == "somatic":
        wl = (f"somatic-standard, synthetic tumor/normal pair {2 * args.depth:g}x/{args.depth:g}x, chr20 shape "
              f"({args.contig_length:,} loci), {READ_LEN} bp (BASELINE.json configs[2]) per GPU")
    else:
        wl = (f"germline-threshold, synthetic chr20 shape ({args.contig_length:,} loci), {args.depth:g}x, {READ_LEN} bp "
              f"(BASELINE.json configs[1]) per GPU")
    config = {"workload": wl, "threshold_percent": 8, "loci_per_gpu": args.contig_length, "depth": args.depth, "read_length": READ_LEN,
              "parallelism": f"loci-partitioned x{world}", "l2": "inputs larger than L2 (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return
        cpu = cpu_leg(args, max(1, args.steps), max(0, min(args.warmup, 1)), n_threads)
        line = {"impl": "reference", "metric": "loci_per_sec", "value": cpu["loci_per_s"], "unit": "loci/s",
                "reads_per_sec": cpu["reads_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": cpu["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64" if args.workload == "somatic" else "u8", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": cpu["loci_per_s"], "unit": "loci/s", "cores": cpu["threads"], "kind": "port",
                                 "sample": f"first {cpu['window']:,} loci of the chr20-shape workload ({cpu['reads']:,} reads); "
                                           "C++ oracle restating the reference's Scala algorithm (no JVM on the box)"},
                "e2e": {"value": cpu["loci_per_s"], "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from guacamole_b200 import callers
    from guacamole_b200._lib import lib
    from guacamole_b200.distributed import gather_records, ranges_of_rank
    from guacamole_b200.loci import partition_loci_uniformly

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    L = lib()
    somatic = args.workload == "somatic"

    # ---- this rank's shard: one chr20-shaped contig of a `world`-contig genome (LociPartitioning over contigs)
    contigs = [(f"20_{r}" if world > 1 else "20", args.contig_length) for r in range(world)]
    names = [c[0] for c in contigs]
    loci_all = [(c, 0, args.contig_length - 1) for c in range(world)]  # LociSet "all" drops the last base of a contig
    my_ranges = ranges_of_rank(partition_loci_uniformly(world, loci_all), rank)
    n_loci = sum(r[2] - r[1] for r in my_ranges)
    t_gen = time.perf_counter()
    samples = [(1, args.depth * 2), (0, args.depth)] if somatic else [(0, args.depth)]   # tumor 60x + normal 30x
    views, keeps, n_reads, algorithmic = [], [], 0, 0.25 * n_loci
    for sample, depth in samples:
        sb = synth.generate(contigs, depth=depth, read_length=READ_LEN, seed=args.seed, sample=sample,
                            window=(rank, 0, args.contig_length))
        n_reads += sb.n_reads
        algorithmic += algorithmic_bytes(sb.c, 0) + ((READ_LEN + 3) * sb.n_reads if somatic else 0)
        v, k = pinned_copy(sb.c)
        sb.free()
        views.append(v)
        keeps.append(k)
    gen_s = time.perf_counter() - t_gen

    ctx = callers.Context(local_rank)
    ctx.set_option(abi.OPT_PACK_QUALITIES, 1 if somatic else 0)
    ctx.set_option(abi.OPT_HOST_THREADS, max(1, n_threads // world))  # ranks share the box's host cores
    packed = [ctx.pack_c(v, names) for v in views]
    pack_ms = sum(p.pack_kernel_ms for p in packed)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def call(reads_list):
        if somatic:
            return callers.somatic_standard(ctx, reads_list[0], reads_list[1], my_ranges, odds_threshold=20, min_alignment_quality=1)
        return callers.germline_threshold(ctx, reads_list[0], my_ranges, threshold=8)

    # ---- device-resident leg
    ctx.set_option(abi.OPT_SORT_RECORDS, 0)
    for _ in range(args.warmup):
        res = call(packed)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.timer_start()
    t0 = time.perf_counter()
    tile_ms, exact_ms, launches = 0.0, 0.0, 0
    for _ in range(args.steps):
        res = call(packed)
        tile_ms += res.stats["tile_kernel_ms"]
        exact_ms += res.stats["exact_kernel_ms"]
        launches += res.stats["kernel_launches"]
    dev_ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    n_records = len(res)
    step_ms = max(dev_ms, 0.0) / args.steps

    # ---- end-to-end leg: host buffers -> pack -> call -> records on the host, every step
    ctx.set_option(abi.OPT_SORT_RECORDS, 1)
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    h2d = d2h = 0
    for i in range(1 + e2e_steps):
        if i == 1:
            barrier()
            t1 = time.perf_counter()
        fresh = [ctx.pack_c(v, names) for v in views]
        out = call(fresh)
        h2d = sum(int(L.guac_reads_h2d_bytes(f._h)) for f in fresh) + int(out.stats["h2d_bytes"])
        d2h = int(out.stats["d2h_bytes"])
        for f in fresh:
            f.free()
    barrier()
    e2e_ms = (time.perf_counter() - t1) * 1e3 / e2e_steps
    assert len(out) == n_records, "e2e and resident legs disagree"

    # ---- gather per-shard records to rank 0 (the path's only exchange), max-over-ranks timing
    t = torch.tensor([step_ms, e2e_ms, float(n_records), tile_ms / args.steps, exact_ms / args.steps], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        t_g = time.perf_counter()
        allrec, _ = gather_records(out.records, out.bytes, dst=0, device=torch.device("cuda", local_rank))
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - t_g) * 1e3
        step_ms, e2e_ms = float(mx[0]), float(mx[1])
        total_records = len(allrec) if rank == 0 else 0
        tile_step_ms, exact_step_ms = float(mx[3]), float(mx[4])
    else:
        total_records = n_records
        gather_ms = 0.0
        tile_step_ms, exact_step_ms = tile_ms / args.steps, exact_ms / args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    alg_bytes = algorithmic
    total_loci = n_loci * world
    total_reads = n_reads * world
    value = total_loci / (step_ms * 1e-3)
    e2e_value = total_loci / (e2e_ms * 1e-3)
    achieved = alg_bytes / (tile_step_ms * 1e-3) / 1e9
    line = {
        "metric": "loci_per_sec", "value": value, "unit": "loci/s", "reads_per_sec": total_reads / (step_ms * 1e-3),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if somatic else "u8", "data": "synthetic", "config": config,
        "records_per_step": total_records, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "loci/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "pinned_host_buffers": sum(len(k) for k in keeps)},
        "roofline": {"bound": "hbm", "kernel": "k_somatic" if somatic else "k_pileup_tile", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic("k_somatic" if somatic else "k_pileup_tile", n_loci, args.depth), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tile_step_ms, "exact_kernel_ms": exact_step_ms,
                     "whole_step_achieved_gbs": alg_bytes / (step_ms * 1e-3) / 1e9},
        "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps, "pack_kernel_ms": pack_ms, "generate_s": gen_s,
        "gather_ms": gather_ms,
    }
    if world == 1:
        cpu = cpu_leg(args, 1, 0, n_threads)
        line["cpu_baseline"] = {"value": cpu["loci_per_s"], "unit": "loci/s", "cores": cpu["threads"], "kind": "port",
                                "sample": f"first {cpu['window']:,} loci of the same workload ({cpu['reads']:,} reads, "
                                          f"{cpu['sec_per_step']:.1f} s); C++ oracle restating the reference's Scala algorithm"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
