#!/usr/bin/env python
"""bench.py — pileup+call throughput of the B200 engine on the synthetic shapes of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                    (the reference algorithm's CPU path: the oracle, all host threads)
    python bench.py --workload amplicon --gpus 8            (BASELINE.json configs[4]: somatic-standard, 10,000x over 1 Mb)

A "step" is one pass of the hot path over one batch of synthetic reads, through the C ABI of include/guac.h.
  N = 1   configs[1]: germline-threshold, chr20 shape (63,025,520 loci), 30x, 150 bp — plus a `somatic` block in the same JSON
          line for configs[2] (somatic-standard, tumor 60x / normal 30x over the same contig).
  N > 1   configs[3]: germline-threshold on the whole-genome shape (GRCh37 contig lengths of the reference's own
          DistributedUtilSuite.scala:72, 3.1 G loci, 30x): partitionLociUniformly(N) cuts the loci into contiguous ranges (mid
          contig, several contigs per rank), every rank generates the reads overlapping its ranges ON ITS DEVICE from the
          seed (boundary reads exist on both neighbours, DistributedUtil.scala:585-597), packs and calls them; the records and
          the depth histogram are gathered to rank 0 by the library's NCCL exchange.  Total work is fixed: strong scaling.
`value` is measured with the packed reads resident in HBM; `e2e` goes from (pinned) HOST buffers through guac_reads_pack +
the call (+ the NCCL gather at N > 1) to records in host memory every step — at N > 1 over a chr20-sized slice of each rank's
shard (the whole shard would need 124 GB of pinned host memory).  `cpu_baseline` / `--impl reference` is the oracle (a C++
restatement of the reference's Scala algorithm — there is no JVM on the box) on a bounded window of the same workload, and
`parity_window` says whether the engine's records inside that window equal the oracle's; `e2e_check` whether the end-to-end
leg handed back exactly the resident leg's records (an order-independent fingerprint of all of them); `gather_check` (N > 1)
whether rank 0's gathered set is the concatenation of the ranks' sets.  Any of them failing exits non-zero after the line.
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
SPAN = READ_LEN + 40          # longest reference span of a generated read


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "germline", "somatic", "amplicon"])
    ap.add_argument("--contig-length", type=int, default=CHR20, help="N = 1: length of the one contig")
    ap.add_argument("--depth", type=float, default=30.0)
    ap.add_argument("--cpu-window", type=int, default=4_000_000, help="loci of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--e2e-wide", dest="e2e_compact", action="store_false",
                    help="end-to-end leg from the wide guac_read_batch (ASCII bases, 64-bit columns) instead of guac_read_batch_v2")
    ap.add_argument("--no-pack-overlap", action="store_true", help="GUAC_OPT_PACK_OVERLAP = 0 (A/B of the chunk-wise pack finish)")
    ap.add_argument("--no-somatic", action="store_true", help="N = 1: leave the configs[2] block out")
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
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r2_traffic.json), scaled
    linearly in loci; None if there is no capture for this shape.  (ncu cannot run inside the timed bench.)"""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(kernel)
    if not t or abs(t["depth"] - depth) > 1e-9:
        return None
    return t["traffic_bytes"] * (n_loci / t["workload_loci"])


def genome(world, workload, contig_length):
    """(contigs, loci in LociSet order = contigs sorted by name, LociMap.contigs) of the workload."""
    if workload == "amplicon":
        contigs = [("amplicon", 1_000_000 + SPAN)]
    elif world == 1:
        contigs = [("20", contig_length)]
    else:  # the sequence dictionary in name order, so that LociSet order (contigs by name) and contig index order agree
        contigs = sorted(synth.GRCH37, key=lambda c: c[0])
    order = sorted(range(len(contigs)), key=lambda i: contigs[i][0])
    if workload == "amplicon":
        loci = [(0, 0, 1_000_000)]
    else:
        loci = [(i, 0, contigs[i][1] - 1) for i in order]   # LociSet "all" drops the last base of a contig
    return contigs, loci


def algorithmic_bytes_of(n_reads, n_ops, n_bases, with_qualities):
    """per read start 4 + ref_len 4 + 4*c + ceil(L/4) + 1 flag bytes (+ L qualities + mapq + mismatch count for
    the likelihood callers); the 0.25 B of reference per locus is added by the caller."""
    b = n_reads * 9 + 4 * n_ops + (n_bases + 3) // 4
    if with_qualities:
        b += n_bases + 3 * n_reads
    return b


def germline_keys(res, contig=None, lo=None, hi=None):
    """(contig, start, ref, alt, gt) of the germline records of `res` inside [lo, hi) of `contig`, sorted."""
    recs, pool = res.records, res.bytes
    if contig is not None:
        recs = recs[(recs["contig"] == contig) & (recs["start"] >= lo) & (recs["start"] < hi)]
    return sorted((int(r["contig"]), int(r["start"]), pool[int(r["ref_off"]):int(r["ref_off"]) + int(r["ref_len"])],
                   pool[int(r["alt_off"]):int(r["alt_off"]) + int(r["alt_len"])], int(r["gt"][0]), int(r["gt"][1])) for r in recs)


def record_digest(res, somatic):
    """Order-independent fingerprint of a whole result (every record's locus, alleles and genotype / depths): the end-to-end leg
    must hand back exactly what the device-resident leg did."""
    import hashlib
    recs, pool = res.records, np.frombuffer(res.bytes, dtype=np.uint8) if not isinstance(res.bytes, np.ndarray) else res.bytes
    if len(recs) == 0:
        return "empty"
    cols = [recs["contig"].astype(np.int64), recs["start"].astype(np.int64), recs["ref_len"].astype(np.int64), recs["alt_len"].astype(np.int64)]
    # first and last allele bytes stand in for the strings (the lengths are in the key as well)
    for off, ln in (("ref_off", "ref_len"), ("alt_off", "alt_len")):
        o, l = recs[off].astype(np.int64), recs[ln].astype(np.int64)
        cols.append(np.where(l > 0, pool[np.minimum(o, len(pool) - 1)], 0).astype(np.int64))
        cols.append(np.where(l > 0, pool[np.minimum(o + np.maximum(l, 1) - 1, len(pool) - 1)], 0).astype(np.int64))
    if somatic:
        cols += [recs["tumor"]["allele_read_depth"].astype(np.int64), recs["tumor"]["read_depth"].astype(np.int64),
                 recs["normal"]["read_depth"].astype(np.int64)]
    else:
        cols += [recs["gt"][:, 0].astype(np.int64), recs["gt"][:, 1].astype(np.int64)]
    m = np.stack(cols, axis=1)
    m = m[np.lexsort(m.T[::-1])]
    return f"{len(recs)}:" + hashlib.sha256(np.ascontiguousarray(m).tobytes()).hexdigest()[:16]


def somatic_keys(res, contig, lo, hi):
    recs, pool = res.records, res.bytes
    recs = recs[(recs["contig"] == contig) & (recs["start"] >= lo) & (recs["start"] < hi)]
    return sorted((int(r["contig"]), int(r["start"]), pool[int(r["ref_off"]):int(r["ref_off"]) + int(r["ref_len"])],
                   pool[int(r["alt_off"]):int(r["alt_off"]) + int(r["alt_len"])], int(r["tumor"]["allele_read_depth"]),
                   int(r["tumor"]["read_depth"]), int(r["normal"]["read_depth"])) for r in recs)


def cpu_leg(args, contigs, window, somatic, depths, steps, warmup, n_threads, keep_records=False):
    """The reference algorithm on host cores: the oracle over the reads STARTING in `window` = (contig, start, end) of the same
    synthetic workload (host build of the same generator), Spark-style loci tasks.  The loci [start + SPAN, end - SPAN) see
    every read that overlaps them."""
    import oracle_binding as orc
    c, s, e = window
    sbs = [synth.generate(contigs, depth=d, read_length=READ_LEN, seed=args.seed, sample=sm, window=(c, s, e)) for sm, d in depths]
    lo, hi = s + (SPAN if s > 0 else 0), e - SPAN
    ranges = orc.partition_loci_uniformly(n_threads, [(c, lo, hi)])
    arr = orc.ranges_array(ranges)
    L = orc.lib()
    times, n_rec, keys = [], 0, None
    for i in range(warmup + steps):  # straight from the generator's buffers (no numpy copy of the columns)
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
        n_rec = int(L.orc_result_n(h))
        if keep_records and keys is None:
            keys = oracle_keys(L, h, n_rec, somatic)
        L.orc_result_free(h)
        if i >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    reads = sum(sb.n_reads for sb in sbs)
    return {"loci_per_s": (hi - lo) / sec, "reads_per_s": reads / sec, "sec_per_step": sec, "loci": hi - lo, "reads": reads,
            "records": n_rec, "threads": n_threads, "window": (c, lo, hi), "keys": keys}


def oracle_keys(L, h, n_rec, somatic):
    nb = C.c_size_t()
    L.orc_result_bytes.restype = C.POINTER(C.c_uint8)
    L.orc_result_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    bp = L.orc_result_bytes(h, C.byref(nb))
    pool = bytes((C.c_uint8 * nb.value).from_address(C.cast(bp, C.c_void_p).value)) if nb.value else b""
    if somatic:
        L.orc_result_somatic_records.restype = C.POINTER(abi.SomaticRecordC)
        L.orc_result_somatic_records.argtypes = [C.c_void_p]
        p = L.orc_result_somatic_records(h)
        return sorted((p[k].contig, p[k].start, pool[p[k].ref_off:p[k].ref_off + p[k].ref_len],
                       pool[p[k].alt_off:p[k].alt_off + p[k].alt_len], p[k].tumor.allele_read_depth, p[k].tumor.read_depth,
                       p[k].normal.read_depth) for k in range(n_rec))
    L.orc_result_threshold_records.restype = C.POINTER(abi.ThresholdRecordC)
    L.orc_result_threshold_records.argtypes = [C.c_void_p]
    p = L.orc_result_threshold_records(h)
    return sorted((p[k].contig, p[k].start, pool[p[k].ref_off:p[k].ref_off + p[k].ref_len],
                   pool[p[k].alt_off:p[k].alt_off + p[k].alt_len], p[k].gt[0], p[k].gt[1]) for k in range(n_rec))


def cpu_baseline_block(cpu, what):
    return {"value": cpu["loci_per_s"], "unit": "loci/s", "cores": cpu["threads"], "kind": "port",
            "sample": f"loci {cpu['window'][1]:,}-{cpu['window'][2]:,} of contig {cpu['window'][0]} of the same {what} workload "
                      f"({cpu['reads']:,} reads, {cpu['sec_per_step']:.1f} s per pass); C++ oracle restating the reference's Scala "
                      "algorithm, one loci task per thread (no JVM on the box)"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_threads = os.cpu_count() or 1
    workload = "germline" if args.workload == "auto" else args.workload
    somatic = workload in ("somatic", "amplicon")
    if workload == "amplicon":
        args.depth = 10000.0
    contigs, loci_all = genome(world, workload, args.contig_length)
    names = [c[0] for c in contigs]
    total_loci = sum(r[2] - r[1] for r in loci_all)
    tumor_depth = args.depth if workload == "amplicon" else 2 * args.depth
    if workload == "amplicon":
        wl = (f"somatic-standard, high-depth amplicon: 10,000x tumor + 10,000x normal over 1,000,000 loci, {READ_LEN} bp "
              f"(BASELINE.json configs[4]), loci partitioned over {world} GPU(s)")
    elif somatic:
        wl = (f"somatic-standard, synthetic tumor/normal pair {tumor_depth:g}x/{args.depth:g}x, chr20 shape "
              f"({args.contig_length:,} loci), {READ_LEN} bp (BASELINE.json configs[2])")
    elif world == 1:
        wl = (f"germline-threshold, synthetic chr20 shape ({args.contig_length:,} loci), {args.depth:g}x, {READ_LEN} bp "
              f"(BASELINE.json configs[1]); `somatic` block: configs[2] (tumor {tumor_depth:g}x / normal {args.depth:g}x, same contig)")
    else:
        wl = (f"germline-threshold, synthetic whole genome {args.depth:g}x, {READ_LEN} bp: the 25 GRCh37 contigs of "
              f"DistributedUtilSuite.scala:72 ({total_loci:,} loci), partitionLociUniformly over {world} GPUs with mid-contig cuts "
              f"and boundary reads on both neighbours, reads generated on each device (BASELINE.json configs[3]); the e2e leg "
              f"runs from pinned host buffers over a chr20-sized slice ({CHR20:,} loci) of every rank's shard")
    config = {"workload": wl, "threshold_percent": 8, "total_loci": total_loci, "depth": args.depth, "read_length": READ_LEN,
              "parallelism": f"loci-partitioned x{world} (partitionLociUniformly)", "l2": "inputs larger than L2 (no flush needed)"}
    depths_of = lambda is_som: [(1, tumor_depth), (0, args.depth)] if is_som else [(0, args.depth)]

    if args.impl == "reference":
        if rank != 0:
            return
        c0 = loci_all[0][0]
        window = (c0, 0, min(args.cpu_window // (3 if somatic else 1), contigs[c0][1] - 1))
        if workload == "amplicon":
            window = (0, 0, 1500)
        cpu = cpu_leg(args, contigs, window, somatic, depths_of(somatic), max(1, args.steps), max(0, min(args.warmup, 1)), n_threads)
        line = {"impl": "reference", "metric": "loci_per_sec", "value": cpu["loci_per_s"], "unit": "loci/s",
                "reads_per_sec": cpu["reads_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": cpu["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f64" if somatic else "u8", "data": "synthetic", "config": config,
                "cpu_baseline": cpu_baseline_block(cpu, workload),
                "e2e": {"value": cpu["loci_per_s"], "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    from guacamole_b200 import callers
    from guacamole_b200.distributed import ranges_of_rank
    from guacamole_b200.loci import partition_loci_uniformly

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctx = callers.Context(local_rank)
    ctx.set_option(abi.OPT_HOST_THREADS, max(1, n_threads // world))
    if args.no_pack_overlap:
        ctx.set_option(abi.OPT_PACK_OVERLAP, 0)
    comm = None
    if world > 1:  # the library's own NCCL communicator: the id travels by the host side's broadcast
        ids = [callers.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = callers.Comm(ctx, ids[0], rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(values, op):
        t = torch.tensor(values, device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=op)
        return [float(x) for x in t]

    my_ranges = ranges_of_rank(partition_loci_uniformly(world, loci_all), rank)
    peak, peak_src = peaks()

    def run_workload(is_somatic, ranges, steps, e2e_steps, e2e_ranges, cpu_window):
        """One workload on this rank's `ranges`: resident leg, the pack + call companion, the gather, the parity check against
        the oracle inside `cpu_window`, the end-to-end leg over `e2e_ranges`.  Returns this rank's numbers."""
        ctx.set_option(abi.OPT_PACK_QUALITIES, 1 if is_somatic else 0)
        samples = depths_of(is_somatic)
        n_loci = sum(r[2] - r[1] for r in ranges)
        t_setup = time.perf_counter()
        packed, n_reads, alg, gen_ms = [], 0, 0.25 * n_loci, 0.0
        for s, d in samples:  # generated on the device, packed there (the large columns move into the read set)
            dv = synth.generate_device(ctx, contigs, depth=d, read_length=READ_LEN, seed=args.seed, sample=s,
                                       windows=synth.shard_windows(ranges, READ_LEN), with_qualities=is_somatic)
            n_reads += dv.n_reads
            alg += algorithmic_bytes_of(dv.n_reads, dv.n_cigar_ops, dv.n_bases, is_somatic)
            gen_ms += dv.kernel_ms
            packed.append(ctx.pack_synth(dv))
            dv.free()
        pack_ms = sum(p.pack_kernel_ms for p in packed)
        expand_ms = sum(p.expand_kernel_ms for p in packed)
        setup_s = time.perf_counter() - t_setup

        def call(reads_list, rngs):
            if is_somatic:
                return callers.somatic_standard(ctx, reads_list[0], reads_list[1], rngs, odds_threshold=20, min_alignment_quality=1)
            return callers.germline_threshold(ctx, reads_list[0], rngs, threshold=8)

        # ---- device-resident leg (records in canonical order)
        for _ in range(args.warmup):
            res = call(packed, ranges)
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        ctx.timer_start()
        t0 = time.perf_counter()
        tile_ms = exact_ms = 0.0
        launches = 0
        for _ in range(steps):
            res = call(packed, ranges)
            st = res.stats
            tile_ms += st["tile_kernel_ms"]
            exact_ms += st["exact_kernel_ms"]
            launches += st["kernel_launches"]
        dev_ms = ctx.timer_stop()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop()
        out = {"n_loci": n_loci, "n_reads": n_reads, "alg": alg, "step_ms": dev_ms / steps, "tile_ms": tile_ms / steps,
               "exact_ms": exact_ms / steps, "launches": launches, "wall_ms": wall_ms / steps, "clocks": clocks, "gen_ms": gen_ms,
               "pack_ms": pack_ms, "expand_ms": expand_ms, "setup_s": setup_s, "records": len(res)}
        # every kernel from the raw columns resident in HBM to the records
        out["pack_plus_call_ms"] = pack_ms + out["tile_ms"] + out["exact_ms"]
        resident_digest = record_digest(res, is_somatic) if e2e_ranges == ranges else None

        # ---- parity: the engine's records inside the CPU window against the oracle's (the oracle's own run of the generator)
        cpu = None
        if cpu_window is not None:
            cpu = cpu_leg(args, contigs, cpu_window, is_somatic, samples, 1, 0, n_threads, keep_records=True)
            c, lo, hi = cpu["window"]
            got = somatic_keys(res, c, lo, hi) if is_somatic else germline_keys(res, c, lo, hi)
            out["parity_window"] = "ok" if got == cpu["keys"] else f"MISMATCH ({len(got)} engine vs {len(cpu['keys'])} oracle records)"
            out["parity_records"] = len(got)
            cpu["keys"] = None
        out["cpu"] = cpu

        # ---- the gather (N > 1, germline): records + depth histogram to rank 0 by the library's NCCL exchange, checked
        if comm is not None and not is_somatic:
            cmp, gen, _ = res.compact()
            mine = [float(len(cmp)), float(len(gen)), float(int(gen["start"].sum()) % (1 << 40)) if len(gen) else 0.0]
            merged = comm.gather(res, 0)
            callers.depth_histogram(ctx, packed[0], ranges, fetch=False)
            hist = comm.reduce_depth_histogram(0)
            tot = reduce(mine, dist.ReduceOp.SUM)
            if rank == 0:
                mc, mg, _ = merged.compact()
                ok = (len(mc) == int(tot[0]) and len(mg) == int(tot[1]) and int(mg["start"].sum()) % (1 << 40) == int(tot[2]) % (1 << 40)
                      and bool(np.all(mc[:-1] <= mc[1:])) and int(hist.sum()) == total_loci)
                out["gather_check"] = "ok" if ok else "MISMATCH"
                out["total_records"] = len(merged)
                out["mean_depth"] = float((hist.astype(np.float64) * np.arange(256)).sum() / max(1.0, float(hist.sum())))
            del merged, cmp, gen
        del res
        for p in packed:
            p.free()

        # ---- end-to-end leg: (pinned) host buffers -> pack -> call (-> NCCL gather) -> records in host memory, every step
        ctx.set_option(abi.OPT_TRIM_CACHE, 1)  # (the resident leg's buffers are of another shape: give them back first)
        e2e_loci = sum(r[2] - r[1] for r in e2e_ranges)
        hosts = []
        for s, d in samples:
            dv = synth.generate_device(ctx, contigs, depth=d, read_length=READ_LEN, seed=args.seed, sample=s,
                                       windows=synth.shard_windows(e2e_ranges, READ_LEN), with_qualities=is_somatic)
            wide = dv.download(pinned=not args.e2e_compact)
            dv.free()
            if args.e2e_compact:  # what a shim fills straight from BAM records: 4-bit bases, 32-bit starts / offsets (guac_read_batch_v2)
                hosts.append(callers.CompactBatch(wide.c, pinned=True, fixed_length=True))
                wide.free()
            else:
                hosts.append(wide)
        h2d = d2h = n_touch = 0
        e2e_digest = None
        t1 = time.perf_counter()
        e2e_warm = max(3, args.warmup)  # (untimed: the pinned-block and device-buffer caches settle over the first few packs)
        for i in range(e2e_warm + e2e_steps):
            if i == e2e_warm:
                barrier()
                t1 = time.perf_counter()
            _t = [time.perf_counter()]
            fresh = [ctx.pack_v2(h, names) if args.e2e_compact else ctx.pack_c(h.c, names) for h in hosts]
            _t.append(time.perf_counter())
            r = call(fresh, e2e_ranges)
            _t.append(time.perf_counter())
            if comm is not None and not is_somatic:
                g = comm.gather(r, 0)
                _t.append(time.perf_counter())
                n_touch = len(g.records) if rank == 0 else 0   # the guac_threshold_record view of every gathered record
                _t.append(time.perf_counter())
                d2h = int(g.stats["d2h_bytes"]) if rank == 0 else 0
                del g
                _t.append(time.perf_counter())
                if os.environ.get("GUAC_TRACE"):
                    print(f"[bench e2e rank {rank}] pack {(_t[1]-_t[0])*1e3:.1f} call {(_t[2]-_t[1])*1e3:.1f} gather {(_t[3]-_t[2])*1e3:.1f} view {(_t[4]-_t[3])*1e3:.1f} del {(_t[5]-_t[4])*1e3:.1f}", file=sys.stderr)
            else:
                n_touch = len(r.records)
                d2h = int(r.stats["d2h_bytes"])
            h2d = sum(f.h2d_bytes for f in fresh) + int(r.stats["h2d_bytes"])
            del r
            for f in fresh:
                f.free()
        barrier()
        out["e2e_ms"] = (time.perf_counter() - t1) * 1e3 / e2e_steps
        out["e2e_loci"] = e2e_loci
        if resident_digest is not None and comm is None:  # one more step, untimed, whose records are fingerprinted
            fresh = [ctx.pack_v2(h, names) if args.e2e_compact else ctx.pack_c(h.c, names) for h in hosts]
            r = call(fresh, e2e_ranges)
            e2e_digest = record_digest(r, is_somatic)
            del r
            for f in fresh:
                f.free()
        out["h2d"], out["d2h"], out["e2e_records"] = h2d, d2h, n_touch
        if resident_digest is not None and e2e_digest is not None:  # the end-to-end leg returns the resident leg's records, all of them
            out["e2e_check"] = "ok" if e2e_digest == resident_digest else f"MISMATCH ({e2e_digest} vs {resident_digest})"
        for h in hosts:
            h.free()
        return out

    def summarise(o, is_somatic, kernel):
        """max-over-ranks timings -> one workload block"""
        step_ms, e2e_ms, tile_ms, exact_ms, pack_ms, expand_ms, gen_ms, ppc = reduce(
            [o["step_ms"], o["e2e_ms"], o["tile_ms"], o["exact_ms"], o["pack_ms"], o["expand_ms"], o["gen_ms"], o["pack_plus_call_ms"]],
            dist.ReduceOp.MAX)
        loci, reads, alg, e2e_loci, recs, h2d = reduce([o["n_loci"], o["n_reads"], o["alg"], o["e2e_loci"], o["records"], o["h2d"]], dist.ReduceOp.SUM)
        # roofline of the dominant kernel: the slowest rank's launch against one rank's share of the algorithmic bytes
        achieved = (alg / world) / (tile_ms * 1e-3) / 1e9
        blk = {"value": loci / (step_ms * 1e-3), "unit": "loci/s", "reads_per_sec": reads / (step_ms * 1e-3), "ms_per_step": step_ms,
               "loci": int(loci), "reads": int(reads), "records_per_step": int(o.get("total_records", recs)),
               "e2e": {"value": e2e_loci / (e2e_ms * 1e-3), "unit": "loci/s", "ms_per_step": e2e_ms, "loci": int(e2e_loci),
                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": o["d2h"], "records_read": o["e2e_records"],
                       "input": "guac_read_batch_v2 (4-bit bases, 32-bit columns)" if args.e2e_compact else "guac_read_batch (ASCII bases, 64-bit columns)"},
               "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": measured_traffic(kernel, loci / world, args.depth), "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": alg / world, "kernel_ms": tile_ms, "other_kernels_ms": exact_ms,
                            "whole_step_achieved_gbs": (alg / world) / (step_ms * 1e-3) / 1e9},
               "pack_plus_call_ms": ppc, "pack_kernel_ms": pack_ms, "expand_kernel_ms": expand_ms, "generate_kernel_ms": gen_ms,
               "pack_plus_call_loci_per_s": loci / (ppc * 1e-3), "gpu_launches": int(o["launches"]), "wall_ms_per_step": o["wall_ms"],
               "clocks": o["clocks"], "setup_s": o["setup_s"]}
        if "parity_window" in o:
            blk["parity_window"] = o["parity_window"]
            blk["parity_window_records"] = o["parity_records"]
        if "e2e_check" in o:
            blk["e2e_check"] = o["e2e_check"]
        if "gather_check" in o:
            blk["gather_check"] = o["gather_check"]
            blk["mean_depth"] = o["mean_depth"]
        if o.get("cpu"):
            blk["cpu_baseline"] = cpu_baseline_block(o["cpu"], "somatic" if is_somatic else "germline")
        return blk

    e2e_steps = args.e2e_steps or min(args.steps, 5)
    # the e2e slice: the whole shard at N = 1; a chr20-sized prefix of the shard at N > 1
    if world == 1 or workload == "amplicon":
        e2e_ranges = my_ranges
    else:
        e2e_ranges, left = [], CHR20
        for r in my_ranges:
            take = min(left, r[2] - r[1])
            if take > 0:
                e2e_ranges.append((r[0], r[1], r[1] + take, r[3]))
                left -= take
    c0, s0 = my_ranges[0][0], my_ranges[0][1]
    if rank != 0:
        cpu_window = None
    elif workload == "amplicon":
        cpu_window = (0, 0, 1500)
    else:
        w = min(args.cpu_window // (3 if somatic else 1), my_ranges[0][2] - s0)
        cpu_window = (c0, max(0, s0 - SPAN), s0 + w)
    main_out = run_workload(somatic, my_ranges, args.steps, e2e_steps, e2e_ranges, cpu_window)
    main_blk = summarise(main_out, somatic, "k_somatic" if somatic else "k_call_tile")
    som_blk = None
    if world == 1 and workload == "germline" and not args.no_somatic:
        w = min(args.cpu_window // 3, my_ranges[0][2])
        som_out = run_workload(True, my_ranges, max(1, min(args.steps, 5)), 2, my_ranges, (c0, 0, w))
        som_blk = summarise(som_out, True, "k_somatic")
        som_blk["workload"] = (f"somatic-standard, synthetic tumor/normal pair {tumor_depth:g}x/{args.depth:g}x, chr20 shape "
                               f"({args.contig_length:,} loci), {READ_LEN} bp (BASELINE.json configs[2])")
        som_blk["dtype"] = "f64"

    failed = False
    if rank == 0:
        line = {"metric": "loci_per_sec", "value": main_blk["value"], "unit": "loci/s", "reads_per_sec": main_blk["reads_per_sec"],
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_blk["ms_per_step"],
                "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f64" if somatic else "u8", "data": "synthetic (generated on the device from the seed)", "config": config}
        for k in ("records_per_step", "gpu_launches", "e2e", "roofline", "clocks", "pack_plus_call_ms", "pack_kernel_ms", "expand_kernel_ms",
                  "generate_kernel_ms", "pack_plus_call_loci_per_s", "wall_ms_per_step", "parity_window", "parity_window_records",
                  "e2e_check", "gather_check", "mean_depth", "cpu_baseline", "setup_s"):
            if k in main_blk:
                line[k] = main_blk[k]
        line["e2e"]["steps"] = e2e_steps
        if som_blk:
            line["somatic"] = som_blk
        bad = [b[k] for b in (main_blk, som_blk or {}) for k in ("parity_window", "e2e_check", "gather_check") if b.get(k, "ok") != "ok"]
        print(json.dumps(line), flush=True)
        if bad:
            print("PARITY FAILURE: " + "; ".join(bad), file=sys.stderr, flush=True)
            failed = True
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    if failed:  # (after the line: a number whose records differ from the oracle's is not a result)
        sys.exit(1)


if __name__ == "__main__":
    main()
