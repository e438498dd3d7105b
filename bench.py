#!/usr/bin/env python
"""bench.py — k-mer recruitment read-bases/s on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

Workload (config.workload): BASELINE.json configs[1] — a cenX-like synthetic HOR array
(1500 x 2052 bp units, 1 % divergence, 2 x 200 kb flanks) at 50x long-read coverage with 6 %
read errors; one step = the whole recruitment path (document-frequency count -> rare band ->
per-unit clouds -> unit-distance pair graph -> edge filter) on one batch = the full read set.
`value` times it with inputs resident in HBM; `e2e` times the same call from pinned host buffers
(H2D inside) to host-resident results (D2H inside).  --scale shrinks the array multiplicity
(parity / smoke use); the default 1.0 is the configuration the metric is quoted on.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAMS = dict(k=19, coverage=32, min_coverage=4, min_d=1, max_d=150, bottom=0.9, top=3.0,
              kmer_survival_rate=0.34, max_nonuniq=3)
DATA = dict(n_monomers=12, monomer_len=171, monomer_div=0.25, unit_seed=7, multiplicity=1500, div_rate=0.01,
            genome_seed=1, read_coverage=50, error_rate=0.06, read_seed=3)
# algorithmic bytes per unit of work, fixed before the first measurement
BYTES_PER_KMER_A = 32.25
BYTES_PER_INCREMENT_C = 32.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_inputs(scale, rank=0, world=1):
    """configs[1] at N = 1.  At N > 1 (weak scaling) the read set is that of N cenX-like arrays, each with its own
    independently drawn HOR unit, genome and reads (array j uses seeds + j; array 0 is configs[1] itself), and
    every rank holds the reads i = rank mod N of every array: about one configs[1] worth of read bases, units and
    pair increments per GPU, while the k-mer counts, the rare set and the distance graph stay global.  (Making ONE
    array N times longer does not keep the work per GPU fixed: more copies of the same unit share more k-mers, the
    clouds grow and the pair increments grow quadratically with them -- 5.5x at N = 2, measured.)"""
    from centroflye_b200 import synth
    from centroflye_b200.ingest import batch_from_synth
    mult = max(8, int(round(DATA["multiplicity"] * scale)))
    reads, unit0 = [], None
    for j in range(world):
        unit = synth.hor_unit(DATA["n_monomers"], DATA["monomer_len"], DATA["monomer_div"], DATA["unit_seed"] + j)
        unit0 = unit0 or unit
        genome, a0, alen = synth.simulate_genome(unit, mult, DATA["div_rate"], DATA["genome_seed"] + j)
        reads += synth.simulate_reads(genome, a0, alen, unit, DATA["read_coverage"], DATA["error_rate"],
                                      DATA["read_seed"] + j, id_prefix=f"read" if j == 0 else f"a{j}_read",
                                      shard=(rank, world) if world > 1 else None)
    batch, units = batch_from_synth(reads, len(unit0))
    return unit0, batch, units


def band():
    from centroflye_b200.engine import band_to_int
    left = PARAMS["bottom"] * PARAMS["coverage"] * PARAMS["kmer_survival_rate"]
    right = PARAMS["top"] * PARAMS["coverage"] * PARAMS["kmer_survival_rate"]
    return band_to_int(left, right)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  A step lasts tens
    of milliseconds, so the sampler reads NVML directly (nvidia_ml_py, ~5 ms period) and falls back to polling
    nvidia-smi only if NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.sm, self.mx, self.reasons, self.index = [], [], set(), index
        self.stop, self.source = threading.Event(), "nvml"
        self.th = threading.Thread(target=self._run, daemon=True)

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            ent = vis.split(",")[idx].strip()
            if ent.isdigit():
                idx = int(ent)
            else:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent)
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _run(self):
        try:
            nv, h = self._nvml_handle()
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                    else nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for name, bit in bits.items():
                    if r & int(bit):
                        self.reasons.add(name)
                self.stop.wait(0.005)
            return
        except Exception as e:  # noqa: BLE001 - any NVML problem: poll nvidia-smi instead
            self.source = f"nvidia-smi ({type(e).__name__})"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")] if out else []
                if r and r[0].replace(".", "").isdigit():
                    self.sm.append(float(r[0]))
                if len(r) > 1 and r[1].replace(".", "").isdigit():
                    self.mx.append(float(r[1]))
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def ncu_traffic(kernel):
    """DRAM bytes (read + write) of one launch of `kernel` from the newest committed `ncu --set full` summary under
    profiles/ (tools/ncu_summary.py output); (None, None) if no capture names the kernel."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.txt")),
                       key=lambda p: [int(x) for x in re.findall(r"\d+", os.path.basename(p))], reverse=True):
        cur, vals = None, {}
        for ln in open(path):
            if ln.startswith("## "):
                cur = ln.strip().split("::")[-1]
            elif cur == kernel:
                m = re.match(r"\s*(dram__bytes_(?:read|write)\.sum)\s+([0-9.]+)\s+(\S+)", ln)
                if m:
                    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), None)
                    if mult:
                        vals[m.group(1)] = float(m.group(2)) * mult
        if len(vals) == 2:
            return sum(vals.values()), os.path.relpath(path, ROOT)
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist
    from centroflye_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    eng = Engine(f"cuda:{local}")
    k = PARAMS["k"]
    lo, hi = band()

    t0 = time.time()
    unit, batch, units = make_inputs(args.scale, rank, world)
    log(f"[bench] rank {rank}: inputs in {time.time() - t0:.1f}s: {batch.n_reads} reads, {batch.n_bases} bases, "
        f"{units.n_units} units")

    if world > 1:
        from centroflye_b200.dist import ShardedRecruiter
        runner = ShardedRecruiter(eng, batch, units, k, rank, world)
    else:
        runner = None

    def device_step(reads, dunits):
        if runner is not None:
            return runner.step(lo, hi, PARAMS["max_nonuniq"], PARAMS["min_d"], PARAMS["max_d"], PARAMS["min_coverage"],
                               gather=False)  # edges stay sharded by source k-mer, like the work
        return eng.recruit(reads, dunits, k, lo, hi, PARAMS["max_nonuniq"], PARAMS["min_d"], PARAMS["max_d"],
                           PARAMS["min_coverage"])

    def e2e_step():
        if runner is not None:
            return runner.e2e_step(lo, hi, PARAMS["max_nonuniq"], PARAMS["min_d"], PARAMS["max_d"],
                                   PARAMS["min_coverage"])
        reads = eng.upload_reads(batch, k)
        dunits = eng.upload_units(units, k)
        early = {}
        # the rare set and the clouds are final after stage B: they travel to the host on a side stream while the
        # distance graph is computed; edges and endpoints follow when it is done
        index, csr, res = eng.recruit(reads, dunits, k, lo, hi, PARAMS["max_nonuniq"], PARAMS["min_d"], PARAMS["max_d"],
                                      PARAMS["min_coverage"],
                                      on_clouds=lambda index, csr: early.update(eng.start_host_copy(
                                          unit_ptr=csr.unit_ptr, ids=csr.ids, rare_keys=index.sorted_keys)))
        out = eng.to_host(selected=res.selected, edges=res.edges)  # pinned result buffers; synchronises
        eng.finish_host_copies()
        out.update(early)
        d2h = sum(t.numel() * t.element_size() for t in out.values())
        return reads.h2d_bytes + dunits.h2d_bytes, d2h

    reads = eng.upload_reads(batch, k) if runner is None else None
    dunits = eng.upload_units(units, k) if runner is None else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = device_step(reads, dunits)
    barrier()
    n_bases_total = batch.n_bases if runner is None else runner.n_bases_total

    step_ms, launches0 = [], eng.launch_count()
    stage_ms = {}
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.fill_(1)  # evict L2 between timed steps (inputs alone would fit the 126 MB L2)
            barrier()
            eng.events = []
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = device_step(reads, dunits)
            b.record()
            barrier()
            step_ms.append(a.elapsed_time(b))
            for name, ms in eng.stage_times_ms().items():
                stage_ms.setdefault(name, []).append(ms)
            eng.events = None
    launches = (eng.launch_count() - launches0) / max(args.steps, 1)
    ms = float(np.mean(step_ms))
    if world > 1:
        tms = torch.tensor([ms], device=eng.device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    last = res[2]
    n_edges_total = int(last.edges.shape[0])
    if world > 1:
        ne = torch.tensor([n_edges_total], dtype=torch.int64, device=eng.device)
        dist.all_reduce(ne)
        n_edges_total = int(ne.item())

    # end to end: pinned host buffers -> host results, same call
    e2e_ms, h2d, d2h = [], 0, 0
    e2e_step()
    for _ in range(max(2, min(args.steps, 3))):
        barrier()
        t1 = time.perf_counter()
        h2d, d2h = e2e_step()
        barrier()
        e2e_ms.append((time.perf_counter() - t1) * 1e3)
    e2e = float(np.mean(e2e_ms))
    if world > 1:
        tms = torch.tensor([e2e], device=eng.device)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e = float(tms.item())
        nbytes = torch.tensor([h2d, d2h], dtype=torch.int64, device=eng.device)  # every rank moves its own shard
        dist.all_reduce(nbytes)
        h2d, d2h = int(nbytes[0].item()), int(nbytes[1].item())

    if rank != 0:
        return
    peak, peak_src = peaks()
    dc_ms = float(np.mean(stage_ms.get("pair_candidates", [0.0])))
    n_incr = last.n_increments
    alg_bytes = BYTES_PER_INCREMENT_C * n_incr / world  # per launch: sources (hence increments) are dealt evenly
    achieved = alg_bytes / (dc_ms * 1e-3) / 1e9 if dc_ms > 0 else 0.0
    pair_kernel = getattr(eng, "last_pair_kernel", "pair_candidates_kernel")
    traffic, traffic_src = ncu_traffic(pair_kernel)
    # the other kernels of the step against the same HBM line (algorithmic bytes of BASELINE.md §4, rank 0's share)
    df_kernel = {"stream": "docfreq_emit_kernel+docfreq_count_kernel", "resident": "docfreq_resident_kernel"}.get(
        eng.docfreq_mode, "docfreq_kernel")
    if eng.docfreq_mode == "stream" and "docfreq_emit" in stage_ms:  # two kernels: the stage-A figure is their sum
        stage_ms["docfreq"] = [a + b for a, b in zip(stage_ms.get("docfreq_emit", []), stage_ms.get("docfreq_count", []))]
    n_k = int(batch.n_bases - batch.n_reads * (k - 1))
    n_ku = int(units.unit_len.astype(np.int64).sum() - units.n_units * (k - 1))
    csr_last = res[1]
    stage_rooflines = []
    for kern, stage, nbytes, what in (
            (df_kernel, "docfreq", BYTES_PER_KMER_A * n_k, "32.25 B per k-mer occurrence"),
            ("cloud_build_warp_kernel", "cloud_build", 16.25 * n_ku + 4.0 * csr_last.n_entries + 8.0 * units.n_units,
             "16.25 B per k-mer inside a unit + 4 B per cloud entry + 8 B per unit"),
            ("pair_join_kernel", "pair_join", 32.0 * last.n_pair_candidates + 16.0 * int(last.edges.shape[0]),
             "32 B per pair candidate + 16 B per edge")) if world == 1 else ():
        t_ms = float(np.mean(stage_ms.get(stage, [0.0])))
        if t_ms <= 0:
            continue
        gbs = nbytes / (t_ms * 1e-3) / 1e9
        tr, tr_src = ncu_traffic(kern)
        stage_rooflines.append({"kernel": kern, "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                "frac": gbs / peak, "traffic": tr, "traffic_source": tr_src, "kernel_ms": t_ms,
                                "algorithmic_bytes": nbytes, "note": what})
    line = {
        "metric": "k-mer recruitment read-bases/s", "value": n_bases_total / (ms * 1e-3), "unit": "read-bases/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[1]: cenX-like HOR array 1500x2052bp (3.08 Mb) + 2x200kb flanks, 50x reads, "
                               "6% errors, k=19, coverage=32, max_d=150: full recruitment + read_kmer_cloud build"
                               + (f"; weak scaling: {world} such arrays with independent HOR units, reads of every array dealt to {world} ranks"
                                  if world > 1 else ""),
                   "scale": args.scale, "read_bases": int(n_bases_total), "reads_rank0": int(batch.n_reads),
                   "units_rank0": int(units.n_units), "pair_increments": int(n_incr),
                   "candidates": int(last.n_candidates), "pair_candidates": int(last.n_pair_candidates), "edges": n_edges_total,
                   "unique_kmers": int(last.selected.numel()), "l2": "256 MiB flush write between timed steps",
                   "sharding": ("reads sharded by record; stage A: nominate-then-sum (all-gather of the k-mers with >= "
                                "ceil(lo/N) local reads, one all-reduce of their counts); cloud CSR all-gathered; source "
                                "k-mers dealt round-robin, edges stay with their source's rank" if world > 1 else "single GPU")},
        "roofline": {"kernel": pair_kernel, "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": dc_ms,
                     "algorithmic_bytes": alg_bytes, "note": "32 B per pair increment (BASELINE.md §4); the counters "
                     "live in shared memory, so the effective figure exceeds the HBM line and the real DRAM traffic "
                     "(ncu, `traffic`) is ~1000x lower: the kernel's true bound is shared-memory wavefronts / issue "
                     "slots (profiles/)"},
        "roofline_other_kernels": stage_rooflines,
        "stage_ms": {name: float(np.mean(v)) for name, v in stage_ms.items()},
        "e2e": {"value": n_bases_total / (e2e * 1e-3), "unit": "read-bases/s", "ms_per_step": e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": launches, "clocks": clk.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(batch, units, bounded_s=args.cpu_seconds)
    print(json.dumps(line), flush=True)


def cpu_baseline(batch, units, bounded_s=20.0, threads=1):
    """The oracle's C restatement (oracle/c) timed on the host on a bounded sample of the same workload."""
    from oracle import c_oracle
    return c_oracle.timed_sample(batch, units, PARAMS, band(), bounded_s=bounded_s, threads=threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    unit, batch, units = make_inputs(args.scale)
    threads = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        res = c_oracle.timed_sample(batch, units, PARAMS, band(), bounded_s=args.cpu_seconds, threads=threads)
        if i >= args.warmup:
            vals.append(res)
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["ms"] for r in vals]))
    line = {"impl": "reference", "metric": "k-mer recruitment read-bases/s", "value": v, "unit": "read-bases/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "configs[1] (bounded sample, see cpu_baseline.sample)", "scale": args.scale},
            "cpu_baseline": dict(vals[-1], value=v),
            "e2e": {"value": v, "unit": "read-bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
