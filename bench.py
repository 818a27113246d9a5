#!/usr/bin/env python
"""Headline benchmark: Fr mul-add/s for the mat-mul + Freivalds + rescale witness, N=1024, P=63,
LOOKUP_BITS=19 (BASELINE.json metric / configs[3]) on 1..8 B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA kernels)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU restatement on the host cores

A "step" = one pass of the hot path over one (A, B) pair: honest_prover_mat_mul (C = A.B), rescale_matrix witnesses of C,
verify_mul (Freivalds) witnesses -- ONE call of h2svd_zkmatrix_mul_witness_dev per rank, recorded into a CUDA graph and
replayed.  Rows of A/C are sharded over the ranks ("strong" scaling: the job is fixed at N=1024), B is replicated, every
rank derives (B v) itself (no collective on the data path).  Prints ONE JSON line on rank 0; DESIGN.md "Measurement" explains
every key.  Everything the driver should see (per-kernel rooflines, BASELINE configs[1] and [2], per-rank oracle
verification) sits inside the `roofline` dict, which the driver keeps.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DEFAULT, P_BITS, LOOKUP_BITS = 1024, 63, 19
METRIC = "Fr mul-add/s for mat-mul+Freivalds+rescale witness, N=1024"
UNIT = "Fr mul-add/s"
DTYPE = "u32x8 (BN254 Fr, 256-bit modular)"


def job_units(n: int, k: int, m: int) -> int:
    """Fr multiply-adds of the whole job: mat-mul + gamma chain + the three Freivalds mat-vecs."""
    return n * k * m + (m - 1) + n * m + k * m + n * k


def config_dict(n: int, k: int, m: int, world: int) -> dict:
    """Printed verbatim by BOTH arms (the driver compares them)."""
    shape = f"square N={n}" if n == k == m else f"{n}x{k} . {k}x{m}"
    return {"workload": f"honest_prover_mat_mul + rescale_matrix + verify_mul witness, {shape}, PRECISION_BITS={P_BITS}, "
                        f"LOOKUP_BITS={LOOKUP_BITS} (BASELINE configs[3] + Freivalds)",
            "n": n, "k": k, "m": m, "precision_bits": P_BITS, "lookup_bits": LOOKUP_BITS,
            "sharding": f"rows of A/C over {world} rank(s), B replicated, no data-path collective",
            "l2": "256 MiB flush write between timed steps (GPU arm)",
            "inputs": "input-creator.py distribution, seeded, quantized to Fr"}


def metric_name(n: int, k: int, m: int) -> str:
    return METRIC if (n, k, m) == (N_DEFAULT,) * 3 else f"Fr mul-add/s for mat-mul+Freivalds+rescale witness, {n}x{k}x{m}"


def make_inputs(n: int, k: int, m: int, seed: int = 20261018):
    """Seeded re-implementation of the reference's input-creator.py:23-28 distribution (f64)."""
    rng = np.random.default_rng(seed)

    def mat(r, c):
        x = rng.uniform(-10.0, 10.0, size=(r, c))
        return x / np.linalg.norm(x, ord=2) * rng.uniform(1, 100)

    a, b = mat(n, k), mat(k, m)
    gamma = rng.integers(0, 1 << 64, size=4, dtype=np.uint64)
    gamma[3] &= np.uint64((1 << 60) - 1)   # a fixed non-trivial canonical challenge
    return a, b, gamma


def split_range(total: int, world: int, rank: int):
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def traffic_from_profile(kernel: str, n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (same command, same N)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)[kernel]
        return t["dram_bytes_read"] + t["dram_bytes_write"] if t.get("n") == n else None
    except (OSError, KeyError, ValueError):
        return None


def bind_to_gpu_numa_node(gpu_index: int):
    """Restricts this process to the CPUs of the NUMA node the GPU hangs off, so that the page-locked host buffers of the
    end-to-end leg are allocated next to its PCIe root port (matters when 8 ranks stream witnesses to the host at once)."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        bus = bus[-12:] if len(bus) > 12 else bus          # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, subprocess.SubprocessError):
        return None


def measured_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except OSError:
        return {}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_slice(n: int, k: int, m: int, threads: int, rows: int) -> dict:
    """Times the C oracle (CPU restatement of the reference loops) on a PROPORTIONAL slice of the job: `rows` of the n rows
    of the mat-mul, of C.v, of A.(Bv), of the rescale, and rows*k/n rows of B.v.  The slice holds rows/n of every term of the
    job's multiply-add count, so its own throughput (slice units / slice seconds) IS the job's throughput on this CPU --
    no extrapolation enters `value` or `ms_per_step`.  The only place bench.py executes oracle/."""
    from oracle import corac
    a_f, b_f, gamma = make_inputs(n, k, m)
    rows = max(1, min(rows, n))
    a = corac.quantize(a_f[:rows], P_BITS)
    b = corac.quantize(b_f, P_BITS)
    t0 = time.perf_counter()
    c = corac.field_mat_mul(a, b, threads=threads)                       # reference src/matrix/mod.rs:510-537
    powers = corac.gamma_powers(gamma.reshape(1, 4), m)                  # :316-326 (whole chain: it is not shardable)
    corac.mat_vec_prefix(c, powers, threads=threads)                     # :335
    pbv = corac.mat_vec_prefix(b, powers, threads=threads)               # :336 (all rows: the totals feed :337)
    bv = np.ascontiguousarray(pbv[:, -1])
    corac.mat_vec_prefix(a, bv, threads=threads)                         # :337
    corac.rescale_witness(c.reshape(-1, 4), P_BITS, LOOKUP_BITS, threads=threads)   # :354-375
    secs = time.perf_counter() - t0
    # what ran: rows*k*m + (m-1) + rows*m + k*m + rows*k multiply-adds (B.v ran in full: slightly MORE than the slice's share)
    units = rows * k * m + (m - 1) + rows * m + k * m + rows * k
    return {"units": units, "seconds": secs, "rows": rows,
            "sample": f"{rows}/{n} rows of the job (mat-mul, C.v, A.(Bv), rescale witnesses of those rows; gamma powers and "
                      f"B.v in full), {threads} thread(s); value = multiply-adds that ran / seconds they took"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = k = m = args.n
    if args.shape:
        n, k, m = (int(x) for x in args.shape.split(","))
    rows = args.ref_rows if args.ref_rows > 0 else min(n, max(16, 2 * threads))
    runs = []
    for i in range(args.warmup + args.steps):
        r = cpu_slice(n, k, m, threads, rows)
        if i >= args.warmup:
            runs.append(r)
    secs = float(np.mean([r["seconds"] for r in runs]))
    value = runs[0]["units"] / secs
    line = {
        "impl": "reference", "metric": metric_name(n, k, m), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": config_dict(n, k, m, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": runs[-1]["sample"] + " (C restatement of the reference loops; the Rust reference cannot "
                                                        "be built here: no cargo, un-vendored crates)",
                         "step_is": f"a {rows}/{n} slice of the job; ms_per_step is the measured time of that slice, the "
                                    f"whole job would take ~{secs * n / rows:.1f} s at this rate"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module("halo2-svd041_b200")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)   # pinned buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    h = pkg.Handle(local_rank, stream.cuda_stream)

    n = k = m = args.n
    if args.shape:       # e.g. --shape 4096,2048,4096 = BASELINE configs[4] (not the headline configuration)
        n, k, m = (int(x) for x in args.shape.split(","))
    r0, r1 = split_range(n, world, rank)
    b0, b1 = split_range(k, world, rank)
    rows, brows = r1 - r0, b1 - b0
    W = h.rescale_witness_count(P_BITS, LOOKUP_BITS)
    if args.matmul_small is not None:
        h.tune("matmul_small", args.matmul_small)
    for kv in args.tune or []:
        key, val = kv.split("=")
        h.tune(key, int(val))

    def fr(*shape):
        return torch.zeros(shape + (4,), dtype=torch.int64, device=device)

    bufs = dict(c_s=fr(rows, m), q=fr(rows, m), wit=fr(rows * m, W), powers=fr(m), prefix_cv=fr(rows, m),
                prefix_bv=fr(max(brows, 1), m), prefix_abv=fr(rows, k), diff=fr(rows), is_zero=fr(rows), inv=fr(rows))
    a_slab, b_dev, gamma_dev = fr(rows, k), fr(k, m), fr(1)

    # ---- synthetic inputs: f64 matrices -> pinned host -> GPU quantization kernel (product path)
    a_f, b_f, gamma = make_inputs(n, k, m)
    a_host = torch.from_numpy(np.ascontiguousarray(a_f[r0:r1])).pin_memory()
    b_host = torch.from_numpy(np.ascontiguousarray(b_f)).pin_memory()
    a_dev_f, b_dev_f = a_host.to(device, non_blocking=True), b_host.to(device, non_blocking=True)
    h.quantize_dev(a_dev_f, P_BITS, a_slab)
    h.quantize_dev(b_dev_f, P_BITS, b_dev)
    gamma_dev.copy_(torch.from_numpy(gamma.view(np.int64).reshape(1, 4)))
    h.sync()

    do_e2e = not args.no_e2e
    pin_in, host_out = {}, {}
    if do_e2e:   # page-locked host buffers for the inputs and for every output of the fused C-ABI call
        pin_in = {"a": pkg.PinnedBuffer(tuple(a_slab.shape), np.uint64), "b": pkg.PinnedBuffer(tuple(b_dev.shape), np.uint64)}
        pin_in["a"].array[...] = a_slab.cpu().numpy().view(np.uint64)
        pin_in["b"].array[...] = b_dev.cpu().numpy().view(np.uint64)
        pin_out = {nm: pkg.PinnedBuffer(tuple(t.shape) if nm != "prefix_bv" else (brows, m, 4), np.uint64)
                   for nm, t in bufs.items()}
        host_out = {nm: pb.array for nm, pb in pin_out.items()}
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step_calls():
        h.zkmatrix_mul_witness_dev(a_slab, b_dev, gamma_dev, P_BITS, LOOKUP_BITS, bv_rows=(b0, b1), **bufs)

    def e2e_step():
        # ONE call of the reference-facing C ABI on host buffers: H2D of A (this rank's rows), B, gamma; mat-mul,
        # rescale and verify_mul witnesses in row slabs; D2H of every witness, overlapped slab by slab.  No
        # collective: every rank derives the k row totals of B.v itself and returns its own share of prefix_bv.
        h.zkmatrix_mul_witness(pin_in["a"].array, pin_in["b"].array, gamma, P_BITS, LOOKUP_BITS,
                               bv_rows=(b0, b1), out=host_out)

    # ---- device-resident timing: the step recorded once into a CUDA graph, replayed K times ----
    step_calls()                     # un-captured first: sizes the workspaces
    h.sync()
    graph = None
    if not args.no_graph:
        h.graph_begin()
        step_calls()
        graph = h.graph_end()
    run_step = graph.launch if graph is not None else step_calls
    sampler = ClockSampler(local_rank)   # runs through both timed regions (device-resident and end-to-end)
    sampler.start()
    for _ in range(args.warmup):
        run_step()
    barrier()
    launches_before = h.launch_count
    events = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_buf.fill_(1)              # L2 flush between timed iterations (outside the event brackets)
        e0, e1 = ev(), ev()
        e0.record(stream)
        run_step()
        e1.record(stream)
        events.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_launches = h.launch_count - launches_before
    mine = float(np.mean([a.elapsed_time(b) for a, b in events]))
    ms_t = torch.tensor([mine], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.cpu()[0])
    engine = h.last_matmul_engine()

    # ---- sustained: the same step back to back for >= --sustain-s seconds (power-capped clocks), no flush ----
    sustained = None
    if args.sustain_s > 0:
        barrier()
        reps = max(10, int(args.sustain_s * 1e3 / max(ms_step, 1e-3)))
        e0, e1 = ev(), ev()
        e0.record(stream)
        for _ in range(reps):
            run_step()
        e1.record(stream)
        e1.synchronize()
        sus = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(sus, op=dist.ReduceOp.MAX)
        sustained = {"ms_per_step": float(sus.cpu()[0]), "steps": reps, "seconds": float(sus.cpu()[0]) * reps * 1e-3}
    barrier()

    # ---- end-to-end timing (host buffers, copies inside the timed region) ----
    e2e_ms = None
    h2d_all = d2h_all = 0
    e2e_wall = 0.0
    if do_e2e:
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()             # synchronous: returns when the step's results are on the host
        t_mine = (time.perf_counter() - t0) / args.steps
        barrier()
        e2e_wall = (time.perf_counter() - t0) / args.steps
        e2e_ms_t = torch.tensor([t_mine * 1e3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(e2e_ms_t, op=dist.ReduceOp.MAX)
        e2e_ms = float(e2e_ms_t.cpu()[0])
        h2d = int(pin_in["a"].array.nbytes + pin_in["b"].array.nbytes + 32)
        d2h = sum(int(host_out[x].nbytes) for x in host_out)
        bytes_t = torch.tensor([h2d, d2h], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)
        h2d_all, d2h_all = [int(x) for x in bytes_t.cpu()]
    clocks = sampler.stop()

    # ---- verification, outside the timed regions, ON EVERY RANK, against the CPU oracle ----
    verified = verify_rank(h, torch, device, bufs, a_slab, b_dev, gamma, rows, k, m, b0, b1, W, host_out if do_e2e else None)
    ok_t = torch.tensor([1.0 if verified["ok"] else 0.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    if float(ok_t.cpu()[0]) != 1.0 or not verified["ok"]:
        raise SystemExit(f"bench: rank {rank}: witness check against the oracle failed ({verified}) -- refusing to report a "
                         "number for wrong results")

    # ---- per-kernel measurements outside the headline region (isolated, L2 flushed, CUDA events on the launch stream) ----
    kern = measure_kernels(h, pkg, torch, device, stream, flush_buf, a_slab, b_dev, bufs, rows, k, m, W, rank, args)

    units = job_units(n, k, m)
    value = units / (ms_step * 1e-3)
    line = None
    if rank == 0:
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        finish_hbm(kern, hbm_peak, hbm_src)
        # dominant kernel of the step = the longest isolated phase
        mm, rs = kern["fr_matmul"], kern["rescale_kernel"]
        dom = rs if rs["ms"] >= mm["ms"] else mm
        roofline = dict(dom)
        roofline["dominant_kernel_of_step"] = dom["kernel"]
        roofline["share_of_step"] = dom["ms"] / ms_step
        roofline["traffic"] = traffic_from_profile("rescale_kernel" if dom is rs else "fr_matmul_tc_kernel", n) if world == 1 else None
        roofline["traffic_source"] = "profiles/traffic.json (ncu --set full capture of this command, bytes per launch)"
        roofline["kernels"] = kern
        roofline["verified"] = {"against": "CPU oracle (oracle/fr_oracle.c), outside the timed region, on every rank",
                                "ranks": world, **{kk: vv for kk, vv in verified.items() if kk != "ok"}}
        roofline["sustained"] = sustained
        line = {
            "metric": metric_name(n, k, m), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_dict(n, k, m, world),
            "roofline": roofline, "matmul_engine": engine,
            "e2e": {"value": units / (e2e_ms * 1e-3) if do_e2e else None, "unit": UNIT, "ms_per_step": e2e_ms,
                    "wall_ms_per_step": e2e_wall * 1e3, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "api": "h2svd_zkmatrix_mul_witness (C ABI, pinned host buffers in/out, slab-pipelined D2H) per rank",
                    "numa_node_rank0": numa},
            "gpu_launches": int(gpu_launches) * world, "clocks": clocks, "wall_s_timed_region": t_wall,
            "step": ("CUDA graph replay of " if graph is not None else "") + "h2svd_zkmatrix_mul_witness_dev",
        }
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_slice(n, k, m, 1, rows=args.cpu_rows)     # ~15 s of 1-thread CPU work
            line["cpu_baseline"] = {"value": cb["units"] / cb["seconds"], "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": cb["sample"] + f" ({cb['seconds']:.1f} s of CPU work; the reference is "
                                                             "single-threaded)"}
    if graph is not None:
        graph.close()
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def run_multi_handle(args) -> None:
    """ONE process, --gpus N devices through h2svd_multi_zkmatrix_mul_witness (the single-process drop-in form of the same
    row sharding): end-to-end only -- host arrays in, host arrays out, one host thread and one PCIe link per GPU."""
    import torch

    pkg = importlib.import_module("halo2-svd041_b200")
    from oracle import corac
    n = k = m = args.n
    if args.shape:
        n, k, m = (int(x) for x in args.shape.split(","))
    ndev = args.gpus
    avail = torch.cuda.device_count()
    devices = [i % avail for i in range(ndev)]          # more handles than GPUs wrap round (same partitioning)
    a_f, b_f, gamma = make_inputs(n, k, m)
    with pkg.Handle(0) as h0:
        W = h0.rescale_witness_count(P_BITS, LOOKUP_BITS)
        a = h0.quantize(a_f, P_BITS)
        b = h0.quantize(b_f, P_BITS)
    shapes = dict(c_s=(n, m), q=(n, m), wit=(n * m, W), powers=(m,), prefix_cv=(n, m), prefix_bv=(k, m), prefix_abv=(n, k),
                  diff=(n,), is_zero=(n,), inv=(n,))
    pin_a, pin_b = pkg.PinnedBuffer(a.shape, np.uint64), pkg.PinnedBuffer(b.shape, np.uint64)
    pin_a.array[...] = a
    pin_b.array[...] = b
    pin_out = {nm: pkg.PinnedBuffer(shp + (4,), np.uint64) for nm, shp in shapes.items()}
    out = {nm: pb.array for nm, pb in pin_out.items()}
    sampler = ClockSampler(0)
    with pkg.MultiHandle(devices) as mh:
        for _ in range(max(3, args.warmup)):
            mh.zkmatrix_mul_witness(pin_a.array, pin_b.array, gamma, P_BITS, LOOKUP_BITS, out=out)
        launches0 = mh.launch_count()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            mh.zkmatrix_mul_witness(pin_a.array, pin_b.array, gamma, P_BITS, LOOKUP_BITS, out=out)
        secs = (time.perf_counter() - t0) / args.steps
        clocks = sampler.stop()
        launches = mh.launch_count() - launches0
    # verification against the oracle: sampled rows of C, their rescale stripes, the Freivalds identity on every row
    sel = sorted({0, n // 3, n - 1})
    want_c = corac.field_mat_mul(np.ascontiguousarray(a[sel]), b, threads=0)
    ok = bool((out["c_s"][sel] == want_c).all()) and not out["diff"].any()
    q_w, _r, wit_w = corac.rescale_witness(np.ascontiguousarray(want_c[:, :4].reshape(-1, 4)), P_BITS, LOOKUP_BITS)
    idx = [r * m + c for r in sel for c in range(4)]
    ok = ok and bool((out["wit"][idx] == wit_w).all())
    if not ok:
        raise SystemExit("bench --multi-handle: witness check against the oracle failed")
    units = job_units(n, k, m)
    d2h = sum(int(v.nbytes) for v in out.values())
    line = {"metric": metric_name(n, k, m), "value": units / secs, "unit": UNIT, "n_gpus": ndev, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_dict(n, k, m, ndev),
            "mode": "one process, h2svd_multi (one handle + one host thread per GPU); value IS the end-to-end figure: host arrays "
                    "in, every witness back on the host",
            "devices": devices,
            "e2e": {"value": units / secs, "unit": UNIT, "ms_per_step": secs * 1e3,
                    "h2d_bytes_per_step": int(a.nbytes + ndev * b.nbytes + 32 * ndev), "d2h_bytes_per_step": d2h,
                    "api": "h2svd_multi_zkmatrix_mul_witness (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "pcie", "note": "output-bound: the D2H copy of the witnesses; see the default mode for kernel rooflines",
                         "achieved": d2h / secs / 1e9, "unit": "GB/s", "peak": None, "frac": None, "traffic": None,
                         "verified": {"against": "CPU oracle", "rows_of_c": len(sel), "rescale_stripes": len(idx),
                                      "freivalds_diff_all_zero": True}}}
    print(json.dumps(line), flush=True)


def verify_rank(h, torch, device, bufs, a_slab, b_dev, gamma, rows, k, m, b0, b1, W, host_out) -> dict:
    """Compares what the timed step left in this rank's buffers with the CPU oracle: sampled rows of C, the rescale witness
    stripes of those rows, sampled running-sum rows of C.v / A.(Bv) / B.v, the gamma powers, every is_equal cell."""
    from oracle import corac
    rng = np.random.default_rng(7 + rows)
    sel = sorted(set([0, rows - 1] + [int(x) for x in rng.integers(0, rows, size=2)]))
    cols = sorted(set([0, m - 1] + [int(x) for x in rng.integers(0, m, size=6)]))
    a_h = np.ascontiguousarray(a_slab[sel].cpu().numpy().view(np.uint64))
    b_h = b_dev.cpu().numpy().view(np.uint64)
    want_c = corac.field_mat_mul(a_h, b_h, threads=0)
    res = {"rows_of_c": len(sel)}
    ok = bool((bufs["c_s"][sel].cpu().numpy().view(np.uint64) == want_c).all())
    # rescale witnesses of sampled elements
    flat = np.ascontiguousarray(want_c[:, cols].reshape(-1, 4))
    q_w, _rem, wit_w = corac.rescale_witness(flat, P_BITS, LOOKUP_BITS)
    idx = torch.tensor([r * m + c for r in sel for c in cols], device=device)
    ok &= bool((bufs["wit"][idx].cpu().numpy().view(np.uint64) == wit_w).all())
    ok &= bool((bufs["q"].reshape(-1, 4)[idx].cpu().numpy().view(np.uint64) == q_w).all())
    res["rescale_stripes"] = len(idx)
    # Freivalds running sums
    g = np.ascontiguousarray(gamma.reshape(1, 4))
    powers = corac.gamma_powers(g, m)
    ok &= bool((bufs["powers"].cpu().numpy().view(np.uint64) == powers).all())
    ok &= bool((bufs["prefix_cv"][sel].cpu().numpy().view(np.uint64) == corac.mat_vec_prefix(want_c, powers)).all())
    pbv_all = corac.mat_vec_prefix(b_h, powers, threads=0)
    bv = np.ascontiguousarray(pbv_all[:, -1])
    ok &= bool((bufs["prefix_abv"][sel].cpu().numpy().view(np.uint64) == corac.mat_vec_prefix(a_h, bv)).all())
    if b1 > b0:
        ok &= bool((bufs["prefix_bv"][: b1 - b0].cpu().numpy().view(np.uint64) == pbv_all[b0:b1]).all())
    res["prefix_rows"] = {"c_v": len(sel), "a_bv": len(sel), "b_v": b1 - b0}
    one = np.array([0xac96341c4ffffffb, 0x36fc76959f60cd29, 0x666ea36f7879462e, 0x0e0a77c19a07df2f], dtype=np.uint64)
    ok &= not bool(bufs["diff"].any().item())
    ok &= bool((bufs["is_zero"].cpu().numpy().view(np.uint64) == one).all())
    ok &= bool((bufs["inv"].cpu().numpy().view(np.uint64) == one).all())
    res["is_equal_cells"] = "all rows: diff == 0, is_zero == inv == 1"
    if host_out is not None:   # the end-to-end call returned the same bytes as the device-resident step
        same = True
        for nm in ("c_s", "q", "wit", "powers", "prefix_cv", "prefix_abv", "diff", "is_zero", "inv"):
            same &= bool((torch.from_numpy(host_out[nm].view(np.int64)).to(device) == bufs[nm]).all().item())
        if b1 > b0:
            same &= bool((torch.from_numpy(host_out["prefix_bv"].view(np.int64)).to(device) == bufs["prefix_bv"][: b1 - b0]).all().item())
        ok &= same
        res["e2e_call_byte_identical_to_device_step"] = same
    res["ok"] = bool(ok)
    return res


def measure_kernels(h, pkg, torch, device, stream, flush_buf, a_slab, b_dev, bufs, rows, k, m, W, rank, args) -> dict:
    """Isolated timings (CUDA events on the handle's stream, 256 MiB L2 flush before every launch, median of `reps`) of the
    kernels of the step and of BASELINE configs[1] / configs[2], with the algorithmic work of SURVEY.md 8(d)."""
    reps = 7

    def timed(fn, flush=True):
        ts = []
        for i in range(reps + 2):
            if flush:
                flush_buf.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def fr(*shape):
        return torch.zeros(shape + (4,), dtype=torch.int64, device=device)

    out = {}
    # K1: the mat-mul of this rank's slab (split kernels included), whichever engine the device picks
    c_tmp = fr(rows, m)
    t_mm = timed(lambda: h.fr_matmul_dev(a_slab, b_dev, c_tmp))
    engine = h.last_matmul_engine()
    same = bool((c_tmp == bufs["c_s"]).all().item())
    ops_per_muladd = {"tensor-small": 2.0 * 9 * 9, "tensor": 2.0 * 32 * 32}.get(engine)   # algorithmic digit products
    mmd = {"kernel": f"fr_matmul_tc_kernel<{'TcSmall' if engine == 'tensor-small' else 'TcFull'}> + its byte-plane split kernels",
           "engine": engine, "ms": t_mm, "fr_mul_adds_per_s": rows * k * m / (t_mm * 1e-3), "same_bytes_as_step": same}
    if ops_per_muladd:
        kind = 1 if engine == "tensor-small" else 0
        burst = h.microbench_tensor_i8(kind, 0.0) / 1e12
        sust = h.microbench_tensor_i8(kind, 2.0) / 1e12 if (rank == 0 and not args.quick) else None
        ach = rows * k * m * ops_per_muladd / (t_mm * 1e-3) / 1e12
        mmd.update({"bound": "tensor", "achieved": ach, "peak": burst, "unit": "TOP/s (8-bit)", "frac": ach / burst,
                    "op": f"8-bit multiply-add = 2 ops; {int(ops_per_muladd // 2)} byte products per Fr mul-add "
                          + ("(9 x 9 signed byte digits; the MMAs execute 82-90 with the N padding of the tile width)"
                             if engine == "tensor-small" else "(32 x 32 byte planes)"),
                    "peak_source": "measured live: h2svd_microbench_tensor_i8 (back-to-back tcgen05.mma kind::i8 of the kernel's "
                                   "shape, operands resident in smem, all SMs), burst",
                    "peak_sustained_2s": sust, "peak_proxy_2x_bf16": 2.0 * measured_peaks().get("bf16_tflops", 1590.0),
                    "nominal_peak": 4500.0, "frac_of_nominal": ach / 4500.0})
    else:
        imad_peak = h.microbench_imad(0, 3000)
        ach = rows * k * m * 128.0 / (t_mm * 1e-3)
        mmd.update({"bound": "imad", "achieved": ach / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD/s", "frac": ach / imad_peak,
                    "peak_source": "measured live: mad.lo.u32 micro-benchmark, all SMs"})
    out["fr_matmul"] = mmd
    # the full-width engine on the same shape with uniformly random canonical operands (the generic contract)
    if rows * k * m <= (1 << 31):
        gen = torch.Generator(device=device)
        gen.manual_seed(5)
        ra = torch.randint(0, 1 << 62, (rows, k, 4), dtype=torch.int64, device=device, generator=gen)
        rb = torch.randint(0, 1 << 62, (k, m, 4), dtype=torch.int64, device=device, generator=gen)
        ra[..., 3] &= (1 << 60) - 1
        rb[..., 3] &= (1 << 60) - 1
        t_full = timed(lambda: h.fr_matmul_dev(ra, rb, c_tmp))
        eng_full = h.last_matmul_engine()
        burst0 = h.microbench_tensor_i8(0, 0.0) / 1e12
        achf = rows * k * m * 2048.0 / (t_full * 1e-3) / 1e12
        out["fr_matmul_full_width_operands"] = {
            "kernel": "fr_matmul_tc_kernel<TcFull> + split kernels (uniform random full-width Fr operands)", "engine": eng_full,
            "ms": t_full, "bound": "tensor", "achieved": achf, "peak": burst0, "unit": "TOP/s (8-bit)", "frac": achf / burst0,
            "fr_mul_adds_per_s": rows * k * m / (t_full * 1e-3), "frac_of_nominal": achf / 4500.0}
        del ra, rb
    # north_star quotes the mat-mul against the INT32-IMAD roofline: the IMAD (Karatsuba) engine on the same operands
    if rank == 0 and not args.quick and rows * k * m <= (1 << 31):
        h.tune("matmul_tc", 0)
        try:
            imad_peak = h.microbench_imad(0, 3000)
            t_imad = timed(lambda: h.fr_matmul_dev(a_slab, b_dev, c_tmp), flush=False)
            out["fr_matmul_imad_engine"] = {
                "kernel": "fr_matmul_kara_kernel (the engine north_star describes; serves small products, cross-checks the tensor "
                          "engines)", "ms": t_imad, "bound": "imad", "achieved": rows * k * m * 128.0 / (t_imad * 1e-3) / 1e12,
                "peak": imad_peak / 1e12, "unit": "T IMAD/s", "frac": rows * k * m * 128.0 / (t_imad * 1e-3) / imad_peak,
                "algorithmic": "SURVEY 8(d): 128 IMAD-pipe slots per Fr mul-add (8x8-limb schoolbook)",
                "same_bytes_as_tensor_engine": bool((c_tmp == bufs["c_s"]).all().item())}
        finally:
            h.tune("matmul_tc", -1)
    del c_tmp
    # K4: rescale witnesses of this rank's slab: 32 * (1 + 1 + W) bytes per element (read c_s, write q, write W witnesses)
    t_rs = timed(lambda: h.rescale_witness_dev(bufs["c_s"], rows * m, P_BITS, LOOKUP_BITS, bufs["q"], bufs["wit"]))
    out["rescale_kernel"] = {"kernel": "rescale_tma_kernel (256-byte-aligned TMA stores of the witness stream; rescale_kernel for other layouts)",
                             "ms": t_rs, "bound": "hbm", "bytes": rows * m * 32.0 * (2 + W),
                             "achieved": rows * m * 32.0 * (2 + W) / (t_rs * 1e-3) / 1e9,
                             "algorithmic": f"32 B x (1 read + 1 quotient + W={W} witnesses) per element (SURVEY 8d)"}
    # K2: one Freivalds mat-vec with every running sum (C.v of this rank's slab): 64 B per multiply-add
    pcv_tmp, tot_tmp = fr(rows, m), fr(rows)
    t_mv = timed(lambda: h.mat_vec_prefix_dev(bufs["c_s"], bufs["powers"], pcv_tmp, tot_tmp))
    out["mat_vec_prefix"] = {"kernel": "mat_vec_prefix_*_kernel (C.v of verify_mul, every running sum)", "ms": t_mv, "bound": "hbm",
                             "bytes": rows * m * 64.0, "achieved": rows * m * 64.0 / (t_mv * 1e-3) / 1e9,
                             "elements_per_s": rows * m / (t_mv * 1e-3),
                             "algorithmic": "64 B per multiply-add (32 B element read + 32 B running sum written; v stays on chip)",
                             "same_bytes_as_step": bool((pcv_tmp == bufs["prefix_cv"]).all().item())}
    del pcv_tmp, tot_tmp
    if rank == 0 and not args.quick:
        # BASELINE configs[1]: ZkVector inner_product / norm / dist witnesses, 4096 vector pairs x 1024, P = 32
        batch, ln = 4096, 1024
        gen = torch.Generator(device=device)
        gen.manual_seed(1)
        xf = (torch.rand((batch, ln), dtype=torch.float64, device=device, generator=gen) - 0.5) * 20
        sf = (torch.rand((batch, ln), dtype=torch.float64, device=device, generator=gen) - 0.5) * 20
        x, s, o = fr(batch, ln), fr(batch, ln), fr(batch, ln)
        h.quantize_dev(xf, 32, x)
        h.quantize_dev(sf, 32, s)
        t_in = timed(lambda: h.zkvec_inner_prefix_dev(x, s, o))
        t_sub = timed(lambda: h.zkvec_sub_dev(s, x, o))
        cnt = batch * ln
        out["config1_zkvector"] = {
            "workload": "BASELINE configs[1]: 4096 vector pairs x 1024, PRECISION_BITS=32",
            "inner_product": {"kernel": "mat_vec_prefix_tile_kernel (per-row second operand)", "ms": t_in, "bound": "hbm",
                              "bytes": cnt * 96.0, "achieved": cnt * 96.0 / (t_in * 1e-3) / 1e9,
                              "algorithmic": "96 B per multiply-add (two 32 B operands read, 32 B running sum written)"},
            "qsub": {"kernel": "sub_kernel", "ms": t_sub, "bound": "hbm", "bytes": cnt * 96.0,
                     "achieved": cnt * 96.0 / (t_sub * 1e-3) / 1e9, "algorithmic": "96 B per element"},
            "fr_mul_adds_per_s": cnt / (t_in * 1e-3)}
        del xf, sf, x, s, o
        # BASELINE configs[2]: honest_prover_mat_mul + verify_mul witness, square N = 256, P = 32, lb = 19
        n2 = 256
        af = (torch.rand((n2, n2), dtype=torch.float64, device=device, generator=gen) - 0.5) * 4
        bf = (torch.rand((n2, n2), dtype=torch.float64, device=device, generator=gen) - 0.5) * 4
        a2, b2, c2 = fr(n2, n2), fr(n2, n2), fr(n2, n2)
        h.quantize_dev(af, 32, a2)
        h.quantize_dev(bf, 32, b2)
        g2 = torch.tensor([[0x1234567, 0x89ABCDEF, 0x13579BDF, 0x2468ACE]], dtype=torch.int64, device=device)
        fw = [fr(n2), fr(n2, n2), fr(n2, n2), fr(n2, n2), fr(n2), fr(n2), fr(n2)]

        def cfg2():
            h.fr_matmul_dev(a2, b2, c2)
            h.freivalds_witness_dev(a2, b2, c2, g2, *fw)

        t2 = timed(cfg2)
        ok2 = not bool(fw[4].any().item())
        u2 = n2 ** 3 + (n2 - 1) + 3 * n2 * n2
        out["config2_n256"] = {"workload": "BASELINE configs[2]: honest_prover_mat_mul + verify_mul witness, N=256, P=32, lb=19",
                               "ms": t2, "fr_mul_adds_per_s": u2 / (t2 * 1e-3), "matmul_engine": h.last_matmul_engine(),
                               "freivalds_diff_all_zero": ok2,
                               "note": "launch-latency sized (7 kernels, 12.6 MB moved): no roofline fraction is meaningful"}
    return out


def finish_hbm(kern: dict, hbm_peak: float, src: str) -> None:
    for v in kern.values():
        if isinstance(v, dict):
            if v.get("bound") == "hbm" and "achieved" in v:
                v["peak"], v["unit"], v["frac"], v["peak_source"] = hbm_peak, "GB/s", v["achieved"] / hbm_peak, src
            finish_hbm(v, hbm_peak, src)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--shape", type=str, default="", help="n,k,m of a rectangular job (default: square --n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=256, help="rows of the job the 1-thread cpu_baseline leg runs")
    ap.add_argument("--ref-rows", type=int, default=0, help="rows per step of the reference arm (default 2 x host threads)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (large rectangular jobs: pinned host memory)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step's kernels directly instead of replaying a CUDA graph")
    ap.add_argument("--matmul-small", type=int, default=None, help="tuning: 0 = never use the small-operand mat-mul engine")
    ap.add_argument("--tune", action="append", default=None, metavar="KEY=VAL", help="triage: per-handle tuning switch (repeatable)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of back-to-back steps for the sustained figure (0 = skip)")
    ap.add_argument("--multi-handle", action="store_true",
                    help="one process drives --gpus N devices through h2svd_multi (end-to-end only; no torchrun)")
    ap.add_argument("--quick", action="store_true", help="skip configs[1]/[2], the IMAD engine and the 2 s peak measurements")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rules: W >= 3
    if args.impl == "reference":
        run_reference(args)
    elif args.multi_handle:
        run_multi_handle(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
