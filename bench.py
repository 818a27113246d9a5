#!/usr/bin/env python
"""Headline benchmark: Fr mul-add/s for the mat-mul + Freivalds + rescale witness, N=1024, P=63,
LOOKUP_BITS=19 (BASELINE.json metric / configs[3]) on 1..8 B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA kernels)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # CPU restatement on the host cores

A "step" = one pass of the hot path over one (A, B) pair: honest_prover_mat_mul (C = A.B),
rescale_matrix witnesses of C, verify_mul (Freivalds) witnesses.  Rows of A/C are sharded over the
ranks ("strong" scaling: the job is fixed at N=1024), B is replicated, (B v) is all-gathered (NCCL).
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every key).
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
def cpu_sample(n: int, threads: int, mm_rows: int, rs_elems: int) -> dict:
    """Times the C oracle (CPU restatement of the reference loops) on a bounded sample of the same
    workload and extrapolates linearly to the full job.  The only place bench.py executes oracle/."""
    from oracle import corac
    a_f, b_f, gamma = make_inputs(n, n, n)
    mm_rows = min(mm_rows, n)
    a = corac.quantize(a_f[:mm_rows], P_BITS)
    b = corac.quantize(b_f, P_BITS)
    t0 = time.perf_counter()
    c = corac.field_mat_mul(a, b, threads=threads)                       # mm_rows of n rows
    t_mm = time.perf_counter() - t0
    t0 = time.perf_counter()
    powers = corac.gamma_powers(gamma.reshape(1, 4), n)
    corac.mat_vec_prefix(c, powers, threads=threads)                     # rows of C.v
    pbv = corac.mat_vec_prefix(b[:mm_rows], powers, threads=threads)     # rows of B.v
    t_fr_rows = time.perf_counter() - t0
    t0 = time.perf_counter()
    flat = np.ascontiguousarray(c.reshape(-1, 4)[:rs_elems])
    corac.rescale_witness(flat, P_BITS, LOOKUP_BITS, threads=threads)
    t_rs = time.perf_counter() - t0
    del pbv
    # extrapolate: mat-mul rows -> n rows; 2 sampled mat-vecs of mm_rows rows -> 3 mat-vecs of n rows
    full = t_mm * n / mm_rows + t_fr_rows * (3 * n) / (2 * mm_rows) + t_rs * (n * n) / flat.shape[0]
    units = n ** 3 + 3 * n * n + (n - 1)
    return {"value": units / full, "seconds_full_job_extrapolated": full,
            "sample": f"{mm_rows}/{n} rows of the mat-mul and of 2 of the 3 Freivalds mat-vecs, "
                      f"{flat.shape[0]}/{n * n} rescale elements; linear extrapolation",
            "t_sample_s": t_mm + t_fr_rows + t_rs}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.n
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(n, threads, mm_rows=min(n, max(64, 8 * threads)), rs_elems=min(n * n, 65536 * max(1, threads // 4)))
        if i >= args.warmup:
            vals.append(r)
            secs.append(r["seconds_full_job_extrapolated"])
    ms = float(np.mean(secs)) * 1e3
    units = n ** 3 + 3 * n * n + (n - 1)
    value = units / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32x8 (BN254 Fr, 256-bit modular)",
        "data": "synthetic",
        "config": {"workload": f"honest_prover_mat_mul + rescale_matrix + verify_mul witness, square N={n}, "
                               f"PRECISION_BITS={P_BITS}, LOOKUP_BITS={LOOKUP_BITS}", "n": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": vals[-1]["sample"] + " (C restatement of the reference loops; the Rust "
                                                        "reference cannot be built here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module("halo2-svd041_b200")
    wl = importlib.import_module("halo2-svd041_b200.workload")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)   # pinned buffers are first-touched on the GPU's own NUMA node
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        comm = dist.group.WORLD

    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    h = pkg.Handle(local_rank, stream.cuda_stream)
    # second handle on a second stream: the C-independent half of verify_mul overlaps the mat-mul
    side_stream = torch.cuda.Stream(device=device)
    h_side = pkg.Handle(local_rank, side_stream.cuda_stream)
    side = wl.SideStream(torch, h_side, side_stream, stream)

    n = k = m = args.n
    if args.shape:       # e.g. --shape 4096,2048,4096 = BASELINE configs[4] (not the headline configuration)
        n, k, m = (int(x) for x in args.shape.split(","))
    plan = wl.ShardPlan(n, k, m, world, rank)
    W = h.rescale_witness_count(P_BITS, LOOKUP_BITS)
    bufs = wl.alloc_buffers(torch, plan, W, device)
    r0, r1 = plan.rows
    b0, b1 = plan.brows
    # Measured: under a full-size mat-mul (2 CTAs/SM hold the whole register file) the side kernels only
    # displace mat-mul CTAs (+0.13 ms at 1 GPU); under the small slabs of 4-8 GPUs they fill idle SM time.
    # side-stream schedule: always with the tensor-core mat-mul engine (it leaves the integer pipe idle), else only for
    # small row slabs (the IMAD mat-mul engines saturate the pipe the mat-vecs need)
    tc_engine = k >= 32 and (r1 - r0) * k * m >= (1 << 18)
    overlap = tc_engine or (r1 - r0) * m < 512 * 1024
    # Measured at N=1024 on one GPU (step, ms): no side stream 1.170; C-independent mat-vecs under the tensor-core
    # mat-mul and C.v next to the rescale 1.116 (the mat-mul gives back most of what the mat-vecs save: they compete
    # for shared-memory bandwidth); all three mat-vecs next to the rescale 1.143 (its CTAs fill the SMs, the side
    # kernels only start as they drain); C-independent mat-vecs beside a hoisted operand split, the mat-mul after both
    # 1.162 (they take 130 us next to the split kernels).  With several ranks the second schedule (the one used) also
    # hides the all-gather latency.
    pre_under_matmul = True
    fused = tc_engine and args.fuse   # experimental: rescale witnesses from the mat-mul epilogue (measured not to pay)
    if fused:
        pkg.set_fuse_rescale(1)

    # ---- synthetic inputs: f64 matrices -> pinned host -> GPU quantization kernel (product path)
    a_f, b_f, gamma = make_inputs(n, k, m)
    a_host = torch.from_numpy(np.ascontiguousarray(a_f[r0:r1])).pin_memory()
    b_host = torch.from_numpy(np.ascontiguousarray(b_f)).pin_memory()
    a_dev_f, b_dev_f = a_host.to(device, non_blocking=True), b_host.to(device, non_blocking=True)
    h.quantize_dev(a_dev_f, P_BITS, bufs.a_slab)
    h.quantize_dev(b_dev_f, P_BITS, bufs.b)
    bufs.gamma.copy_(torch.from_numpy(gamma.view(np.int64).reshape(1, 4)))
    h.sync()
    # end-to-end leg: page-locked host buffers for the inputs and for every output of the fused C-ABI call
    def pinned_like(t):
        pb = pkg.PinnedBuffer(tuple(t.shape), np.uint64)
        return pb

    do_e2e = not args.no_e2e
    if not do_e2e:
        pinned_like = lambda t: pkg.PinnedBuffer((1, 4), np.uint64)   # noqa: E731  (placeholders, never used)
    pin_in = {"a": pinned_like(bufs.a_slab), "b": pinned_like(bufs.b)}
    if do_e2e:
        pin_in["a"].array[...] = bufs.a_slab.cpu().numpy().view(np.uint64)
        pin_in["b"].array[...] = bufs.b.cpu().numpy().view(np.uint64)
    out_shapes = {"c_s": (r1 - r0, m), "q": (r1 - r0, m), "wit": ((r1 - r0) * m, W), "powers": (m,),
                  "prefix_cv": (r1 - r0, m), "prefix_bv": (b1 - b0, m), "prefix_abv": (r1 - r0, k), "diff": (r1 - r0,),
                  "is_zero": (r1 - r0,), "inv": (r1 - r0,)}
    pin_out = {nm: pkg.PinnedBuffer((shp if do_e2e else (1,) * len(shp)) + (4,), np.uint64) for nm, shp in out_shapes.items()}
    host_out = {nm: pb.array for nm, pb in pin_out.items()}
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def device_step(times=None):
        # same schedule as wl.run_step(..., side=side), with phase events on the main stream
        e = [ev() for _ in range(4)] if times is not None else None
        if e: e[0].record(stream)
        if fused:
            # mat-mul + rescale witnesses in one launch (tensor-core engine, witnesses written from its epilogue); the
            # C-independent Freivalds mat-vecs under it on the side stream, C.v after it
            bv = side.run(lambda be: wl.step_freivalds_pre(be, plan, bufs, dist, comm))
            wl.step_matmul_rescale(h, plan, bufs, P_BITS, LOOKUP_BITS)
            if e: e[1].record(stream)
            if e: e[2].record(stream)
            side.join()
            wl.step_freivalds_post(h, plan, bufs, bv)
            if e:
                e[3].record(stream)
                times.append(e)
            return
        if overlap and pre_under_matmul:
            bv = side.run(lambda be: wl.step_freivalds_pre(be, plan, bufs, dist, comm))   # under the mat-mul
        wl.step_matmul(h, plan, bufs)
        if e: e[1].record(stream)
        if overlap:
            # the mat-vecs (integer-pipe bound) on the side stream next to the rescale kernel (HBM bound)
            side.run(lambda be: wl.step_freivalds_post(be, plan, bufs, bv))
            wl.step_rescale(h, plan, bufs, P_BITS, LOOKUP_BITS)
            if e: e[2].record(stream)
            side.join()
        else:
            wl.step_rescale(h, plan, bufs, P_BITS, LOOKUP_BITS)
            if e: e[2].record(stream)
            bv = wl.step_freivalds_pre(h, plan, bufs, dist, comm)
            wl.step_freivalds_post(h, plan, bufs, bv)
        if e:
            e[3].record(stream)
            times.append(e)

    def e2e_step():
        # ONE call of the reference-facing C ABI on host buffers: H2D of A (this rank's rows), B, gamma; mat-mul,
        # rescale and verify_mul witnesses in row slabs; D2H of every witness, overlapped slab by slab.  No
        # collective: every rank derives the k row totals of B.v itself and returns its own share of prefix_bv.
        h.zkmatrix_mul_witness(pin_in["a"].array, pin_in["b"].array, gamma, P_BITS, LOOKUP_BITS,
                               bv_rows=(b0, b1), out=host_out)

    # ---- device-resident timing ----
    sampler = ClockSampler(local_rank)   # runs through both timed regions (device-resident and end-to-end)
    sampler.start()
    for _ in range(args.warmup):
        device_step()
    barrier()
    step_events = []
    launches_before = h.launch_count + h_side.launch_count
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_buf.fill_(1)              # L2 flush between timed iterations (outside the event brackets)
        device_step(step_events)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gpu_launches = h.launch_count + h_side.launch_count - launches_before
    tot = [e[0].elapsed_time(e[3]) for e in step_events]
    t_mm = [e[0].elapsed_time(e[1]) for e in step_events]
    t_rs = [e[1].elapsed_time(e[2]) for e in step_events]
    t_fr = [e[2].elapsed_time(e[3]) for e in step_events]
    my_ms = torch.tensor([sum(tot) / len(tot), float(np.mean(t_mm)), float(np.mean(t_rs)), float(np.mean(t_fr))],
                         dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(my_ms, op=dist.ReduceOp.MAX)
    ms_step, ms_mm, ms_rs, ms_fr = [float(x) for x in my_ms.cpu()]

    # ---- end-to-end timing (host buffers, copies inside the timed region) ----
    for _ in range(max(1, min(args.warmup, 3)) if do_e2e else 0):
        e2e_step()
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps if do_e2e else 0):
        e2e_step()             # synchronous: returns when the step's results are on the host
    t_mine = (time.perf_counter() - t0) / e2e_steps
    barrier()
    e2e_wall = (time.perf_counter() - t0) / e2e_steps
    e2e_ms_t = torch.tensor([t_mine * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_ms_t, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms_t.cpu()[0])
    clocks = sampler.stop()
    h2d = int(pin_in["a"].array.nbytes + pin_in["b"].array.nbytes + 32)
    d2h = sum(int(host_out[x].nbytes) for x in host_out)
    bytes_t = torch.tensor([h2d, d2h], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(bytes_t, op=dist.ReduceOp.SUM)
    h2d_all, d2h_all = [int(x) for x in bytes_t.cpu()]

    # ---- sanity: the timed buffers hold a correct witness (honest product => every diff is zero)
    ok = bool((bufs.diff == 0).all().item()) and (not do_e2e or not host_out["diff"].any())
    # both legs computed the same witness: the fused host call and the device-resident building blocks agree
    for nm, dev_t in ((("c_s", bufs.c_slab), ("q", bufs.q_slab), ("prefix_cv", bufs.prefix_cv), ("prefix_abv", bufs.prefix_abv),
                       ("wit", bufs.wit_slab)) if do_e2e else ()):
        ok = ok and bool((torch.from_numpy(host_out[nm].view(np.int64)).to(device) == dev_t).all().item())
    if not ok:
        raise SystemExit("bench: witness check failed -- refusing to report a number for wrong results")

    units = plan.mul_adds()
    value = units / (ms_step * 1e-3)
    line = None
    if rank == 0:
        peaks = measured_peaks()
        rows = r1 - r0
        engine = pkg.last_matmul_engine()
        if engine == "tensor":
            # roofline of the dominant kernel (fr_matmul_tc_kernel): tensor-pipe bound.  Algorithmic work = 1024 u8
            # multiply-adds (2048 ops) per Fr mul-add: 32 x 32 byte-plane products (DESIGN.md "K1t").  Peak = twice the
            # measured dense bf16 rate (8-bit operands run at twice the 16-bit rate on the same pipe).
            bf16 = peaks.get("bf16_tflops")
            tensor_peak = 2.0 * (bf16 if bf16 else 1590.0)
            ops = rows * k * m * 2048.0
            achieved_t = ops / (ms_mm * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "fr_matmul_tc_kernel (tcgen05.mma kind::i8, u8 x u8 -> s32 in TMEM) + the two "
                                                     "byte-plane split kernels, which are inside the timed phase",
                        "achieved": achieved_t, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved_t / tensor_peak,
                        "op": "u8 multiply-add = 2 ops; 1024 multiply-adds per Fr mul-add",
                        "peak_source": ("2 x MEASURED_PEAKS.json bf16_tflops (burst: the kernel runs ~0.6 ms per step); the cuBLAS "
                                        "bf16 GEMM behind that figure reaches 73 % of the nominal pipe rate, so frac can exceed 1"
                                        if bf16 else "2 x 1590 TF/s of fallback (B200_PROFILING.md)"),
                        "frac_of_nominal": achieved_t / 4500.0, "nominal_peak": 4500.0,
                        "traffic": traffic_from_profile("fr_matmul_tc_kernel", n) if world == 1 else None,
                        "traffic_source": "profiles/traffic.json (ncu --set full capture of this command, bytes per launch)"}
        else:
            # IMAD engines (small or very short-k products): integer-pipe bound.  Algorithmic work = 128 IMAD-pipe slots
            # per Fr mul-add; peak = IMAD issue rate measured live.
            imad_peak = h.microbench_imad(0, 3000)
            achieved = rows * k * m * 128.0 / (ms_mm * 1e-3)
            roofline = {"bound": "imad", "kernel": "fr_matmul_kara_kernel" if engine == "karatsuba" else "fr_matmul_kernel",
                        "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "T IMAD/s", "frac": achieved / imad_peak,
                        "peak_source": "measured live: mad.lo.u32 micro-benchmark, all SMs (h2svd_microbench_imad kind 0)",
                        "algorithmic": "SURVEY 8(d): 128 IMAD-pipe slots per Fr mul-add (8x8-limb schoolbook); the Karatsuba "
                                       "kernel executes 108 (profiles/r01c_ncu_full_summary.md), so frac can exceed 1",
                        "traffic": traffic_from_profile("fr_matmul_kara_kernel", n) if world == 1 else None,
                        "traffic_source": "profiles/traffic.json"}
        rs_bytes = rows * m * 32.0 * (1 + W)
        fr_bytes = 2.0 * 32.0 * (rows * m + rows * k + (0 if overlap else (b1 - b0) * m))   # mat-vecs inside the timed phase
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": METRIC if (n, k, m) == (N_DEFAULT,) * 3 else f"Fr mul-add/s for mat-mul+Freivalds+rescale witness, {n}x{k}x{m}",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x8 (BN254 Fr, 256-bit modular)", "data": "synthetic",
            "config": {"workload": (f"honest_prover_mat_mul + rescale_matrix + verify_mul witness, square N={n}, "
                                    f"PRECISION_BITS={P_BITS}, LOOKUP_BITS={LOOKUP_BITS} (BASELINE configs[3] + Freivalds)")
                       if n == k == m else
                       (f"honest_prover_mat_mul + rescale_matrix + verify_mul witness, {n}x{k} . {k}x{m}, "
                        f"PRECISION_BITS={P_BITS}, LOOKUP_BITS={LOOKUP_BITS}"), "n": n, "k": k, "m": m, "sharding": f"rows of A/C over {world} rank(s), B replicated, "
                                   "(B v) all-gathered", "l2": "256 MiB flush write between timed steps",
                       "inputs": "input-creator.py distribution, seeded, quantized on the GPU"},
            "phase_ms": {"fr_matmul": ms_mm, "rescale": ms_rs, "freivalds_after_matmul": ms_fr, "freivalds_pre_overlapped_with_matmul": bool(overlap),
                         "rescale_fused_into_matmul_epilogue": bool(fused)},
            "roofline": roofline, "matmul_engine": engine,
            "roofline_hbm": {
                # fused: the witness stream is written during the whole mat-mul launch (phase "fr_matmul")
                "rescale": {"bound": "hbm", "achieved": rs_bytes / ((ms_mm if fused else ms_rs) * 1e-3) / 1e9, "peak": hbm_peak,
                            "unit": "GB/s", "frac": rs_bytes / ((ms_mm if fused else ms_rs) * 1e-3) / 1e9 / hbm_peak,
                            "timed_over": "the fused mat-mul + rescale launch" if fused else "rescale_kernel"},
                "freivalds": {"bound": "hbm", "achieved": fr_bytes / (ms_fr * 1e-3) / 1e9, "peak": hbm_peak,
                              "unit": "GB/s", "frac": fr_bytes / (ms_fr * 1e-3) / 1e9 / hbm_peak},
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"},
            "e2e": {"value": units / (e2e_ms * 1e-3) if do_e2e else None, "unit": UNIT, "ms_per_step": e2e_ms if do_e2e else None,
                    "wall_ms_per_step": e2e_wall * 1e3, "h2d_bytes_per_step": h2d_all,
                    "d2h_bytes_per_step": d2h_all,
                    "api": "h2svd_zkmatrix_mul_witness (C ABI, pinned host buffers in/out, slab-pipelined D2H) per rank",
                    "numa_node_rank0": numa},
            "gpu_launches": int(gpu_launches) * world, "clocks": clocks, "wall_s_timed_region": t_wall,
            "verified": "Freivalds diff == 0; fused host call and device building blocks byte-identical",
        }
        if world == 1 and engine == "tensor" and rows * k * m <= (1 << 31):
            # north_star quotes the mat-mul against the INT32-IMAD roofline: time the IMAD (Karatsuba) engine on the same
            # operands too (outside the timed region; three launches) and report it in that accounting
            pkg.set_matmul_tc(0)
            try:
                imad_peak = h.microbench_imad(0, 3000)
                c_ref = torch.empty_like(bufs.c_slab)
                h.fr_matmul_dev(bufs.a_slab, bufs.b, c_ref)
                ts = []
                for _ in range(3):
                    e0, e1 = ev(), ev()
                    e0.record(stream)
                    h.fr_matmul_dev(bufs.a_slab, bufs.b, c_ref)
                    e1.record(stream)
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                t_imad = float(np.mean(ts))
                line["roofline_imad_engine"] = {
                    "bound": "imad", "kernel": "fr_matmul_kara_kernel (the engine north_star describes; kept for small products "
                                               "and as the cross-check of the tensor-core engine)",
                    "ms": t_imad, "achieved": rows * k * m * 128.0 / (t_imad * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                    "unit": "T IMAD/s", "frac": rows * k * m * 128.0 / (t_imad * 1e-3) / imad_peak,
                    "algorithmic": "SURVEY 8(d): 128 IMAD-pipe slots per Fr mul-add (8x8-limb schoolbook); the Karatsuba kernel "
                                   "executes 108, so frac can exceed 1",
                    "peak_source": "measured live: mad.lo.u32 micro-benchmark, all SMs",
                    "same_bytes_as_tensor_engine": bool((c_ref == bufs.c_slab).all().item()),
                    "speedup_of_tensor_engine": t_imad / ms_mm}
                del c_ref
            finally:
                pkg.set_matmul_tc(-1)
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_sample(n, 1, mm_rows=256, rs_elems=262144) if n == k == m else None   # ~15 s of 1-thread CPU work
            if cb is None:
                cb = {"value": None, "sample": "square jobs only", "t_sample_s": 0.0}
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": cb["sample"] + f" ({cb['t_sample_s']:.1f} s of CPU work, 1 thread: "
                                                             "the reference is single-threaded)"}
    h_side.close()
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--shape", type=str, default="", help="n,k,m of a rectangular job (default: square --n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fuse", action="store_true", help="experimental fused mat-mul + rescale launch (tuning)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (large rectangular jobs: pinned host memory)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rules: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
