"""Triage: phase timeline of CTA 0 of the tensor-core mat-mul kernels (small-operand and full-width engines)."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("halo2-svd041_b200")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rows = int(sys.argv[2]) if len(sys.argv) > 2 else n
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
gen = torch.Generator(device=dev)
gen.manual_seed(1)


def fr(*shape):
    return torch.zeros(shape + (4,), dtype=torch.int64, device=dev)


def run(label, a, b):
    c = fr(a.shape[0], b.shape[1])
    for _ in range(3):
        h.fr_matmul_dev(a, b, c)
    h.sync()
    h.matmul_timeline(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h.fr_matmul_dev(a, b, c)
    e1.record(stream)
    e1.synchronize()
    tl = h.matmul_timeline(False).astype(np.int64)
    print(f"== {label}: {a.shape[0]}x{a.shape[1]}x{b.shape[1]} engine={h.last_matmul_engine()} whole call {e0.elapsed_time(e1) * 1e3:.1f} us")
    t0 = tl[tl > 0].min()
    names = ["mma_start", "mma_issued", "epi_sees_acc", "tmem_read+carry", "zeroed+released", "field+stores", "tma_first", "tma_last"]
    print("   round " + " ".join(f"{x:>16s}" for x in names) + "   (us since first stamp)")
    for r in range(16):
        if not tl[r].any():
            continue
        print(f"   {r:5d} " + " ".join(f"{(v - t0) / 1e3:16.2f}" if v else f"{'-':>16s}" for v in tl[r]))


af = (torch.rand((rows, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
bf = (torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
a, b = fr(rows, n), fr(n, n)
h.quantize_dev(af, 63, a)
h.quantize_dev(bf, 63, b)
run("quantized operands", a, b)
ra = torch.randint(0, 1 << 62, (rows, n, 4), dtype=torch.int64, device=dev, generator=gen)
rb = torch.randint(0, 1 << 62, (n, n, 4), dtype=torch.int64, device=dev, generator=gen)
ra[..., 3] &= (1 << 60) - 1
rb[..., 3] &= (1 << 60) - 1
run("full-width operands", ra, rb)
h.close()
