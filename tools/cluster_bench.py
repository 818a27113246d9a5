"""Times the small-operand tensor-core mat-mul with and without cluster multicast of the A planes (developer tool, run
under gpurun): tile widths 8/16/24/28 x cluster sizes 1/2 on the shapes of the sharded jobs; every variant must produce
the same bytes."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("halo2-svd041_b200")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
for kind, label in ((0, "u8 N256"), (1, "s8 N256"), (2, "s8 N144"), (3, "s8 N80")):
    print(f"tensor pipe {label}: {h.microbench_tensor_i8(kind) / 1e15:.3f} P op/s", flush=True)
shapes = [(1024, 1024, 1024), (512, 1024, 1024), (256, 1024, 1024), (128, 1024, 1024), (4096, 2048, 4096), (512, 2048, 4096)]
if os.environ.get("SHAPES"):
    shapes = [tuple(int(x) for x in sh.split("x")) for sh in os.environ["SHAPES"].split(",")]
for (n, k, m) in shapes:
    af = (torch.rand((n, k), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
    bf = (torch.rand((k, m), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
    a = torch.zeros((n, k, 4), dtype=torch.int64, device=dev)
    b = torch.zeros((k, m, 4), dtype=torch.int64, device=dev)
    c = torch.zeros((n, m, 4), dtype=torch.int64, device=dev)
    h.quantize_dev(af, 63, a)
    h.quantize_dev(bf, 63, b)
    ref = None
    for width in (0, 28, 24, 16, 8):
        for cluster in (0, 2):
            h.tune("matmul_small_width", width)
            h.tune("matmul_cluster", cluster)
            ts = []
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                h.fr_matmul_dev(a, b, c)
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            h.sync()
            if ref is None:
                ref = c.clone()
            same = bool((ref == c).all())
            c.zero_()
            t = min(ts[2:])
            print(f"{n}x{k}x{m} width {width:2d} cluster {cluster}: {t * 1e3:8.1f} us  {n * k * m / t / 1e9:7.2f} T mul-add/s  "
                  f"engine={h.last_matmul_engine()} same={same}", flush=True)
h.tune("matmul_small_width", 0)
h.tune("matmul_cluster", 0)
h.close()
