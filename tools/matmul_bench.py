"""Times the field mat-mul on the shapes of the row-sharded N=1024 job (developer tool, run under gpurun):
plain one-CTA-per-tile schedule vs stream-K."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
def rand_fr(*shape):
    t = torch.randint(-(1 << 63), (1 << 63) - 1, shape + (4,), dtype=torch.int64, device="cuda", generator=gen)
    t[..., 3] &= (1 << 60) - 1
    return t
shapes = [(1024, 1024, 1024), (512, 1024, 1024), (256, 1024, 1024), (128, 1024, 1024), (256, 256, 256), (512, 2048, 4096)]
if os.environ.get("SHAPES"):
    shapes = [tuple(int(x) for x in sh.split("x")) for sh in os.environ["SHAPES"].split(",")]
variants = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
kara_modes = [int(x) for x in os.environ.get("KARA", "1,2,3").split(",") if x]
for (n, k, m) in shapes:
    a, b = rand_fr(n, k), rand_fr(k, m)
    c = torch.empty((n, m, 4), dtype=torch.int64, device="cuda")
    ref = None
    for variant in variants:
        h.tune("matmul_variant", variant)
        modes = [("sk", 0), ("sk", 1), ("sk", -1)] + [("kara", x) for x in kara_modes]
        for kind, sk in modes:
            h.tune("matmul_karatsuba", sk if kind == "kara" else 0)
            h.tune("matmul_streamk", sk if kind == "sk" else -1)
            ts = []
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); h.fr_matmul_dev(a, b, c); e1.record(stream); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            h.sync()
            if ref is None:
                ref = c.clone()
            same = bool((ref == c).all())
            t = min(ts[2:])
            print(f"{n}x{k}x{m} variant {variant} {kind}={sk:2d}: {t:.4f} ms  {n*k*m/t/1e6:.1f} G mul-add/s  same={same}", flush=True)
h.tune("matmul_streamk", -1); h.tune("matmul_variant", 0); h.tune("matmul_karatsuba", -1)
h.close()
