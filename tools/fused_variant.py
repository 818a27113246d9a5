"""mat-mul + rescale at N = 1024 (quantized operands): two kernels vs the fused epilogue, and the timeline of the fused kernel."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("halo2-svd041_b200")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rows = int(sys.argv[2]) if len(sys.argv) > 2 else n
def fr(*s): return torch.zeros(s + (4,), dtype=torch.int64, device=dev)
x = (torch.rand((rows, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
y = (torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
a, b = fr(rows, n), fr(n, n)
h.quantize_dev(x, 63, a); h.quantize_dev(y, 63, b); h.sync()
W = h.rescale_witness_count(63, 19)
outs = {}
for fuse in (0, 1):
    h.tune("fuse_rescale", fuse)
    c, q, wit = fr(rows, n), fr(rows, n), fr(rows * n, W)
    ts = []
    for i in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); h.fr_matmul_rescale_dev(a, b, c, 63, 19, q, wit); e1.record(stream); e1.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    h.sync()
    outs[fuse] = (c, q, wit)
    print(f"{rows}x{n}x{n} fuse={fuse}: mat-mul + rescale median {np.median(ts):.1f} us min {min(ts):.1f} us engine={h.last_matmul_engine()}")
same = all(bool((u == v).all().item()) for u, v in zip(outs[0], outs[1]))
print("fused output byte-identical to the two-kernel output:", same)
h.tune("fuse_rescale", 0)
h.close()
