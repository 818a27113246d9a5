"""K2 sweep: several-warps-per-row kernel with 2 / 4 / 8 warps per row, 1024 x 1024 and 128 x 1024."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("halo2-svd041_b200")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(3)
n = 1024
a = torch.randint(0, 1 << 62, (n, n, 4), dtype=torch.int64, device=dev, generator=gen); a[..., 3] &= (1 << 60) - 1
v = torch.randint(0, 1 << 62, (n, 4), dtype=torch.int64, device=dev, generator=gen); v[..., 3] &= (1 << 60) - 1
out = torch.zeros((n, n, 4), dtype=torch.int64, device=dev); tot = torch.zeros((n, 4), dtype=torch.int64, device=dev)
def timed(fn):
    ts = []
    for i in range(9):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    return round(float(np.median(ts)), 1)
ref = None
for rows in (1024, 128):
    for segs in (8, 4, 2):
        h.tune("matvec_seg", 1); h.tune("matvec_segs", segs)
        o = out[:rows]; o.fill_(-1)
        t = timed(lambda: h.mat_vec_prefix_dev(a[:rows], v, o, tot[:rows]))
        if ref is None: ref = o.clone()
        print(f"mat_vec_prefix {rows}x1024, {segs} warps per row: {t} us same={bool((o == ref[:rows]).all().item())}")
h.close()
