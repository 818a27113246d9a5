"""Times the rescale kernel alone (N=1024, P=63, lb=19) -- developer tool, run under gpurun."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for (N, P, lb) in ((1024, 63, 19), (1024, 32, 19), (2048, 63, 19)):
    W = h.rescale_witness_count(P, lb)
    a_f = (torch.rand((N, N), dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 20
    a = torch.empty((N, N, 4), dtype=torch.int64, device="cuda")
    h.quantize_dev(a_f, P, a)
    cs = torch.empty_like(a); h.fr_matmul_dev(a, a, cs)
    q = torch.empty_like(a); wit = torch.empty((N * N, W, 4), dtype=torch.int64, device="cuda")
    ts = []
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); h.rescale_witness_dev(cs, N * N, P, lb, q, wit); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    b = N * N * 32 * (1 + W)
    chk = int(wit.sum().item()) ^ int(q.sum().item())
    print(f"chk={chk & 0xffffffff:08x} N={N} P={P} lb={lb} W={W}: best {min(ts[2:]):.4f} ms med {np.median(ts[2:]):.4f} ms -> {b/min(ts[2:])/1e6:.0f} GB/s", flush=True)
    del wit, q, cs, a
h.close()
