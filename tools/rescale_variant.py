"""Times the N = 1024 rescale with whatever library H2SVD_LIB points at; prints a checksum of the witness stream."""
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
def fr(*s): return torch.zeros(s + (4,), dtype=torch.int64, device=dev)
x = (torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
y = (torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
a, b, c = fr(n, n), fr(n, n), fr(n, n)
h.quantize_dev(x, 63, a); h.quantize_dev(y, 63, b); h.fr_matmul_dev(a, b, c); h.sync()
W = h.rescale_witness_count(63, 19)
q, wit = fr(n, n), fr(n * n, W)
if os.environ.get("RESCALE_CH"):
    h.tune("rescale_ch", int(os.environ["RESCALE_CH"]))
for kv in filter(None, os.environ.get("TUNE", "").split(",")):
    key, val = kv.split("=")
    h.tune(key, int(val))
ts = []
for i in range(9):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); h.rescale_witness_dev(c, n * n, 63, 19, q, wit); e1.record(stream); e1.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
chk = int(wit.sum().item()) ^ int(q.sum().item())
print(f"{os.environ.get('H2SVD_LIB', 'default')} rescale_ch={os.environ.get('RESCALE_CH', 'default')} {os.environ.get('TUNE', '')}: rescale median {np.median(ts):.1f} us min {min(ts):.1f} us checksum {chk & 0xffffffffffff:x}")
