"""One launch of the N = 1024 rescale per running-sum variant (Montgomery step + modular add, then fr::SmallSum): the
target of an ncu capture (-k regex:rescale_)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("halo2-svd041_b200")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
h = pkg.Handle(0, stream.cuda_stream)
gen = torch.Generator(device=dev); gen.manual_seed(3)
n = 1024
def fr(*s): return torch.zeros(s + (4,), dtype=torch.int64, device=dev)
x = (torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
y = (torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
a, b, c = fr(n, n), fr(n, n), fr(n, n)
h.quantize_dev(x, 63, a); h.quantize_dev(y, 63, b); h.fr_matmul_dev(a, b, c); h.sync()
W = h.rescale_witness_count(63, 19)
q, wit = fr(n, n), fr(n * n, W)
for fast in (0, 1):
    h.tune("rescale_fast_sums", fast)
    h.rescale_witness_dev(c, n * n, 63, 19, q, wit)
    h.sync()
h.tune("rescale_fast_sums", 0)
h.close()
