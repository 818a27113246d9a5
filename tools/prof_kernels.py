"""Short driver for ncu captures of the HBM-side kernels (run under gpurun):
rescale (N=1024, P=63, lb=19), mat-vec prefix (1024x1024, shared v), ZkVector inner prefix (4096x1024)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")


def rand_fr(gen, *shape):
    t = torch.randint(-(1 << 63), (1 << 63) - 1, shape + (4,), dtype=torch.int64, device="cuda", generator=gen)
    t[..., 3] &= (1 << 60) - 1
    return t


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h = pkg.Handle(0, stream.cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    N, P, lb = 1024, 63, 19
    W = h.rescale_witness_count(P, lb)
    a_f = (torch.rand((N, N), dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 20
    a = torch.empty((N, N, 4), dtype=torch.int64, device="cuda")
    h.quantize_dev(a_f, P, a)
    cs = torch.empty_like(a)
    h.fr_matmul_dev(a, a, cs)
    q = torch.empty_like(a)
    wit = torch.empty((N * N, W, 4), dtype=torch.int64, device="cuda")
    v = rand_fr(gen, N)
    pcv = torch.empty_like(a)
    tot = torch.empty((N, 4), dtype=torch.int64, device="cuda")
    x, s = rand_fr(gen, 4096, 1024), rand_fr(gen, 4096, 1024)
    o = torch.empty_like(x)
    for _ in range(reps):
        h.rescale_witness_dev(cs, N * N, P, lb, q, wit)
        h.mat_vec_prefix_dev(cs, v, pcv, tot)
        h.zkvec_inner_prefix_dev(x, s, o)
        h.zkvec_sub_dev(s, x, o)
    h.sync()
    h.close()
    print("ok")


if __name__ == "__main__":
    main()
