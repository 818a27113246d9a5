"""Checks the tensor-core mat-mul engine against the IMAD (Karatsuba / schoolbook) engines and times it
(run under gpurun):  python tools/tc_check.py [n,k,m ...]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")


def rand_fr(gen, *shape):
    t = torch.randint(-(1 << 63), (1 << 63) - 1, shape + (4,), dtype=torch.int64, device="cuda", generator=gen)
    t[..., 3] &= (1 << 60) - 1
    return t


def main():
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [
        (128, 128, 8), (128, 32, 8), (5, 7, 3), (16, 16, 16), (32, 32, 32), (64, 64, 64), (128, 128, 128), (130, 300, 20),
        (256, 256, 256), (300, 1500, 77), (1000, 8, 1000), (128, 1024, 1024), (512, 1024, 1024), (1024, 1024, 1024), (2048, 2048, 2048)]
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h = pkg.Handle(0, stream.cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(11)
    for n, k, m in shapes:
        a, b = rand_fr(gen, n, k), rand_fr(gen, k, m)
        if n >= 4 and k >= 4 and m >= 2:  # edge values: zero, one (Montgomery R mod r), r-1 patterns already canonical
            a[0, 0] = 0
            b[1, 1] = 0
            a[1, :] = 0
        ref, out = torch.empty((n, m, 4), dtype=torch.int64, device="cuda"), torch.empty((n, m, 4), dtype=torch.int64, device="cuda")
        h.tune("matmul_tc", 0)
        h.fr_matmul_dev(a, b, ref)
        h.sync()
        h.tune("matmul_tc", 1)
        out.fill_(-1)
        h.fr_matmul_dev(a, b, out)
        h.sync()
        same = bool((ref == out).all())
        nbad = int((ref != out).any(dim=-1).sum())
        msg = f"{n}x{k}x{m}: same={same} bad={nbad}/{n*m}"
        if not same:
            bad = (ref != out).any(dim=-1).nonzero()[:6].tolist()
            msg += f" first bad (i,j): {bad}"
        def bench(inner=4):
            ts = []
            for it in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(inner):
                    h.fr_matmul_dev(a, b, out)
                e1.record(stream)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) / inner)
            return min(ts)
        best = bench()
        h.tune("matmul_tc", 0)
        imad = bench()
        msg += (f"  tc {best:.4f} ms -> {n*k*m/(best*1e-3)/1e9:.1f} G mul-add/s ({2*1024*n*k*m/(best*1e-3)/1e12:.0f} T int8 op/s)"
                f"  imad {imad:.4f} ms")
        print(msg, flush=True)
    h.tune("matmul_tc", 0)
    h.close()


if __name__ == "__main__":
    main()
