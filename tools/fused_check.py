"""Checks the fused mat-mul + rescale launch (witnesses written from the tensor-core epilogue) against the two separate
kernels, and times both (run under gpurun):  python tools/fused_check.py [n,k,m ...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")


def main():
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [
        (64, 64, 64), (130, 300, 20), (100, 1500, 77), (128, 1024, 1024), (1024, 1024, 1024)]
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h = pkg.Handle(0, stream.cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    for P, lb in ((63, 19), (32, 12)):
        W = h.rescale_witness_count(P, lb)
        for n, k, m in shapes:
            af = (torch.rand((n, k), dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 20
            bf = (torch.rand((k, m), dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 20
            a = torch.empty((n, k, 4), dtype=torch.int64, device="cuda")
            b = torch.empty((k, m, 4), dtype=torch.int64, device="cuda")
            h.quantize_dev(af, P, a)
            h.quantize_dev(bf, P, b)
            outs = []
            times = []
            for fuse in (0, 1):
                h.tune("fuse_rescale", fuse)
                c = torch.full((n, m, 4), -1, dtype=torch.int64, device="cuda")
                q = torch.full((n, m, 4), -1, dtype=torch.int64, device="cuda")
                wit = torch.full((n * m, W, 4), -1, dtype=torch.int64, device="cuda")
                h.fr_matmul_rescale_dev(a, b, c, P, lb, q, wit)
                h.sync()
                ts = []
                for _ in range(4):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    h.fr_matmul_rescale_dev(a, b, c, P, lb, q, wit)
                    e1.record(stream)
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                outs.append((c, q, wit))
                times.append(min(ts))
            same = all(bool((x == y).all()) for x, y in zip(outs[0], outs[1]))
            nbad = int((outs[0][2] != outs[1][2]).any(dim=-1).sum())
            print(f"P={P} lb={lb} W={W} {n}x{k}x{m}: same={same} bad_wit={nbad} engine={h.last_matmul_engine()} "
                  f"separate {times[0]:.4f} ms fused {times[1]:.4f} ms", flush=True)
            del outs
    h.tune("fuse_rescale", 0)
    h.close()


if __name__ == "__main__":
    main()
