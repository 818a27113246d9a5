"""Times the running-sum (prefix) mat-vec / ZkVector inner-product kernels on a few shapes and checks the tile
kernel against the warp-per-segment kernel (run under gpurun):  python tools/matvec_bench.py [rows x len ...]"""
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


def timeit(fn, stream, reps=7, warm=3, inner=10):
    """`inner` back-to-back launches per event pair: a ~100 us kernel is otherwise timed together with the host's
    launch latency (the GPU idles between the start event and the first launch)."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(inner):
            fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return min(ts), float(np.median(ts))


def main():
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(4096, 1024), (8192, 512), (5000, 1000),
                                                                             (3000, 1024), (2048, 1024), (700, 4100)]
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h = pkg.Handle(0, stream.cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    for rows, ln in shapes:
        x, s = rand_fr(gen, rows, ln), rand_fr(gen, rows, ln)
        o, o_ref = torch.empty_like(x), torch.empty_like(x)
        h.tune("matvec_warp_kernel", True)
        h.zkvec_inner_prefix_dev(x, s, o_ref)
        h.tune("matvec_warp_kernel", False)
        best, med = timeit(lambda: h.zkvec_inner_prefix_dev(x, s, o), stream)
        h.sync()
        same = bool((o == o_ref).all())
        print(f"inner {rows}x{ln}: best {best*1e3:.1f} us med {med*1e3:.1f} us -> {rows*ln*96/(best*1e-3)/1e12:.2f} TB/s"
              f" same={same}", flush=True)
        # shared-vector mat-vec with row totals
        v = rand_fr(gen, ln)
        tot, tot_ref = torch.empty((rows, 4), dtype=torch.int64, device="cuda"), torch.empty((rows, 4), dtype=torch.int64, device="cuda")
        h.tune("matvec_warp_kernel", True)
        h.mat_vec_prefix_dev(x, v, o_ref, tot_ref)
        h.tune("matvec_warp_kernel", False)
        best, med = timeit(lambda: h.mat_vec_prefix_dev(x, v, o, tot), stream)
        h.sync()
        same = bool((o == o_ref).all() and (tot == tot_ref).all())
        print(f"mat-vec {rows}x{ln}: best {best*1e3:.1f} us -> {rows*ln*64/(best*1e-3)/1e12:.2f} TB/s same={same}", flush=True)
    h.close()


if __name__ == "__main__":
    main()
