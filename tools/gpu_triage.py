"""GPU triage / tuning run (developer tool, run under gpurun): integer-pipe micro-benchmarks,
mat-mul tile-variant timings, per-kernel timings at the BASELINE sizes.  Writes gpurun_out/triage.json."""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")


def rand_fr(gen, *shape):
    t = torch.randint(-(1 << 63), (1 << 63) - 1, shape + (4,), dtype=torch.int64, device="cuda", generator=gen)
    t[..., 3] &= (1 << 60) - 1
    return t


def timeit(fn, stream, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    out = {}
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h = pkg.Handle(0, stream.cuda_stream)
    out["sm_count"] = h.sm_count
    names = {0: "imad_lo", 1: "imad_wide", 2: "wide_chain", 3: "mulacc_8x8"}
    for kind, name in names.items():
        v = h.microbench_imad(kind, 4000)
        out[f"mb_{name}_Tops"] = v / 1e12
        print(f"microbench {name}: {v/1e12:.3f} T mul-instr/s", flush=True)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    for N in (256, 1024):
        a, b = rand_fr(gen, N, N), rand_fr(gen, N, N)
        c = torch.empty_like(a)
        ref = None
        h.tune("matmul_karatsuba", 0)          # tile variants of the schoolbook engine
        for variant in (0, 1, 2, 3):
            h.tune("matmul_variant", variant)
            best, med = timeit(lambda: h.fr_matmul_dev(a, b, c), stream)
            h.sync()
            if ref is None:
                ref = c.clone()
            same = bool((ref == c).all())
            rate = N ** 3 / (best * 1e-3)
            out[f"matmul_N{N}_v{variant}"] = dict(ms_best=best, ms_med=med, gmuladd_s=rate / 1e9, same_as_v0=same)
            print(f"matmul N={N} variant {variant}: best {best:.3f} ms med {med:.3f} ms -> {rate/1e9:.1f} G mul-add/s"
                  f" same={same}", flush=True)
        h.tune("matmul_variant", 0)
        h.tune("matmul_karatsuba", -1)
        best, med = timeit(lambda: h.fr_matmul_dev(a, b, c), stream)
        h.sync()
        out[f"matmul_N{N}_default"] = dict(ms_best=best, ms_med=med, gmuladd_s=N ** 3 / (best * 1e-3) / 1e9, same_as_v0=bool((ref == c).all()))
        print(f"matmul N={N} default (Karatsuba engine): best {best:.3f} ms -> {N**3/(best*1e-3)/1e9:.1f} G mul-add/s same={bool((ref == c).all())}", flush=True)
    # Freivalds + rescale at N=1024
    N, P, lb = 1024, 63, 19
    a, b = rand_fr(gen, N, N), rand_fr(gen, N, N)
    cs = torch.empty_like(a)
    h.fr_matmul_dev(a, b, cs)
    gamma = rand_fr(gen, 1)
    W = h.rescale_witness_count(P, lb)
    bufs = dict(powers=torch.empty((N, 4), dtype=torch.int64, device="cuda"),
                pcv=torch.empty_like(a), pbv=torch.empty_like(a), pabv=torch.empty_like(a),
                diff=torch.empty((N, 4), dtype=torch.int64, device="cuda"),
                z=torch.empty((N, 4), dtype=torch.int64, device="cuda"),
                inv=torch.empty((N, 4), dtype=torch.int64, device="cuda"))
    best, med = timeit(lambda: h.freivalds_witness_dev(a, b, cs, gamma, bufs["powers"], bufs["pcv"], bufs["pbv"],
                                                       bufs["pabv"], bufs["diff"], bufs["z"], bufs["inv"]), stream)
    fre_bytes = 2 * 3 * N * N * 32
    out["freivalds_N1024"] = dict(ms_best=best, ms_med=med, gbps=fre_bytes / (best * 1e-3) / 1e9,
                                  diff_zero=bool((bufs["diff"] == 0).all()))
    print("freivalds N=1024:", out["freivalds_N1024"], flush=True)
    q = torch.empty_like(a)
    wit = torch.empty((N * N, W, 4), dtype=torch.int64, device="cuda")
    best, med = timeit(lambda: h.rescale_witness_dev(cs, N * N, P, lb, q, wit), stream)
    rs_bytes = N * N * 32 * (1 + W)
    out["rescale_N1024"] = dict(ms_best=best, ms_med=med, gbps=rs_bytes / (best * 1e-3) / 1e9, W=W)
    print("rescale N=1024:", out["rescale_N1024"], flush=True)
    wit_ref, q_ref = torch.empty_like(wit), torch.empty_like(q)
    h.tune("rescale_generic", True)
    best, med = timeit(lambda: h.rescale_witness_dev(cs, N * N, P, lb, q_ref, wit_ref), stream, reps=2, warm=1)
    h.tune("rescale_generic", False)
    h.sync()
    out["rescale_N1024_generic"] = dict(ms_best=best, same_as_staged=bool((wit_ref == wit).all() and (q_ref == q).all()))
    print("rescale generic:", out["rescale_N1024_generic"], flush=True)
    del wit_ref, q_ref
    tot = torch.empty((N, 4), dtype=torch.int64, device="cuda")
    best, med = timeit(lambda: h.mat_vec_prefix_dev(cs, bufs["powers"], bufs["pcv"], tot), stream)
    out["mat_vec_prefix_N1024"] = dict(ms_best=best, ms_med=med, gbps=N * N * 64 / (best * 1e-3) / 1e9,
                                       gmuladd_s=N * N / (best * 1e-3) / 1e9)
    print("mat_vec_prefix N=1024:", out["mat_vec_prefix_N1024"], flush=True)
    best, med = timeit(lambda: h.gamma_powers_dev(gamma, N, bufs["powers"]), stream)
    out["gamma_powers_1024"] = dict(ms_best=best, ms_med=med)
    print("gamma_powers 1024:", out["gamma_powers_1024"], flush=True)
    # zkvec config 1
    B_, L_ = 4096, 1024
    x, s = rand_fr(gen, B_, L_), rand_fr(gen, B_, L_)
    o = torch.empty_like(x)
    best, med = timeit(lambda: h.zkvec_inner_prefix_dev(x, s, o), stream)
    out["zkvec_inner_4096x1024"] = dict(ms_best=best, ms_med=med, gbps=B_ * L_ * 96 / (best * 1e-3) / 1e9)
    print("zkvec inner 4096x1024:", out["zkvec_inner_4096x1024"], flush=True)
    best, med = timeit(lambda: h.zkvec_sub_dev(s, x, o), stream)
    out["zkvec_sub_4096x1024"] = dict(ms_best=best, ms_med=med, gbps=B_ * L_ * 96 / (best * 1e-3) / 1e9)
    print("zkvec sub 4096x1024:", out["zkvec_sub_4096x1024"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "triage.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    h.close()


if __name__ == "__main__":
    main()
