"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (cp.async.bulk), LDGSTS (cp.async),
plus HMMA / IMMA (legacy mma.sync -- expected 0).  Runs here, no GPU:  python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "halo2-svd041_b200", "libh2svd_b200.so")
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "HMMA", "IMMA", "IMAD.WIDE"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("h2svd::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["instructions"] += 1
            for w in WATCH:
                if op.startswith(w):
                    cur[w] += 1
                    total[w] += 1
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    print("# r02 SASS summary of `halo2-svd041_b200/libh2svd_b200.so` (`cuobjdump -sass`, sm_100a)\n")
    print("ELF images: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))) + "\n")
    print("| kernel | SASS instructions | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    for name, c in counts.items():
        print(f"| `{name[:78]}` | {c['instructions']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")
    print("| **total** | | " + " | ".join(str(total[w]) for w in WATCH) + " |")
    print("\n`UTCIMMA` = `tcgen05.mma.kind::i8` (the tensor-core mat-mul engines and the tensor-pipe micro-benchmark); `LDTM`/`STTM` = "
          "`tcgen05.ld`/`st` (accumulator read-out and re-zeroing); `UTMALDG` = TMA tensor loads of the byte planes; `UTMASTG` = TMA "
          "tensor stores (the optional witness-stream path of `rescale_tma_kernel`); `UBLKCP` = `cp.async.bulk` (operand slabs of the "
          "IMAD engines, the witness stream of the staged kernels); `LDGSTS` = `cp.async` (mat-vec tile staging).  No `HMMA`/`IMMA`: "
          "nothing runs on the legacy `mma.sync` path.")


if __name__ == "__main__":
    main()
