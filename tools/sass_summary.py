"""Per-kernel SASS evidence from the built library (no GPU needed): which kernels contain TMA bulk copies (UBLKCP /
UTMALDG), tcgen05 tensor-core instructions (UTCHMMA / UTCMMA, TMEM loads LDTM), packed fp32 FMAs (FFMA2), fp64 FMAs,
cluster barriers, local-memory traffic, and how many registers each one uses.

    python tools/sass_summary.py [gmmloc_b200/lib/libgmmloc_b200.so] [profiles/sass_r02.txt]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
MNEMONICS = ["UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCMMA", "LDTM", "STTM", "FFMA2", "FFMA", "DFMA", "HMMA", "SHFL", "LDGSTS",
             "UCGABAR", "BAR", "ATOM", "RED", "LDL", "STL", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(lib, out):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = (int(m.group(1)), int(m.group(2)))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMONICS:
                if op == k or (k in ("UBLKCP", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCMMA", "UCGABAR") and op.startswith(k)):
                    counts[cur][k] += 1
    names = demangle(list(counts))
    lines = [f"SASS summary of {Path(lib).name} (cuobjdump -sass; sm_100a).  Columns: instructions in the kernel's code "
             "(static counts, not executed counts).",
             "UBLKCP = TMA bulk copy (cp.async.bulk), UTMALDG = TMA tensor load, UTCHMMA/UTCMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), "
             "FFMA2 = packed fma.rn.f32x2, UCGABAR = cluster barrier (barrier.cluster), LDL/STL = local memory (spills, arrays)", ""]
    hdr = f"{'kernel':58s} {'regs':>5s} {'smem':>6s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in MNEMONICS)
    lines.append(hdr)
    for fn, c in counts.items():
        short = names.get(fn, fn).replace("(anonymous namespace)::", "").replace("gmmloc::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        r = regs.get(fn, (0, 0))
        lines.append(f"{short[:58]:58s} {r[0]:5d} {r[1]:6d} {c['_total']:6d} " + " ".join(f"{c[k]:7d}" for k in MNEMONICS))
    Path(out).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gmmloc_b200" / "lib" / "libgmmloc_b200.so"),
         sys.argv[2] if len(sys.argv) > 2 else str(ROOT / "profiles" / "sass_r02.txt"))
