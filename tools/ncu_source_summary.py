"""Per-source-line summary of an .ncu-rep captured with --import-source on (read here, no GPU needed): share of the
warp-stall samples and of the executed warp instructions per line and per code region.

    python tools/ncu_source_summary.py gpurun_out/s6_gnm.ncu-rep profiles/frame_gn_mixed_source_r01.txt [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, out, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    cur = None
    agg = collections.defaultdict(lambda: [0, 0, ""])  # (file, line) -> [samples, warp instructions, source text]
    for r in csv.reader(io.StringIO(txt)):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0].isdigit() and len(r) > 7:
            try:
                s, n = int(r[4]), int(r[7])
            except ValueError:
                continue
            k = (cur, int(r[0]))
            agg[k][0] += s
            agg[k][1] += n
            agg[k][2] = r[1]
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    lines = [f"source-level summary of {rep}: {tot_s} stall samples, {tot_i} warp instructions",
             "per file (share of samples / of instructions):"]
    byfile = collections.defaultdict(lambda: [0, 0])
    for (f, _), v in agg.items():
        byfile[f][0] += v[0]
        byfile[f][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"  {f:28s} {100 * v[0] / tot_s:6.2f} %  {100 * v[1] / tot_i:6.2f} %")
    lines.append(f"top {top} lines by stall samples (file:line, % samples, % instructions, source):")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        lines.append(f"  {f}:{l:<5d} {100 * v[0] / tot_s:6.2f} %  {100 * v[1] / tot_i:6.2f} %  {v[2].strip()[:100]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
