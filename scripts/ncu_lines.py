"""Per-source-line share of stall samples / executed instructions of one kernel from an .ncu-rep (cuda,sass view)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; agg = []; ts = ti = 0
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        try: s = float(r[6]); n = float(r[7])
        except ValueError: continue
        agg.append((s, n, cur, int(r[0]), r[1].strip()[:120])); ts += s; ti += n
agg.sort(reverse=True)
for s, n, f, l, src in agg[:top]:
    print(f"{100 * s / ts:5.2f}% smp {100 * n / ti:5.2f}% ins {f}:{l} {src}")
