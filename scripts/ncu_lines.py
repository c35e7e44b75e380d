"""Per-source-line stall samples of an .ncu-rep captured with --import-source on (-lineinfo build)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
i_s = hdr.index("# Samples"); i_e = hdr.index("Instructions Executed")
lines = []
cur = None
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= i_s: continue
    if r[0] != "" and r[0].isdigit():
        lines.append([r[0], r[1].strip(), int(r[i_s]) if r[i_s].isdigit() else 0, int(r[i_e]) if r[i_e].isdigit() else 0])
tot = sum(l[2] for l in lines)
print(f"# {tot} samples")
for l in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{l[2]:6d} {100.0*l[2]/tot:5.1f}%  exec={l[3]:9d}  L{l[0]:>4}: {l[1][:110]}")
