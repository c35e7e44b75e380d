"""Decode gpurun_out/tc_trace.bin (PMF_TC_TRACE): per-tile event times of CTA 0, in cycles relative to the first stamp."""
import sys
import numpy as np
EV = ["tmaA_issue", "M1_issue", "G_seen", "M2_issue", "M3_issue", "epi_top", "Z_seen", "A_seen", "epi_done", "G_arrived",
      "dx_begin", "DXFULL_seen", "dx_end", "M1_top", "M1_issued", "M3_issued"]
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 32)
t0 = a[a > 0].min()
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20, 44)
print("tile " + " ".join(f"{e:>11}" for e in EV))
for g in range(lo, hi):
    print(f"{g:4d} " + " ".join(f"{(a[g, i] - t0) if a[g, i] > 0 else -1:11d}" for i in range(len(EV))))
d = a[lo:hi]
print("per-tile period (G_arrived):", np.diff(d[:, 9]).mean())
for name, x, y in [("A wait (A_seen - Z_seen)", 7, 6), ("Z wait (Z_seen - epi_top)", 6, 5), ("epi body (epi_done - A_seen)", 8, 7),
                   ("drain+arrive (G_arrived - epi_done)", 9, 8), ("MMA sees G (G_seen - G_arrived)", 2, 9),
                   ("M2 issue after G_seen", 3, 2), ("M3 issue after M2", 4, 3), ("DXFULL wait", 11, 10), ("dx_out body", 12, 11),
                   ("TMA A issue(g+3) - M2_issue(g)", None, None)]:
    if x is None:
        v = a[lo + 3:hi + 3, 0] - a[lo:hi, 3]
    else:
        v = d[:, x] - d[:, y]
    print(f"{name:40s} mean {v.mean():9.1f}  min {v.min():7d}  max {v.max():7d}")
for name, x, y in [("M1: XK wait (M1_issue - M1_top)", 1, 13), ("M1: 24 MMAs issue (M1_issued - M1_issue)", 14, 1),
                   ("M3: 8 MMAs issue (M3_issued - M3_issue)", 15, 4), ("M2 16 MMAs+XM wait (M3_issue - M2_issue)", 4, 3)]:
    v = d[:, x] - d[:, y]
    print(f"{name:40s} mean {v.mean():9.1f}  min {v.min():7d}  max {v.max():7d}")
v = a[lo + 3:hi + 3, 7] - a[lo + 3:hi + 3, 0]
print(f"{'A_seen(g) - tmaA_issue(g)':40s} mean {v.mean():9.1f}  min {v.min():7d}  max {v.max():7d}")

print("item boundaries (events keyed by the first tile of the next item; A = group 0, B = group 1):")
names = ["flushed", "DYFULL_seen", "dY_stored", "Y_computed", "Y_fenced", "Y_arrived"]
for g in range(1, a.shape[0]):
    if a[g, 17] > 0 or a[g, 23] > 0:
        print(f" tile {g}: " + " | ".join(f"{grp}:" + ",".join(f"{n}={a[g, 16 + 6 * k + i] - t0 if a[g, 16 + 6 * k + i] > 0 else -1}" for i, n in enumerate(names)) for k, grp in enumerate("AB")),
              f"| MMA Y_READY->M1_issue(g)={a[g, 1] - t0}")

# finer stamps of the dX flush (events 28..31): TMEM read done, staging buffer free, staged + arrived, issuer woke up
if a.shape[1] >= 32 and (a[lo:hi, 28] > 0).all():
    print("dX flush of tile g (cycles after DXFULL_seen):")
    for name, ev in [("TMEM read + DX_EMPTY arrive", 28), ("staging buffer free (DXS_DONE x2)", 29), ("staged, DXS_FULL arrive", 30),
                     ("issuer saw DXS_FULL", 31), ("TMA read of staging done", 12)]:
        v = d[:, ev] - d[:, 11]
        print(f"  {name:38s} mean {v.mean():9.1f}  min {v.min():7d}  max {v.max():7d}")
    v = d[2:, 5] - d[:-2, 30]
    print(f"  {'epi_top(g+2) - staged(g-2..)':38s} (next tile of the group starts) see rows")
