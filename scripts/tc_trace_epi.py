"""Decode a PMF_TC_TRACE dump taken with PMF_TC_FLAGS=64: phases of ONE epilogue warp (group 0, quarter 0, half 0) per tile."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64)[:96 * 32].reshape(-1, 32)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20, 60)
rows = [g for g in range(lo, hi) if a[g, 5] > 0 and a[g, 9] > 0]
names = [("Z wait", 6, 5), ("A wait", 7, 6), ("h0 loads (LDTM+LDS+wait::ld)", 16, 7), ("h0 math", 17, 16), ("h0 stores issued", 18, 17),
         ("h1 loads", 19, 18), ("h1 math", 20, 19), ("h1 stores issued", 21, 20), ("wait::st", 8, 21), ("fence+arrive", 9, 8),
         ("whole tile (arrive - top)", 9, 5)]
for n, x, y in names:
    v = np.array([a[g, x] - a[g, y] for g in rows])
    print(f"{n:34s} mean {v.mean():8.1f} min {v.min():6d} max {v.max():6d}")
tops = np.array([a[g, 5] for g in rows])
print("period of this group's tiles (top to top):", np.diff(tops).mean(), " idle between arrive and next top:",
      np.mean([a[rows[i + 1], 5] - a[rows[i], 9] for i in range(len(rows) - 1)]))
