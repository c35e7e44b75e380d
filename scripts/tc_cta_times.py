"""Per-CTA durations (globaltimer ns) from a PMF_TC_TRACE dump: load balance of the cost-weighted range cut."""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], dtype=np.int64)
t = a[96 * 32:96 * 32 + 2 * 160].reshape(-1, 2)[:148]
d = (t[:, 1] - t[:, 0]) / 1e3
print("per-CTA us: min %.1f max %.1f mean %.1f; start spread %.1f us" % (d.min(), d.max(), d.mean(), (t[:, 0].max() - t[:, 0].min()) / 1e3))
print("by CTA (us):", " ".join(f"{x:.0f}" for x in d))
