"""Breakdown of gpurun_out/f11_stamps.npy (profiles/tools/f11_stamps.py): where an epilogue group's period goes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
s = np.load(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "f11_stamps.npy"))
for ps, pname in ((0, "ROW"), (1, "GENE")):
    for cta in (0, 1):
        a = s[ps, cta].astype(np.int64)
        t0 = a[a > 0].min()
        a = np.where(a > 0, a - t0, -1)
        n = int((a[9] >= 0).sum())
        J = range(16, min(n, 84))
        def d(x, y):
            v = [a[y][j] - a[x][j] for j in J if a[x][j] >= 0 and a[y][j] >= 0]
            return np.mean(v)
        per = np.mean([a[9][j + 4] - a[9][j] for j in J if j + 4 < n])
        gap = np.mean([a[9][j + 4] - a[8][j] for j in J if j + 4 < n])
        print(f"{pname} cta{'0' if cta == 0 else '74'}: period/half-unit {per:.0f} cycles = wait x {d(9, 4):.0f} + x read, wait acc1 {d(4, 5):.0f} "
              f"+ tmem ld {d(5, 6):.0f} + first half to a2_empty {d(6, 7):.0f} + second half, st, arrive {d(7, 8):.0f} + bookkeeping {gap:.0f}"
              f" | x refill issued {np.mean([a[0][j + 4] - a[4][j] for j in J if j + 4 < n and a[0][j + 4] >= 0]):.0f} after x ready,"
              f" T load -> MMA1 {np.mean([a[2][i] - a[1][i] for i in range(8, 40)]):.0f}")
