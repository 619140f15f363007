import sys, numpy as np
rows = [np.array(l.split(), dtype=np.uint64).astype(np.int64) for l in open(sys.argv[1])]
t0 = min(r[0] for r in rows if r[0] > 0)
ev = []
for r in rows:
    r = r[r > 0] - t0
    ev.append(r[: (len(r) // 4) * 4].reshape(-1, 4))
for pair in ((0, 4), (1, 5)):
    a, b = ev[pair[0]], ev[pair[1]]
    T0 = max(a[2, 0], b[2, 0]); T1 = min(a[-2, 3], b[-2, 3])
    grid = np.arange(T0, T1)
    def busy(e):
        m = np.zeros(grid.size, bool)
        for s, t in zip(e[:, 2], e[:, 3]):
            lo, hi = max(s, T0) - T0, min(t, T1) - T0
            if hi > lo: m[lo:hi] = True
        return m
    ma, mb = busy(a), busy(b)
    print(f"warps {pair}: window {T1-T0} clks; both computing {np.mean(ma&mb):.3f}, exactly one {np.mean(ma^mb):.3f}, none {np.mean(~ma&~mb):.3f}")
    # compute duration vs overlap: chunks
    d = a[:, 3] - a[:, 2]
    print("   warp", pair[0], "compute per chunk: median", np.median(d), "min", d.min(), "max", d.max(), " period median", np.median(np.diff(a[:, 0])))
    print("   lead of warp", pair[1], "over", pair[0], "(chunk start diff, clks):", (a[:24, 0] - b[:24, 0]).tolist())
