import sys, numpy as np
rows = [np.array(l.split(), dtype=np.uint64).astype(np.int64) for l in open(sys.argv[1])]
t0 = min(r[0] for r in rows if r[0] > 0)
for w, r in enumerate(rows):
    r = r[r > 0] - t0
    ev = r[: (len(r) // 4) * 4].reshape(-1, 4)      # per chunk: [start, after producer, after wait, after k-loop]
    prod = ev[:, 1] - ev[:, 0]; wait = ev[:, 2] - ev[:, 1]; comp = ev[:, 3] - ev[:, 2]
    gap = ev[1:, 0] - ev[:-1, 3]
    print(f"warp {w}: chunks {len(ev)} first start {ev[0,0]:6d}  producer {np.median(prod):5.0f} (max {prod.max()})  wait {np.median(wait):5.0f} (p90 {np.percentile(wait,90):.0f}, max {wait.max()})  compute {np.median(comp):6.0f}  gap(end->next start) median {np.median(gap):5.0f} p90 {np.percentile(gap,90):.0f} max {gap.max()}")
    if w in (0, 4):
        print("   chunk starts:", ev[:40, 0].tolist())
        print("   waits      :", wait[:40].tolist())
        print("   gaps       :", gap[:40].tolist())
