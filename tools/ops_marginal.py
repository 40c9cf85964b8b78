"""Split every op's time into fixed + marginal from two profile_ops.py tables at batch B and 2B:
    python tools/ops_marginal.py ops_B.txt ops_2B.txt
marginal = t(2B) - t(B) (time the extra B images cost), fixed = 2 t(B) - t(2B)."""
import sys


def load(path):
    rows = {}
    for line in open(path):
        f = line.split()
        if len(f) >= 11 and f[0].isdigit():
            rows[f[2]] = (float(f[3]), float(f[5]), float(f[6]), f[1])
    return rows


a, b = load(sys.argv[1]), load(sys.argv[2])
tot_f = tot_m = 0.0
print(f"{'op':28s} {'kind':6s} {'t(B) ms':>8s} {'t(2B) ms':>8s} {'fixed':>7s} {'marg':>7s}  marg TFLOP/s  marg GB/s")
for name, (t1, gf, mb, kind) in a.items():
    if name not in b:
        continue
    t2 = b[name][0]
    m, fx = t2 - t1, 2 * t1 - t2
    tot_f += fx
    tot_m += m
    print(f"{name:28s} {kind:6s} {t1:8.4f} {t2:8.4f} {fx:7.4f} {m:7.4f}  {gf / max(m, 1e-6) / 1e3:10.1f}  {mb / max(m, 1e-6) / 1e3:9.0f}")
print(f"# total fixed {tot_f:.3f} ms, total marginal {tot_m:.3f} ms per batch B")
