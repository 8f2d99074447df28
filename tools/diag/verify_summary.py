import re, sys
bad = []
n = 0
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r'bounce (\d+) item (\d+) .* fused t (\S+) prim (\d+) prop (\d+) \| reference t (\S+) prim (\d+) prop (\d+)', line)
    if not m:
        continue
    n += 1
    g = m.groups()
    tf, tr = float(g[2]), float(g[5])
    if g[2] != g[5] or g[4] != g[7]:
        bad.append((abs(tf - tr) / max(tr, 1e-9), line.strip()))
bad.sort(reverse=True)
print(sys.argv[1], "mismatches", n, "with another t or prop", len(bad), "relative t difference > 1e-3:", sum(1 for r, _ in bad if r > 1e-3))
for r, b in bad[:4]:
    print("  %.2e" % r, b[:330])
