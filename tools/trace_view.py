"""Timeline of block 0's roles from a QDM_TRACE run (stderr lines 'QDMTRACE region tag clock').
regions: 0 TMA producer (1=raw slot free, 2=A stage free), 1 MMA issuer (1=accumulator free, 2=stage full),
2 epilogue warp 4 (1=accumulator full, 2=drained), 3/4 dequant group 0/1 (1=raw landed, 2=B stage free, 3=arrived)."""
import sys
ev = []
for line in open(sys.argv[1]):
    if line.startswith("QDMTRACE begin"):
        ev = []
        print(line.strip())
    elif line.startswith("QDMTRACE "):
        _, r, tag, t = line.split()
        ev.append((int(t), int(r), int(tag) // 1000000, int(tag) % 1000000))
ev.sort()
t0 = ev[0][0]
names = {0: "tma", 1: "mma", 2: "epi", 3: "dq0", 4: "dq1", 5: "dq2", 6: "dq3", 7: "raw"}
names.update({k + 8: "P:" + v for k, v in list(names.items())})   # peer CTA (block 1); clocks of the two SMs are not aligned
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10 ** 9)
for t, r, e, it in ev:
    if lo <= t - t0 <= hi:
        print(f"{t - t0:8d}  {'      ' * r}{names[r]}.{e}#{it}")
