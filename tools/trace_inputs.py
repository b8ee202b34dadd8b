"""Per k-block: when the A load was issued (leader / peer), when the dequant warps arrived, when the MMA saw `full`.
python tools/trace_inputs.py gpurun_out/trace_x.txt [k0 k1]"""
import sys
ev = []
for line in open(sys.argv[1]):
    if line.startswith('QDMTRACE begin'):
        ev = []; hdr = line.strip()
    elif line.startswith('QDMTRACE '):
        _, r, tag, t = line.split(); ev.append((int(t), int(r), int(tag) // 1000000, int(tag) % 1000000))
print(hdr)
t0 = {0: min(t for t, r, e, i in ev if r < 8), 1: min([t for t, r, e, i in ev if r >= 8] or [0])}
d = {}
ngroups = 1 + max(r % 8 for t, r, e, i in ev if 3 <= r % 8 <= 6) - 3
for t, r, e, i in ev:
    d[(r, e, i)] = t - t0[r // 8]
k0, k1 = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (10, 40)
print("  k   A-issue(L/P)      dq-arrive(L/P)     mma-full   full-A   full-dq   epi(1=acc full,2=drained)")
for k in range(k0, k1):
    ai, ap, g = d.get((0, 2, k)), d.get((8, 2, k)), 3 + k % ngroups
    dl, dp, mf = d.get((g, 3, k)), d.get((g + 8, 3, k)), d.get((1, 2, k))
    if mf is None: break
    print(f"{k:3d} {ai!s:>7}/{ap!s:<7}   {dl!s:>7}/{dp!s:<7} {mf:>9}   {mf - max(ai or 0, ap or 0):6d}  {mf - max(dl or 0, dp or 0):6d}")
