"""Tile / wave cost model of the CTA-pair W4A16 kernel against the measured per-shape tables (profiles/gemm_layers*_r01.json).

Model (profiles/_notes.md): a pair processes one k-block of a 256 x tile_n tile in ~(440 + 1.6 tile_n) cycles (A box + packed
rows + MMA), tiles run in waves of 74 pairs, plus a fixed cost per launch.  The script prints, per shape, the measured time,
the modelled time with the widths the kernel can use today (<= 256 columns, double-buffered accumulators) and what a
single-wave tiling would give if tiles up to 512 columns (single-buffered TMEM) were available -- the evidence behind item 1
of DESIGN.md section 9.      python tools/wave_model.py [sd15|sdxl|sd35]"""
import json
import math
import os
import sys

R = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
PAIRS, GHZ, FIXED_US = 74, 1.9, 6.0


def cost(m, n, k, tn):
    tiles = math.ceil(m / 256) * math.ceil(n / tn)
    waves = math.ceil(tiles / PAIRS)
    return waves * (k / 64) * (440 + 1.6 * tn) / (GHZ * 1e3) + FIXED_US, waves, tiles


def best(m, n, k, max_tn):
    cands = [(cost(m, n, k, tn), tn) for tn in range(32, max_tn + 1, 16)]
    (us, waves, tiles), tn = min(cands)
    return us, waves, tiles, tn


def main():
    mdl = sys.argv[1] if len(sys.argv) > 1 else "sd15"
    name = "gemm_layers_r01.json" if mdl == "sd15" else f"gemm_layers_{mdl}_r01.json"
    rows = json.load(open(os.path.join(R, name)))["rows"]
    print(f"{'M':>6} {'N':>6} {'K':>6} {'calls':>5} | {'measured':>8} | {'model<=256':>10} waves tile | {'model<=512':>10} waves tile | saving/step")
    tot_meas = tot_new = 0.0
    for r in rows:
        m, n, k, c = r["M"], r["N"], r["K"], r.get("calls_per_step", 1)
        if m <= 128:
            continue
        meas = r["w4a16"]["ms"] * 1e3
        us1, w1, _, t1 = best(m, n, k, 256)
        us2, w2, _, t2 = best(m, n, k, 512)
        gain = max(0.0, us1 - us2) * c
        tot_meas += meas * c
        tot_new += (meas - max(0.0, us1 - us2)) * c
        print(f"{m:6d} {n:6d} {k:6d} {c:5d} | {meas:8.1f} | {us1:10.1f} {w1:5d} {t1:4d} | {us2:10.1f} {w2:5d} {t2:4d} | {gain / 1e3:6.2f} ms")
    print(f"large-M launches of the step: {tot_meas / 1e3:.2f} ms measured -> {tot_new / 1e3:.2f} ms with single-wave tiles "
          f"({100 * (1 - tot_new / tot_meas):.0f} % less)")


if __name__ == "__main__":
    main()
