"""Summarise an .ncu-rep (one or more kernel launches) into a small text file for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_r01.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of {rep} ({len(rows) - 2} launch(es)); full metric set, --clock-control none")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n## {d.get('Kernel Name', '?')}  grid={d.get('Grid Size', '?')} block={d.get('Block Size', '?')}")
        for k in KEYS:
            for h, u in zip(hdr, units):
                if h == k or h.endswith("." + k) or h.endswith(k):
                    print(f"{k:90s} {d[h]:>16s} {u}")
                    break
    if len(sys.argv) > 2 and sys.argv[2] == "--source":
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr) and (r[ix["# Samples"]] or "0").isdigit()]   # several launches repeat the header
        uniq, seen = [], set()
        for r in data:          # the page lists an instruction once per launch of the report: keep the first
            if tuple(r) not in seen:
                seen.add(tuple(r))
                uniq.append(r)
        data = uniq
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        print("\n## top sampled SASS instructions (share of warp-stall samples, dominant stall reasons)")
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]:
            n = int(r[ix["# Samples"]] or 0)
            st = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:3]
            print(f"{100 * n / tot:5.1f}%  {r[ix['Source']][:64]:64s} " + " ".join(f"{s}={v}" for v, s in st if v))


if __name__ == "__main__":
    main()
