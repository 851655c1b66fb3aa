"""Print key metrics + hottest SASS lines of an ncu report (first kernel)."""
import collections
import csv
import subprocess
import sys


def main(rep, top=18):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    print("kernel:", vals[hdr.index("Kernel Name")][:90])
    for i, h in enumerate(hdr):
        if h in want or (h.startswith("smsp__average_warps_issue_stalled") and float(vals[i] or 0) > 0.2):
            print(f"  {h} [{units[i]}] = {vals[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    isrc, iex, ismp, ithr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    tot = sum(int(r[iex]) for r in data)
    tot_s = sum(int(r[ismp]) for r in data)
    print(f"total warp-inst {tot}, samples {tot_s}, SASS lines {len(data)}")
    ops, smp = collections.Counter(), collections.Counter()
    for r in data:
        t = r[isrc].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[iex]); smp[op] += int(r[ismp])
    print("  by opcode:", ", ".join(f"{op} ex={c / tot * 100:.1f}%/smp={smp[op] / tot_s * 100:.1f}%" for op, c in ops.most_common(12)))
    lst = sorted([(int(r[ismp]), n, int(r[iex]), r[ithr], r[isrc].strip()[:72]) for n, r in enumerate(data)], reverse=True)[:top]
    for s, n, ex, thr, sr in lst:
        print(f"  {n:4d} smp={s / tot_s * 100:5.2f}% ex={ex / tot * 100:5.2f}% thr={thr:>4} {sr}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 18)
