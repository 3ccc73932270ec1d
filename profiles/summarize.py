"""Summarise ncu exports kept under profiles/: launch list (gpu__time_duration) and `--set full` raw page.

    python profiles/summarize.py launches profiles/<round>_launches.csv
    python profiles/summarize.py raw      <ncu --page raw --csv export>
"""
import collections
import csv
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except Exception:
            continue
        ms = v / 1e6 if r[ui] in ("nsecond", "ns") else v / 1e3 if r[ui] in ("usecond", "us") else v
        agg[r[ki].split("(")[0]][0] += 1
        agg[r[ki].split("(")[0]][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| %s | %d | %.3f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))


WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    idx = {w: H.index(w) for w in WANT if w in H}
    ki = H.index("Kernel Name")
    print("| metric | unit | " + " | ".join(r[ki].split("(")[0] for r in rows[2:]) + " |")
    print("|---|---|" + "---|" * (len(rows) - 2))
    for w in WANT:
        if w in idx:
            print("| %s | %s | " % (w, U[idx[w]]) + " | ".join(r[idx[w]][:14] for r in rows[2:]) + " |")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else raw)(sys.argv[2])
