"""Turns ncu outputs into the text summaries committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.txt> "<command line>"
  python tools/ncu_summary.py full <report.ncu-rep> <out.txt> "<command line>"
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path, out, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"])
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else (v * 1e6 if unit in ("s", "second") else v))
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(out, "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none (one timed step, cudaProfilerStart/Stop window)\n")
        f.write("command: %s\nper-launch times are cold-cache and serialised: compare SHARES, not absolutes\n" % cmd)
        f.write("total %.1f us over %d launches\n\n" % (tot, sum(a[0] for a in agg.values())))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-78s launches=%6d total_us=%11.1f share=%5.1f%% avg_us=%9.1f\n" % (k[:78], n, t, 100 * t / tot, t / n))
    print(open(out).read())


WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
]


def full(path, out, cmd):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("ncu --set full --clock-control none --import-source on\ncommand: %s\n\n" % cmd)
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].replace("void ", "")[:60]
            grid = r[idx["Grid Size"]]
            parts = []
            for key, lab in WANT:
                if key in idx:
                    parts.append("%s=%s%s" % (lab, r[idx[key]], units[idx[key]] if lab in ("dur", "dram_rd", "dram_wr") else ""))
            f.write("%-60s grid=%-16s %s\n" % (name, grid, " ".join(parts)))
    print(open(out).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
