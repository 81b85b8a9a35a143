"""Top stall sites of one kernel from an ncu report's source page (SASS view), grouped by reason.

    python tools/ncu_stalls.py <report.ncu-rep> <kernel regex | launch index> [top]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    sel = ["--launch-skip", pat, "--launch-count", "1"] if pat.isdigit() else ["--kernel-name", "regex:" + pat]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
    reasons = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print("kernel %s (%s): %d samples" % (pat, lines[start - 1][:110] if start else "", tot))
    agg = {k: sum(int(r[k] or 0) for r in rows) for k in reasons}
    print("by reason: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
    rows_s = sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]
    for i, r in sorted(rows_s):
        n = int(r["# Samples"] or 0)
        why = sorted(((int(r[k] or 0), k[6:]) for k in reasons), reverse=True)[:2]
        print("%5.1f%%  %-8s %-70s %s" % (100.0 * n / max(tot, 1), r["Address"][-5:], r["Source"][:70],
                                         " ".join("%s=%d" % (k, v) for v, k in why if v)))


if __name__ == "__main__":
    main()
