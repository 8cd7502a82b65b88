"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the
last complete forward (cold-cache, serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, seq = None, []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        val = float(d["Metric Value"].replace(",", ""))
        val *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
        seq.append((re.sub(r"\(.*", "", d["Kernel Name"])[:60], val))
    starts = [i for i, (n, _) in enumerate(seq) if "k_csr_hist" in n]
    print(f"launches captured: {len(seq)}; forwards seen: {len(starts)}")
    if "--all" in sys.argv:                      # everything captured (e.g. whole training steps)
        fwd = seq
    elif len(starts) < 2:
        fwd = seq[starts[-1]:] if starts else seq
    else:
        fwd = seq[starts[-2]:starts[-1]]
    agg = collections.OrderedDict()
    for n, v in fwd:
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"{'kernel':62s} {'n':>3s} {'ms':>9s} {'share':>7s}")
    for n, (v, c) in agg.items():
        print(f"{n:62s} {c:3d} {v:9.3f} {100 * v / tot:6.1f}%")
    print(f"{'sum of kernel durations':62s} {len(fwd):3d} {tot:9.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
