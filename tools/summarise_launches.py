"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def short(n):
    m = re.search(r"gemm_kernel<(\d+), ?(\d+), ?(\d+), ?(\d+)>", n)
    if m:
        names = ["F32", "BIAS_BF16", "RELU", "RELU_POOL22", "RELU_POOL12", "STATS", "LSTM", "LOGITS", "XPROJ", "CONV_STORE(dgrad)", "RELU_POOL22_T", "RELU_POOL12_T"]
        epi = names[int(m.group(3))] if int(m.group(3)) < len(names) else m.group(3)
        return f'gemm<{m.group(1)},{"CONV3" if m.group(2) == "1" else "PLAIN"},{epi}>'
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("(anonymous namespace)::", "")[:60]


def main(path, skip=0):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))[skip:]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total device time (cold-cache, serialised under ncu)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:45s} n={c:4d} total={t / 1e3:10.1f} us  avg={t / c / 1e3:9.2f} us  share={t / tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
