"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file ...`, optionally .gz):
per-kernel total time, share of all launches, count.  usage: python tools/launch_list_summary.py <csv[.gz]> [top]"""
import collections
import csv
import gzip
import io
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    lines = raw.splitlines()
    i0 = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    tot, cnt = collections.Counter(), collections.Counter()
    for r in csv.DictReader(io.StringIO("\n".join(lines[i0:]))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        k = r["Kernel Name"].split("(")[0][:70]
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    print(f"{sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time (cold-cache, serialised by ncu)")
    for k, v in tot.most_common(top):
        print(f"{v / 1e3:9.3f} ms {100 * v / T:5.1f}%  x{cnt[k]:5d}  avg {v / cnt[k]:8.1f} us  {k}")
    ours = sum(v for k, v in tot.items() if k.startswith(("void tc::", "tc::", "void blk::", "blk::", "void ln2d::", "ln2d::", "void slak", "slak", "void g2::", "g2::", "void mlp::", "mlp::", "void dense::", "dense::")))
    print(f"kernels of this repo: {100 * ours / T:.1f}% of the kernel time")


if __name__ == "__main__":
    main()
