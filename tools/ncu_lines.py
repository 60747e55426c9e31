"""Top source lines by warp-stall samples from an `ncu --page source --csv --print-source cuda,sass` export.
usage: python tools/ncu_lines.py profiles/<file>.csv.gz [top] [--sass]"""
import csv
import gzip
import io
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
    want_sass = "--sass" in sys.argv
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    lines, sass = [], []
    cur_file = ""
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = r
            i_all = hdr.index("Warp Stall Sampling (All Samples)")
            i_exec = hdr.index("Instructions Executed")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        try:
            smp = float(r[i_all])
        except ValueError:
            continue
        st = sorted(((float(r[i]) if r[i] not in ("", "-") else 0.0, h) for i, h in stall_cols), reverse=True)[:3]
        rec = (smp, cur_file, r[0], r[1].strip()[:90] if r[0] else r[3].strip()[:90], r[i_exec], st)
        (lines if r[0] else sass).append(rec)
    tot = sum(x[0] for x in lines)
    print(f"total samples {tot:.0f}")
    for smp, f, ln, src, ex, st in sorted(sass if want_sass else lines, reverse=True)[:top]:
        sts = " ".join(f"{h[6:]}={v:.0f}" for v, h in st if v > 0)
        print(f"{100 * smp / tot:5.1f}% {f}:{ln:>4s} exec={ex:>8s} | {src} | {sts}")


if __name__ == "__main__":
    main()
