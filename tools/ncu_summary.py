"""Summarise an `ncu --page raw --csv` export (optionally .gz): one line per launch with duration, DRAM bytes,
achieved DRAM GB/s, L2/L1/shared throughput, issue-slot use, registers, occupancy.
usage: python tools/ncu_summary.py profiles/r01_block_raw.csv.gz [--md]"""
import csv
import gzip
import io
import sys


def load(path):
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    hdr, units, data = load(path)
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, default=float("nan")):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return default
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return default

    def unit(name):
        i = col.get(name)
        return units[i] if i is not None else ""

    def to_bytes(v, u):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def to_us(v, u):
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)

    out = []
    for r in data:
        name = r[col["Kernel Name"]]
        dur = to_us(get(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
        rd = to_bytes(get(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"))
        wr = to_bytes(get(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
        out.append(dict(
            name=name, grid=r[col["Grid Size"]], block=r[col["Block Size"]], us=dur, rd=rd, wr=wr,
            gbs=(rd + wr) / dur / 1e3 if dur else 0,
            dram_pct=get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            l2_pct=get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            l1_pct=get(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            sm_pct=get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            issue=get(r, "sm__inst_issued.avg.pct_of_peak_sustained_active", get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
            regs=get(r, "launch__registers_per_thread"),
            occ=get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            smem=get(r, "launch__shared_mem_per_block_dynamic"),
            tensor=get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                       get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")),
        ))
    if md:
        print("| # | kernel | grid x block | µs | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | L2 % | L1 % | SM % | tensor % | regs | occ % |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for i, o in enumerate(out):
        nm = o["name"].split("(")[0][-44:]
        if md:
            print(f"| {i} | `{nm}` | {o['grid']} x {o['block']} | {o['us']:.1f} | {o['rd'] / 1e6:.1f} | {o['wr'] / 1e6:.1f} | "
                  f"{o['gbs']:.0f} | {o['dram_pct']:.0f} | {o['l2_pct']:.0f} | {o['l1_pct']:.0f} | {o['sm_pct']:.0f} | {o['tensor']:.0f} | {o['regs']:.0f} | {o['occ']:.0f} |")
        else:
            print(f"{i:3d} {nm:44s} {o['grid']:>14s}x{o['block']:<14s} {o['us']:8.1f}us rd {o['rd'] / 1e6:7.1f} wr {o['wr'] / 1e6:7.1f} MB "
                  f"{o['gbs']:6.0f} GB/s dram {o['dram_pct']:4.0f}% l2 {o['l2_pct']:4.0f}% l1 {o['l1_pct']:4.0f}% sm {o['sm_pct']:4.0f}% tc {o['tensor']:4.0f}% regs {o['regs']:4.0f} occ {o['occ']:4.0f}%")


if __name__ == "__main__":
    main()
