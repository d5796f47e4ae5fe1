"""Summarise an `ncu --set full` report into a per-kernel table for profiles/.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv
    python tools/ncu_summary.py x_raw.csv [launches.csv] > profiles/rNN_ncu_summary.md

Columns: duration, DRAM read / write bytes (dram__bytes_{read,write}.sum), DRAM GB/s over the launch, DRAM % of peak,
tensor-pipe active % (sm__pipe_tensor_cycles_active_realtime), TMEM-pipe instruction %, achieved occupancy, registers.
With a second argument (the `--metrics gpu__time_duration.sum` launch list of a timed bench region) a per-kernel share
table of the step is appended.
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name[:44]


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    H, U = rows[0], rows[1]

    def col(name):
        return H.index(name) if name in H else None

    cols = {
        "dur": col("gpu__time_duration.sum"),
        "rd": col("dram__bytes_read.sum"),
        "wr": col("dram__bytes_write.sum"),
        "dram_pct": col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "tensor": col("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        "tmem": col("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"),
        "occ": col("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "regs": col("launch__registers_per_thread"),
        "grid": col("Grid Size"),
        "block": col("Block Size"),
        "l2": col("lts__t_bytes.sum"),
    }
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    print("| kernel | grid x block | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | tensor pipe % | tmem inst % | occupancy % | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        name = short(r[col("Kernel Name")])
        t = fnum(r[cols["dur"]]) * tscale.get(U[cols["dur"]], 1.0)
        rd = fnum(r[cols["rd"]]) * scale.get(U[cols["rd"]], 1.0)
        wr = fnum(r[cols["wr"]]) * scale.get(U[cols["wr"]], 1.0)

        def g(k, fmt="%.1f"):
            i = cols[k]
            if i is None or r[i] == "":
                return "-"
            v = fnum(r[i])
            return "-" if v != v else fmt % v
        print(f"| {name} | {r[cols['grid']]} x {r[cols['block']]} | {t * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
              f"{(rd + wr) / t / 1e9:.0f} | {g('dram_pct')} | {g('tensor')} | {g('tmem')} | {g('occ')} | {g('regs', '%.0f')} |")
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        H2 = rows[hdr]
        d = collections.OrderedDict()
        for r in rows[hdr + 1:]:
            if len(r) < len(H2) or not r[0].isdigit():
                continue
            name = short(r[H2.index("Kernel Name")])
            v = fnum(r[-1])
            d.setdefault(name, [0, 0.0])
            d[name][0] += 1
            d[name][1] += v
        tot = sum(v[1] for v in d.values())
        print()
        print(f"Launch list of the timed region (gpu__time_duration.sum, serialised, cold cache): {sum(v[0] for v in d.values())} launches, "
              f"{tot / 1e6:.3f} ms of kernel time")
        print()
        print("| kernel | launches | total us | share % |")
        print("|---|---|---|---|")
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
            print(f"| {k} | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f} |")


if __name__ == "__main__":
    main()
