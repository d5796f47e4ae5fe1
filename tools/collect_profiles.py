"""Copy this round's evidence from gpurun_out/ (scratch) into profiles/ (tracked) and print the scaling table of DESIGN.md section 6.

    python tools/collect_profiles.py            # after tools/run_scaling.sh N for N in 2 4 8 and a 1-GPU `bench.py` run
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    try:
        lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except Exception:
        return None


def main():
    table = {}
    for cfg in ("config3", "config4", "config5"):
        for n in (1, 2, 4, 8):
            name = f"r02_bench_{cfg}_n{n}.json"
            d = last_json(os.path.join(SRC, name))
            if d is None:
                continue
            shutil.copy(os.path.join(SRC, name), os.path.join(DST, name))
            table[(cfg, n)] = d
    for n in (2, 4, 8):
        for p in (0, 1):
            f = f"r02_dp_equivalence_n{n}_p2p{p}.log"
            if os.path.exists(os.path.join(SRC, f)):
                keep = [l for l in open(os.path.join(SRC, f)).read().splitlines()
                        if l.startswith(("world=", "sharded", "peer all-reduce", "DP EQUIVALENCE", "rc="))]
                open(os.path.join(DST, f), "w").write("\n".join(keep) + "\n")
    for cfg in ("config3", "config4", "config5"):
        base = table.get((cfg, 1))
        row = []
        for n in (1, 2, 4, 8):
            d = table.get((cfg, n))
            if d is None:
                row.append("-")
                continue
            s = f"{d['value'] / 1e6:.1f} ({d['ms_per_step']:.2f} ms"
            if base is not None and n > 1:
                s += f", {100.0 * d['value'] / (n * base['value']):.1f} %"
            if "volume_query" in d:
                s += f"; volume {d['volume_query']['ms']:.1f} ms"
            row.append(s + ")")
        print(f"| {cfg} | " + " | ".join(row) + " |")


if __name__ == "__main__":
    sys.exit(main())
