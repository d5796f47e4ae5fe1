"""Fill the numeric placeholders of DESIGN.md (scaling table, config-3 table) from the bench lines kept under profiles/."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(cfg, n):
    try:
        ls = [l for l in open(os.path.join(ROOT, "profiles", f"r02_bench_{cfg}_n{n}.json")).read().splitlines() if l.startswith("{")]
        return json.loads(ls[-1])
    except Exception:
        return None


def main():
    p = os.path.join(ROOT, "DESIGN.md")
    s = open(p).read()
    for tag, cfg in (("SCALE3", "config3"), ("SCALE4", "config4"), ("SCALE5", "config5")):
        base = line(cfg, 1)
        for n in (1, 2, 4, 8):
            d = line(cfg, n)
            if d is None:
                txt = "n/a"
            else:
                txt = f"{d['value'] / 1e6:.1f} ({d['ms_per_step']:.2f} ms/step"
                if n > 1 and base is not None:
                    txt += f", {100.0 * d['value'] / (n * base['value']):.1f} % of linear"
                if "volume_query" in d:
                    txt += f"; {d['volume_query']['ms']:.1f} ms"
                txt += ")"
            s = s.replace(f"{tag}_{n} ", txt + " ")
    d = line("config3", 1)
    if d is not None:
        r = d["roofline"]
        rep = {"BENCH3_VALUE": f"{d['value'] / 1e6:.2f} M", "BENCH3_EARLY_MS": f"{d['early_window']['ms_per_step']:.2f}",
               "BENCH3_EARLY": f"{d['early_window']['value'] / 1e6:.1f} M",
               "BENCH3_MS": f"{d['ms_per_step']:.3f} ({d['spread']['min_ms_per_step']:.3f}–{d['spread']['max_ms_per_step']:.3f})",
               "BENCH3_E2E": f"{d['e2e']['value'] / 1e6:.2f} M", "BENCH3_TFLOPS": f"{r['achieved']:.0f}",
               "BENCH3_FRACS": f"{r['frac_of_sustained_peak']:.2f}", "BENCH3_FRAC": f"{r['frac']:.2f}"}
        for k, v in rep.items():
            s = s.replace(k, v)
    open(p, "w").write(s)


if __name__ == "__main__":
    main()
