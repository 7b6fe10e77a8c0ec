"""One line per leg of a bench.py JSON line:  python tools/bench_summary.py gpurun_out/bench_x.json"""
import json
import sys

d = json.load(open(sys.argv[1]))


def row(name, v):
    if not isinstance(v, dict):
        return
    if "error" in v:
        print(f"{name:28s} ERROR {v['error']}")
        return
    r = v.get("roofline")
    if r:
        mb = r.get("min_bytes", {})
        print(f"{name:28s} ms/step {v['ms_per_step']:.4f} fwd {r['fwd']['ms']:.4f} ({r['fwd']['frac']:.3f}) bwd {r['bwd']['ms']:.4f} "
              f"({r['bwd']['frac']:.3f}) fwd+bwd {r['fwd_bwd_frac']:.3f} [min-bytes {mb.get('fwd_bwd_frac', 0):.3f}] traffic {r.get('traffic')}")
    else:
        print(f"{name:28s} " + " ".join(f"{k}={v[k]:.4g}" if isinstance(v[k], float) else f"{k}={v[k]}" for k in v
                                         if k not in ("note", "unit", "config", "sample") and not isinstance(v[k], (dict, list)))[:300])
        for k in v:
            if isinstance(v[k], dict) and k != "roofline":
                print(f"{'':28s}   {k}: " + " ".join(f"{kk}={vv:.4g}" if isinstance(vv, float) else f"{kk}={vv}" for kk, vv in v[k].items())[:260])


row("HEADLINE " + d["config"]["fragments"], d)
print(f"{'':28s} value {d['value']:.4g} {d['unit']}  n_gpus {d['n_gpus']}  repeats {['%.4f' % t for t in d.get('repeat_ms_per_step', [])]}  peak mem {d.get('peak_memory_bytes')}")
for k, v in (d.get("also") or {}).items():
    if k == "configs":
        for kk, vv in v.items():
            row(kk, vv)
    else:
        row(k, v)
print("e2e", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d.get("e2e") or {}).items() if k != "note"})
print("cpu_baseline", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
