"""Human-readable digest of a bench.py JSON line:  python tools/show_bench.py out.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
r = lambda x: round(x, 3) if isinstance(x, float) else x  # noqa: E731
print(f"{d['metric']} = {d['value']:.1f} {d['unit']}  n_gpus {d['n_gpus']}  {d['ms_per_step']:.3f} ms/step  scaling {d['scaling']}  "
      f"frac {d['roofline']['frac']:.3f}  clocks {d['clocks']}")
print("workload:", d["config"]["workload"])
if "parity" in d:
    print("parity:", d["parity"])
e = d["e2e"]
print("e2e:", {k: r(v) for k, v in e.items() if k != "note"})
for key in ("sustained", "weak_c2"):
    if key in d:
        print(key + ":", {k: r(v) for k, v in d[key].items() if k not in ("config",)})
if "c4" in d:
    print("c4:", {k: r(v) for k, v in d["c4"].items()})
if "decode" in d:
    x = d["decode"]
    print(f"decode C3: {x['value']:.0f} GB/s  {x['us_per_step']:.1f} us  frac {x['roofline']['frac']:.3f}  launches {x['gpu_launches']}")
    for row in x.get("cases", []):
        print(f"   {row['workload']:62s} splits {row['num_splits']:2d}  eager {row['eager_us']:6.1f} us {row['eager_gbs']:5.0f} GB/s  "
              f"graph {row['graph_us']:6.1f} us {row['graph_gbs']:5.0f} GB/s  parity {row['parity']['ok']}")
    if "cpu_baseline" in x:
        print("   cpu:", r(x["cpu_baseline"]["value"]), x["cpu_baseline"]["unit"], x["cpu_baseline"]["cores"], "threads")
for row in d.get("shapes", []):
    print(f"  shape {row['workload']:58s} {row['ms']:8.3f} ms {row['tflops']:7.1f} TFLOP/s  frac {row['roofline']['frac']:.3f}")
if "cpu_baseline" in d:
    c = d["cpu_baseline"]
    print("cpu_baseline:", r(c["value"]), "best", r(c.get("best")), c["unit"], c["cores"], "threads;", c["pass_times_s"])
s = d.get("strong_scaling_configs", {})
if "c4_prefill_65536" in s:
    print("C4:", {k: r(v) for k, v in s["c4_prefill_65536"].items() if k != "workload"})
for row in s.get("c5_decode_b256", {}).get("sweep", []):
    print("  C5", {k: r(v) for k, v in row.items() if k not in ("l2_note", "kv_bytes_per_gpu")})
