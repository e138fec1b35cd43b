import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"])
for r in d["strong_scaling_configs"]["c5_decode_b256"]["sweep"]:
    print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items() if k.endswith("us") or k == "ctx"})
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["strong_scaling_configs"]["c4_prefill_65536"].items() if k != "workload"})
