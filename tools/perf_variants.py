"""Time C2 prefill (causal + non-causal) with several builds of the library (tuning aid).
    python tools/perf_variants.py [variant ...]     # '' = product build, else build/libpli_attention_<variant>.so
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import torch
import physics_llm_inference_b200 as pli
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
out = []
for causal in (True, False):
    for _ in range(3):
        pli.flash_attention_forward(q, k, v, causal=causal)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pli.flash_attention_forward(q, k, v, causal=causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out.append(f"causal={int(causal)} {ms:.3f} ms {pli.prefill_algorithmic_flops(B, Hq, N, N, D, causal) / ms / 1e9:.0f} TFLOP/s")
print("  ".join(out))
''' % ROOT

for variant in (sys.argv[1:] or [""]):
    env = dict(os.environ)
    if variant:
        env["PLI_LIB_PATH"] = os.path.join(ROOT, "physics_llm_inference_b200", "build", f"libpli_attention_{variant}.so")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(f"[{variant or 'product'}] {r.stdout.strip()} {r.stderr.strip()[-300:] if r.returncode else ''}", flush=True)
