"""A/B library builds on the long-context shard shapes (C4: N 65536): python tools/ab_c4.py '' pb8 ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %r)
import torch
import physics_llm_inference_b200 as pli
def run(B, Hq, Hkv, N, D=128, reps=3):
    q = torch.randn(B, Hq, N, D, device="cuda").bfloat16(); k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16(); v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
    for _ in range(2): pli.flash_attention_forward(q, k, v, causal=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): pli.flash_attention_forward(q, k, v, causal=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, pli.prefill_algorithmic_flops(B, Hq, N, N, D, True) / ms / 1e9
print("C4 full 32q/8kv %%.2f ms %%.0f TFLOP/s | C4 shard of 8 GPUs (4q/1kv) %%.2f ms %%.0f | N16384 B2 %%.2f ms %%.0f" %% (
    *run(1, 32, 8, 65536), *run(1, 4, 1, 65536, reps=10), *run(2, 32, 8, 16384, reps=5)))
''' % ROOT
for rep in range(2):
    for variant in sys.argv[1:]:
        env = dict(os.environ)
        if variant:
            env["PLI_LIB_PATH"] = os.path.join(ROOT, "physics_llm_inference_b200", "build", f"libpli_attention_{variant}.so")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
        print(f"[{variant or 'product':8s}] {r.stdout.strip()} {r.stderr.strip()[-300:] if r.returncode else ''}", flush=True)
