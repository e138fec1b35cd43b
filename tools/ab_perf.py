"""A/B several library builds on the SAME box: burst (20 launches) and sustained (last 0.6 s of 2 s) C2 causal.
    python tools/ab_perf.py '' prefix other ...      ('' = product build)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
import physics_llm_inference_b200 as pli
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16(); k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16(); v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
fl = pli.prefill_algorithmic_flops(B, Hq, N, N, D, True)
def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): pli.flash_attention_forward(q, k, v, causal=True)
    e1.record(); torch.cuda.synchronize()
    return fl / (e0.elapsed_time(e1) / n) / 1e9
time.sleep(1.0)
for _ in range(3): pli.flash_attention_forward(q, k, v, causal=True)
torch.cuda.synchronize()
burst = timed(20)
timed(700)
sus = timed(300)
print(f"burst {burst:7.1f}  sustained {sus:7.1f} TFLOP/s")
''' % ROOT
for rep in range(2):
    for variant in sys.argv[1:]:
        env = dict(os.environ)
        if "=" in variant:                      # NAME=VALUE: same library, different environment switch
            k_, v_ = variant.split("=", 1)
            env[k_] = v_
        elif variant:
            env["PLI_LIB_PATH"] = os.path.join(ROOT, "physics_llm_inference_b200", "build", f"libpli_attention_{variant}.so")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(f"[{variant or 'product':10s}] {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
