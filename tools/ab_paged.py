"""A/B library builds on C2 read from 16-token pages against the contiguous call: python tools/ab_paged.py '' name ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time
sys.path.insert(0, %r)
import torch
import physics_llm_inference_b200 as pli
B, Hq, Hkv, N, D, bs = 4, 32, 8, 8192, 128, 16
P = B * N // bs
kp = torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16(); vp = torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16()
table = torch.randperm(P).to(torch.int32).view(B, N // bs).cuda(); lens = torch.full((B,), N, dtype=torch.int32, device="cuda")
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16(); k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16(); v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
fl = pli.prefill_algorithmic_flops(B, Hq, N, N, D, True)
def timed(fn, n=20):
    time.sleep(1.0)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return fl / (e0.elapsed_time(e1) / n) / 1e9
c = timed(lambda: pli.flash_attention_forward(q, k, v, causal=True))
p = timed(lambda: pli.flash_attention_paged(q, kp, vp, table, lens, max_seq_len=N))
print("contiguous %%6.0f  paged %%6.0f  ratio %%.3f" %% (c, p, p / c))
''' % ROOT
for rep in range(2):
    for variant in sys.argv[1:]:
        env = dict(os.environ)
        if "=" in variant:
            k_, v_ = variant.split("=", 1); env[k_] = v_
        elif variant:
            env["PLI_LIB_PATH"] = os.path.join(ROOT, "physics_llm_inference_b200", "build", f"libpli_attention_{variant}.so")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(f"[{variant or 'product':10s}] {r.stdout.strip()} {r.stderr.strip()[-200:] if r.returncode else ''}", flush=True)
