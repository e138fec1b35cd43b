"""A/B several library builds on a few prefill shapes, each build in its own process, interleaved twice:
    python tools/ab_shapes.py '' prev nohint       ('' = product build; names = build/libpli_attention_<name>.so)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
import physics_llm_inference_b200 as pli
def run(B, Hq, Hkv, N, D, causal, dtype=torch.bfloat16, reps=20):
    reps = max(reps, int(reps * 8192 / N))
    q = torch.randn(B, Hq, N, D, device="cuda").to(dtype); k = torch.randn(B, Hkv, N, D, device="cuda").to(dtype); v = torch.randn(B, Hkv, N, D, device="cuda").to(dtype)
    for _ in range(3): pli.flash_attention_forward(q, k, v, causal=causal)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): pli.flash_attention_forward(q, k, v, causal=causal)
    e1.record(); torch.cuda.synchronize()
    return pli.prefill_algorithmic_flops(B, Hq, N, N, D, causal) / (e0.elapsed_time(e1) / reps) / 1e9
time.sleep(0.5)
out = []
only = os.environ.get("PLI_AB_ONLY", "").split(",") if os.environ.get("PLI_AB_ONLY") else None      # e.g. PLI_AB_ONLY=C2,D64
def add(name, *a, **kw):
    if only is None or name in only:
        out.append(name + " %%6.0f" %% run(*a, **kw))
add("C2", 4, 32, 8, 8192, 128, True)
add("N2048", 16, 32, 8, 2048, 128, True)
add("N512", 64, 32, 8, 512, 128, True)
add("N128", 256, 32, 8, 128, 128, True)
add("noncausal", 4, 32, 8, 8192, 128, False, reps=10)
add("fp16", 4, 32, 8, 8192, 128, True, torch.float16)
add("D64", 4, 32, 8, 8192, 64, True)
add("D64nc", 4, 32, 8, 8192, 64, False, reps=10)
add("MHA", 4, 32, 32, 8192, 128, True)
print("  ".join(out))
''' % ROOT
for rep in range(2):
    for variant in sys.argv[1:]:
        env = dict(os.environ)
        if variant:
            env["PLI_LIB_PATH"] = os.path.join(ROOT, "physics_llm_inference_b200", "build", f"libpli_attention_{variant}.so")
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
        print(f"[{variant or 'product':8s}] {r.stdout.strip()} {r.stderr.strip()[-300:] if r.returncode else ''}", flush=True)
