"""How the 20-step C2 burst number depends on what ran just before it: python tools/burst_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
g = torch.Generator().manual_seed(1)
host = [torch.randn(B, h, N, D, generator=g).to(torch.bfloat16).pin_memory() for h in (Hq, Hkv, Hkv)]
q, k, v = (t.cuda() for t in host)
fl = pli.prefill_algorithmic_flops(B, Hq, N, N, D, True)
def burst(warm, n=20, idle=2.0):
    torch.cuda.synchronize(); time.sleep(idle)
    for _ in range(warm): pli.flash_attention_forward(q, k, v, causal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): pli.flash_attention_forward(q, k, v, causal=True)
    e1.record(); torch.cuda.synchronize()
    return fl / (e0.elapsed_time(e1) / n) / 1e9
for rep in range(2):
    for idle in (2.0, 0.2):
        print(f"idle {idle:.1f} s: " + "  ".join(f"warm {w:3d}: {burst(w, idle=idle):6.0f}" for w in (0, 3, 10, 30, 100)), flush=True)
qd = torch.randn(B, Hq, N, D, device="cuda").bfloat16(); kd = torch.randn(B, Hkv, N, D, device="cuda").bfloat16(); vd = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
q, k, v = qd, kd, vd
print("device-generated inputs: " + "  ".join(f"warm {w:3d}: {burst(w):6.0f}" for w in (3, 10, 30)), flush=True)
print("long idle: " + "  ".join(f"idle {i:4.0f} s warm 3: {burst(3, idle=i):6.0f}" for i in (8.0, 15.0)), flush=True)
import threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
with bench.ClockSampler(0) as cs:
    r = burst(3, idle=2.0)
print(f"with bench.ClockSampler polling NVML every 10 ms: {r:6.0f}  {cs.summary()}", flush=True)
