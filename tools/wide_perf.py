"""C2 causal with the wide prefill kernel and debug flags: python tools/wide_perf.py FLAGS [FLAGS ...]
(4 = wide kernel with three softmax warpgroups, 36 = with two, +8 = null softmax (needs a -DPLI_PROFILE=1 build), 0 = default kernel)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib
lib = _lib.load()
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16(); k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16(); v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
fl = pli.prefill_algorithmic_flops(B, Hq, N, N, D, True)
def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): pli.flash_attention_forward(q, k, v, causal=True)
    e1.record(); torch.cuda.synchronize()
    return fl / (e0.elapsed_time(e1) / n) / 1e9
for rep in range(2):
    for flags in [int(a) for a in sys.argv[1:]]:
        lib.pli_debug_prefill_trace(None, 0, flags)
        time.sleep(1.0)
        for _ in range(3): pli.flash_attention_forward(q, k, v, causal=True)
        torch.cuda.synchronize()
        print(f"flags {flags:3d}: burst {timed(20):7.1f} TFLOP/s", flush=True)
lib.pli_debug_prefill_trace(None, 0, 0)
