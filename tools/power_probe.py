"""Sustained-load probe: clocks / power / throttle reasons while the prefill kernel runs for seconds."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
import physics_llm_inference_b200 as pli

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
samples = []
stop = False
def sampler():
    while not stop:
        try:
            samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        except Exception as e:
            samples.append((time.time(), -1, -1, repr(e)))
        time.sleep(0.02)
th = threading.Thread(target=sampler, daemon=True); th.start()
fl = pli.prefill_algorithmic_flops(B, Hq, N, N, D, True)
for _ in range(5):
    pli.flash_attention_forward(q, k, v, causal=True)
torch.cuda.synchronize()
t_start = time.time()
for win in range(int(os.environ.get("PLI_WINDOWS", "12"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100):
        pli.flash_attention_forward(q, k, v, causal=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    now = time.time()
    recent = [s for s in samples if s[0] > now - 0.15]
    clk = sorted(s[1] for s in recent)[len(recent) // 2] if recent else -1
    pw = max((s[2] for s in recent), default=-1)
    rs = 0
    for s in recent:
        if isinstance(s[3], int): rs |= s[3]
    print(f"t={now - t_start:5.2f}s  {ms:.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  sm_clk {clk} MHz  power {pw:.0f} W  reasons 0x{rs:x}", flush=True)
stop = True
