"""Timeline of CTA 0 of the prefill kernel (debug aid): python tools/trace_prefill.py [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib

flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
lib = _lib.load()
for _ in range(2):
    pli.flash_attention_forward(q, k, v, causal=True)
cap = 20000
buf = torch.zeros(4 * cap * 2, dtype=torch.int64, device="cuda")
lib.pli_debug_prefill_trace(buf.data_ptr(), cap, flags)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pli.flash_attention_forward(q, k, v, causal=True)
e1.record()
torch.cuda.synchronize()
lib.pli_debug_prefill_trace(None, 0, 0)
print(f"flags={flags} kernel {e0.elapsed_time(e1):.3f} ms (with tracing)")
h = buf.cpu().tolist()
recs = [(h[2 * i + 1], h[2 * i] & 0xFF, (h[2 * i] >> 8) & 0xFF, (h[2 * i] >> 16) & 0xFFFF) for i in range(4 * cap) if h[2 * i] >> 40]
recs.sort()
t0 = recs[0][0]
names = {1: "S_ready", 2: "max_done", 3: "P_posted", 4: "mma_inputs_ready", 5: "mma_issued", 6: "mma: V landed", 7: "mma: O corrected", 8: "mma: P seen"}
print("first item: events of half-steps 40..44 (regions: softmax tile 0/1, MMA warp of tile 0/1)")
for clk, ev, t, j in recs:
    if 40 <= j <= 44 and clk - t0 < 800000:
        print(f"  {clk - t0:8d}  tile{t} j={j:3d} {names[ev]}")
# statistics over the first item: per-tile durations
import collections
ev = collections.defaultdict(dict)
first_item = [r for r in recs if r[0] - t0 < 10**9]
seen = set()
for clk, e, t, j in first_item:
    key = (e, t, j)
    if key in seen:      # later items reuse (t, j): keep the first item only
        continue
    seen.add(key)
    ev[(t, j)][e] = clk
def avg(xs):
    return sum(xs) / max(1, len(xs))
for t in (0, 1):
    js = sorted(j for (tt, j) in ev if tt == t and 10 <= j <= 110)
    soft = [ev[(t, j)][3] - ev[(t, j)][1] for j in js if 3 in ev[(t, j)] and 1 in ev[(t, j)]]
    mx = [ev[(t, j)][2] - ev[(t, j)][1] for j in js if 2 in ev[(t, j)] and 1 in ev[(t, j)]]
    p2go = [ev[(t, j)][4] - ev[(t, j)][3] for j in js if 4 in ev[(t, j)] and 3 in ev[(t, j)]]
    issue = [ev[(t, j)][5] - ev[(t, j)][4] for j in js if 5 in ev[(t, j)] and 4 in ev[(t, j)]]
    turn = [ev[(t, j + 1)][1] - ev[(t, j)][5] for j in js if (t, j + 1) in ev and 1 in ev[(t, j + 1)] and 5 in ev[(t, j)]]
    period = [ev[(t, j + 1)][1] - ev[(t, j)][1] for j in js if (t, j + 1) in ev and 1 in ev[(t, j + 1)] and 1 in ev[(t, j)]]
    print(f"tile{t}: softmax(S_ready->P_posted) {avg(soft):.0f}  of which ld+max {avg(mx):.0f} | P_posted->mma_inputs_ready {avg(p2go):.0f} | "
          f"issue {avg(issue):.0f} | issued->next S_ready {avg(turn):.0f} | period {avg(period):.0f} cycles")
