"""Timeline of CTA 0 of the prefill kernel (debug aid): python tools/trace_prefill.py [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib

flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
# argv[3], argv[4]: batch and sequence length (short sequences: "0 20 64 512" prints the FULL timeline of CTA 0's first items)
if len(sys.argv) > 4:
    B, N = int(sys.argv[3]), int(sys.argv[4])
SHORT = N <= 2048
q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
lib = _lib.load()
PAGED = os.environ.get("PLI_TRACE_PAGED") == "1"          # C2 read in place from 16-token pages (the kPaged instance)
if PAGED:
    bs_ = 16
    P_ = B * N // bs_
    kp_ = torch.randn(P_, 1, bs_, Hkv, D, device="cuda").bfloat16()
    vp_ = torch.randn(P_, 1, bs_, Hkv, D, device="cuda").bfloat16()
    table_ = torch.randperm(P_).to(torch.int32).view(B, N // bs_).cuda()
    lens_ = torch.full((B,), N, dtype=torch.int32, device="cuda")

    class _Paged:
        @staticmethod
        def flash_attention_forward(q_, k_, v_, causal=True):
            return _real.flash_attention_paged(q_, kp_, vp_, table_, lens_, max_seq_len=N)
    _real = pli
    pli = _Paged
for _ in range(2):
    pli.flash_attention_forward(q, k, v, causal=True)
cap = 20000
buf = torch.zeros(5 * cap * 2, dtype=torch.int64, device="cuda")
lib.pli_debug_prefill_trace(buf.data_ptr(), cap, flags)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
# argv[2] = number of untraced launches issued back to back in front of the traced one (the traced launch then
# runs under the clocks / power state of a burst of that length)
lead = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if lead:
    lib.pli_debug_prefill_trace(None, 0, 0)
    for _ in range(lead):
        pli.flash_attention_forward(q, k, v, causal=True)
    lib.pli_debug_prefill_trace(buf.data_ptr(), cap, flags)
e0.record()
pli.flash_attention_forward(q, k, v, causal=True)
e1.record()
torch.cuda.synchronize()
lib.pli_debug_prefill_trace(None, 0, 0)
ms = e0.elapsed_time(e1)
print(f"flags={flags} kernel {ms:.3f} ms (with tracing), after {lead} back-to-back launches")
h = buf.cpu().tolist()
recs = [(h[2 * i + 1], h[2 * i] & 0xFF, (h[2 * i] >> 8) & 0xFF, (h[2 * i] >> 16) & 0xFFFF) for i in range(5 * cap) if h[2 * i] >> 40]
recs.sort()
t0 = recs[0][0]
span = recs[-1][0] - t0
print(f"CTA 0: {span} SM cycles between its first and last event = {span / ms / 1e3:.0f} MHz if they span the kernel "
      f"(the SM clock the kernel actually ran at; nvidia-smi's samples are too coarse to see it)")
names = {1: "S_ready", 2: "max_done", 3: "P_posted", 4: "mma_inputs_ready", 5: "mma_issued", 6: "S_in_registers", 7: "exp_done", 8: "corr_start", 9: "corr_posted", 10: "epilogue_start", 11: "epilogue_done", 12: "q_k_landed"}
if SHORT:
    print("every event of CTA 0's first ~60k cycles (all items)")
    for clk, ev, t, j in recs:
        if clk - t0 < 60000:
            print(f"  {clk - t0:8d}  tile{t} j={j:3d} {names[ev]}")
    sys.exit(0)
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
    ld = [ev[(t, j)][6] - ev[(t, j)][1] for j in js if 6 in ev[(t, j)] and 1 in ev[(t, j)]]
    ex = [ev[(t, j)][7] - ev[(t, j)][2] for j in js if 7 in ev[(t, j)] and 2 in ev[(t, j)]]
    tail = [ev[(t, j)][3] - ev[(t, j)][7] for j in js if 7 in ev[(t, j)] and 3 in ev[(t, j)]]
    print(f"tile{t}: softmax phases: TMEM load {avg(ld):.0f} | max + scale post {avg(mx) - avg(ld):.0f} | exp2/pack/store issue {avg(ex):.0f} | "
          f"wait::st + fence + arrive {avg(tail):.0f}")
    p2go = [ev[(t, j)][4] - ev[(t, j)][3] for j in js if 4 in ev[(t, j)] and 3 in ev[(t, j)]]
    issue = [ev[(t, j)][5] - ev[(t, j)][4] for j in js if 5 in ev[(t, j)] and 4 in ev[(t, j)]]
    turn = [ev[(t, j + 1)][1] - ev[(t, j)][5] for j in js if (t, j + 1) in ev and 1 in ev[(t, j + 1)] and 5 in ev[(t, j)]]
    period = [ev[(t, j + 1)][1] - ev[(t, j)][1] for j in js if (t, j + 1) in ev and 1 in ev[(t, j + 1)] and 1 in ev[(t, j)]]
    cw = [ev[(t, j)][9] - ev[(t, j)][8] for j in js if 9 in ev[(t, j)] and 8 in ev[(t, j)]]
    clag = [ev[(t, j)][9] - ev[(t, j)][3] for j in js if 9 in ev[(t, j)] and 3 in ev[(t, j)]]
    cst = [ev[(t, j)][8] - ev[(t, j)][2] for j in js if 8 in ev[(t, j)] and 2 in ev[(t, j)]]
    print(f"tile{t}: correction warp: scale posted -> seen {avg(cst):.0f} | its share of P {avg(cw):.0f} | posted relative to the softmax warp's P_posted {avg(clag):+.0f}")
    print(f"tile{t}: softmax(S_ready->P_posted) {avg(soft):.0f}  of which ld+max {avg(mx):.0f} | P_posted->mma_inputs_ready {avg(p2go):.0f} | "
          f"issue {avg(issue):.0f} | issued->next S_ready {avg(turn):.0f} | period {avg(period):.0f} cycles")
