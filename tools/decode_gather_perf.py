"""C5 ctx 1024 share of one GPU with the output gather fused in, N ranks (torchrun): decode alone vs decode + NCCL all-gather
vs the single-launch fused gather, eager (DecodePlan) and as 10-step CUDA graphs.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/decode_gather_perf.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import physics_llm_inference_b200 as pli

rank, world, local = pli.init_distributed("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
Hq, Hkv, D, B, bs = 32, 8, 128, 256, 16
shard = pli.make_shard(rank, world, Hq, Hkv, B)
hq_l, hkv_l = shard.q_end - shard.q_start, shard.kv_end - shard.kv_start


def timed(fn, reps):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for L in (1024, 8192):
    pages = B * L // bs
    kp = torch.randn(pages, 1, bs, hkv_l, D, device=dev).bfloat16()
    vp = torch.randn(pages, 1, bs, hkv_l, D, device=dev).bfloat16()
    table = torch.randperm(pages).to(torch.int32).view(B, L // bs).to(dev)
    lens = torch.full((B,), L, dtype=torch.int32, device=dev)
    q = torch.randn(B, hq_l, 1, D, device=dev).bfloat16()
    S = pli.decode_num_splits(B, hkv_l, L)
    ws = pli.decode_workspace(B, hq_l, D, S, dev)
    od = torch.empty(B, hq_l, D, device=dev, dtype=torch.bfloat16)
    po = pli.PeerOutput(B, Hq, D, torch.bfloat16, shard, device=dev)
    plain = pli.DecodePlan(q, kp, vp, lens, block_tables=table, max_seq_len=L, workspace=ws, out=od)
    fused = pli.DecodePlan(q, kp, vp, lens, block_tables=table, max_seq_len=L, workspace=ws, peer_out=po)
    both = lambda: (plain(), pli.gather_heads(od, shard))  # noqa: E731
    plain()
    ref = pli.gather_heads(od, shard)
    eq = all(torch.equal(fused(), ref) for _ in range(3))

    def graphed(step, after=None, n=10):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step(); torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for _ in range(n): step()
        torch.cuda.current_stream(dev).wait_stream(side)

        def replay():
            g.replay()
            if after: after()
        return timed(replay, 5) / n

    res = {"decode": timed(plain, 20), "decode+nccl": timed(both, 20), "fused": timed(fused, 20),
           "graph decode": graphed(plain), "graph decode+nccl": graphed(both), "graph fused": graphed(fused, lambda: po.advance(10))}
    eq2 = torch.equal(po.buffer(0), ref)
    nbytes = (2 * B * L * Hkv * D * 2 + 2 * B * Hq * D * 2 + 4 * B * (L // bs)) / world
    if rank == 0:
        print(f"ctx {L} world {world} (S={S}) fused==nccl {eq and eq2}: " + "  ".join(f"{k} {v:.1f} us" for k, v in res.items()) +
              f" | per GPU: graph decode {nbytes / res['graph decode'] / 1e3:.0f} GB/s, graph fused {nbytes / res['graph fused'] / 1e3:.0f} GB/s",
              flush=True)
    del kp, vp
dist.barrier()
dist.destroy_process_group()
