"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): head-sharded decode with the output all-gather
fused into the kernel's peer stores == NCCL gather of the same shards == the unsharded kernel."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    import physics_llm_inference_b200 as pli
    from oracle import attention_oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    pli.init_distributed("nccl")
    dev = torch.device("cuda", rank)
    B, Hq, Hkv, D, bs = 6, 16, 4, 128, 16
    lens_l = [513, 77, 1024, 300, 16, 999]
    q, kp, vp, table, lens = orc.seeded_paged(41, B, Hq, Hkv, D, bs, lens_l, dtype=torch.bfloat16)
    q, kp, vp, table, lens = q.to(dev), kp.to(dev), vp.to(dev), table.to(dev), lens.to(dev)
    full = pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=max(lens_l))[:, :, 0]
    shard = pli.make_shard(rank, world, Hq, Hkv, B)
    qs = q[:, shard.q_start:shard.q_end]
    kps, vps = kp[:, :, :, shard.kv_start:shard.kv_end], vp[:, :, :, shard.kv_start:shard.kv_end]
    po = pli.PeerOutput(B, Hq, D, torch.bfloat16, shard)
    ok, why = True, ""
    for step in range(4):
        for splits in (None, 2):
            kw = dict(block_tables=table, max_seq_len=max(lens_l), num_splits=splits)
            ref = pli.gather_heads(pli.flash_decode(qs, kps, vps, lens, **kw)[:, :, 0], shard)   # NCCL all-gather
            o = pli.flash_decode(qs, kps, vps, lens, peer_out=po, **kw)                         # fused peer stores
            if not torch.equal(o, ref):
                ok, why = False, f"step {step} splits {splits}: fused != gathered, max diff {(o.float() - ref.float()).abs().max().item()}"
            if (ref.float() - full.float()).abs().max().item() > 1e-2:      # split counts differ: not bit-equal
                ok, why = False, f"step {step} splits {splits}: sharded != unsharded"
    # the same step captured in a CUDA graph (the step counter and the buffer parity are read on the device)
    kw = dict(block_tables=table, max_seq_len=max(lens_l))
    ref = pli.gather_heads(pli.flash_decode(qs, kps, vps, lens, **kw)[:, :, 0], shard)
    ws = pli.decode_workspace(qs.shape[0], qs.shape[1], D, pli.decode_num_splits(qs.shape[0], kps.shape[3], max(lens_l)), dev)
    # ONE step per graph with a CONSUMER captured in the same graph (stands for o_proj / the next layer): it must see the
    # step's output on every replay.  The single-launch gather writes one buffer at a fixed address, so nothing special is
    # needed (round 1 returned an alternating double buffer here; the two-buffer protocol still needs graph_safe).
    torch.cuda.synchronize()
    dist.barrier()
    if po.mode != "gather":
        ok, why = False, f"the TMA decode path should use the single-launch gather, mode is {po.mode}"
    pog = pli.PeerOutput(B, Hq, D, torch.bfloat16, shard)
    consumed = torch.zeros(B, Hq, D, device=dev, dtype=torch.float32)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        pli.flash_decode(qs, kps, vps, lens, peer_out=pog, workspace=ws, **kw)         # warm-up outside the graph
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            og = pli.flash_decode(qs, kps, vps, lens, peer_out=pog, workspace=ws, **kw)
            consumed.copy_(og.float() * 2)
    torch.cuda.current_stream(dev).wait_stream(side)
    for rep in range(5):
        consumed.zero_()
        graph.replay()
        o = pog.advance()
        torch.cuda.synchronize()
        if not torch.equal(o, ref):
            ok, why = False, f"graph replay {rep}: fused != gathered"
        if not torch.equal(consumed, ref.float() * 2):
            ok, why = False, f"graph replay {rep}: the consumer captured in the graph read a stale buffer"
    # prefill: O tiles TMA-stored into every rank's full output
    Bp, N = 2, 640
    qf, kf, vf = orc.seeded_qkv(43, Bp, Hq, Hkv, N, N, D)
    qf, kf, vf = qf.to(dev).bfloat16(), kf.to(dev).bfloat16(), vf.to(dev).bfloat16()
    shp = pli.make_shard(rank, world, Hq, Hkv, Bp)
    qs2, ks2, vs2 = pli.shard_kv_heads(qf, kf, vf, shp)
    ref = pli.gather_heads(pli.flash_attention_forward(qs2, ks2, vs2, causal=True), shp)
    pop = pli.PeerOutput(Bp, Hq, D, torch.bfloat16, shp, seq_len=N)
    for step in range(3):
        o = pli.flash_attention_forward(qs2, ks2, vs2, causal=True, peer_out=pop)
        if o.shape != ref.shape or not torch.equal(o, ref):
            ok, why = False, f"prefill step {step}: fused != gathered"
    if not ok:
        print(f"[rank {rank}] {why}", flush=True)
    torch.cuda.synchronize()
    torch.save(torch.tensor(int(ok)), os.path.join(out_dir, f"ok{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_decode_with_fused_gather(tmp_path):
    world = 2 if torch.cuda.device_count() < 4 else 4
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(torch.load(os.path.join(tmp_path, f"ok{r}.pt"))) == 1
