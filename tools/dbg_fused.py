import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib
from oracle import attention_oracle as orc
big = torch.empty(int(40e9), dtype=torch.uint8, device="cuda"); del big       # the caching allocator now owns a 40 GB segment
B, G, D, L, splits = 1, 4, 128, 32768, 37
Hkv, bs = 2, 16
Hq = Hkv * G
q, kp, vp, table, lens = orc.seeded_paged(91, B, Hq, Hkv, D, bs, [L], dtype=torch.bfloat16)
qd, kd, vd, td, ld = q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens.cuda()
lib = _lib.load()
need = int(lib.pli_decode_workspace_bytes(B, Hq, D, splits))
print("need", need, "table", tuple(td.shape), "pools", tuple(kd.shape), flush=True)
ws2 = torch.empty(need // 4 + 2, dtype=torch.float32, device="cuda")
o2 = torch.empty(B, Hq, D, device="cuda", dtype=torch.bfloat16)
l2 = torch.empty(B, Hq, device="cuda", dtype=torch.float32)
stream = torch.cuda.current_stream().cuda_stream
q3 = qd[:, :, 0, :]
_lib.check(lib.pli_decode_splitkv(q3.data_ptr(), kd.data_ptr(), vd.data_ptr(), td.data_ptr(), ld.data_ptr(), B, Hq, Hkv, D,
                                  L, bs, td.stride(0), 0, kd.shape[0], _lib.i64(q3.stride(0), q3.stride(1)),
                                  _lib.i64(*kd.stride()[:4]), D ** -0.5, _lib.dtype_code(qd.dtype), splits,
                                  ws2.data_ptr(), ws2.numel() * 4, stream))
torch.cuda.synchronize(); print("splitkv ok", flush=True)
_lib.check(lib.pli_decode_combine(ws2.data_ptr(), o2.data_ptr(), l2.data_ptr(), B, Hq, D, splits,
                                  _lib.i64(o2.stride(0), o2.stride(1)), _lib.dtype_code(qd.dtype), stream))
torch.cuda.synchronize(); print("combine ok", flush=True)
for S in (37, 32, 5):
    ws = torch.empty(int(lib.pli_decode_workspace_bytes(B, Hq, D, S)) // 4 + 2, dtype=torch.float32, device="cuda")
    ws.view(torch.uint8).fill_(0xFF)
    o = pli.flash_decode(qd, kd, vd, ld, block_tables=td, max_seq_len=L, num_splits=S, workspace=ws)
    torch.cuda.synchronize(); print("fused ok", S, (o.float()[:, :, 0] - o2.float()).abs().max().item(), flush=True)
