"""Layer-1 TGAT attention at config-3 size (12 000 seeds x 20 neighbours, node 1, edge 172, time 100)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tgm_b200.nn import TemporalAttention, Time2Vec
dev = torch.device('cuda', 0)
torch.manual_seed(0)
S, k, ND, D, TD = int(os.environ.get('S', 12000)), 20, 1, 172, 100
att = TemporalAttention(2, ND, D, TD, dropout=0.0).to(dev).eval()
te = Time2Vec(TD).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
node_x = torch.randn(S, ND, device=dev, generator=g)
nbr = torch.randn(S, k, ND, device=dev, generator=g)
edge = torch.randn(S, k, D, device=dev, generator=g)
st = torch.randint(100_000, 2_600_000, (S,), device=dev, generator=g)
nt = (st[:, None] - torch.randint(1, 90_000, (S, k), device=dev, generator=g)).clamp_(min=0)
nid = torch.randint(0, 9000, (S, k), device=dev, generator=g, dtype=torch.int32)
nid[torch.rand(S, k, device=dev, generator=g) < 0.2] = -1
with torch.no_grad():
    for _ in range(3):
        out = att.forward_fused(te, node_x, nbr, edge, st, nt, nid)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        out = att.forward_fused(te, node_x, nbr, edge, st, nt, nid)
    b.record(); torch.cuda.synchronize()
print(f'forward_fused S={S}: {a.elapsed_time(b) / 20:.4f} ms  checksum {float(out.double().sum()):.6f}')
