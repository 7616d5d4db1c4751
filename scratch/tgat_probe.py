import sys, numpy as np, torch
sys.path.insert(0, '.')
from tgm_b200.nn import TGAT
DEV = torch.device('cuda', 0)
rng = np.random.default_rng(0); torch.manual_seed(0)
N, D, TD, EMB, k, S0 = 9227, 172, 100, 172, 20, 600
model = TGAT(node_dim=1, edge_dim=D, time_dim=TD, embed_dim=EMB, num_layers=2, n_heads=2).to(DEV).eval()
node_x = torch.randn(N, 1, device=DEV)
sizes = [S0, S0 * k]; hop = {}
for h, S in enumerate(sizes):
    nid = rng.integers(0, N, (S, k)).astype(np.int32)
    pad = np.arange(k)[None, :] < rng.integers(0, k + 1, S)[:, None]
    nid[pad] = -1
    st = rng.integers(100_000, 2_600_000, S)
    nt = np.sort(np.clip(st[:, None] - rng.integers(1, 90_000, (S, k)), 0, None), 1); nt[pad] = 0
    ex = rng.standard_normal((S, k, D)).astype(np.float32); ex[pad] = 0
    seeds = rng.integers(0, N, S).astype(np.int32) if h == 0 else hop[0][2].reshape(-1)
    stt = st if h == 0 else hop[0][3].reshape(-1)
    hop[h] = (seeds, stt, nid, nt, ex)
dv = lambda i: [torch.from_numpy(np.ascontiguousarray(hop[h][i])).to(DEV) for h in range(2)]
args = (node_x, dv(0), dv(1), dv(2), dv(4), dv(3))
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        model(*args)
torch.cuda.synchronize()
