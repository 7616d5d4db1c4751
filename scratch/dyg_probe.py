import sys, numpy as np, torch
sys.path.insert(0, '.')
from tgm_b200.nn import DyGFormer
DEV = torch.device('cuda', 0)
torch.manual_seed(0); rng = np.random.default_rng(0)
N, B, L, dN, dE, dT, C, out = 100_000, 200, 32, 128, 16, 100, 50, 172
m = DyGFormer(dN, dE, dT, C, output_dim=out, patch_size=1, num_layers=2, num_heads=2, max_input_sequence_length=L).to(DEV).eval()
k = L - 1
node_x = torch.randn(N, dN, device=DEV)
src, dst = rng.integers(0, N, B), rng.integers(0, N, B)
t = rng.integers(10_000, 2_000_000, B)
nbrs = rng.integers(0, N, (2 * B, k)).astype(np.int32)
nt = np.sort(np.clip(np.tile(t, 2)[:, None] - rng.integers(1, 9000, (2 * B, k)), 0, None), 1)
ef = rng.standard_normal((2 * B, k, dE)).astype(np.float32)
dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
args = (node_x, dv(np.stack([src, dst])), dv(t), dv(nbrs), dv(nt), dv(ef))
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
        m(*args)
torch.cuda.synchronize()
