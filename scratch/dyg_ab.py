"""A/B of the token-linear engine policies on one DyGFormer forward (same box, interleaved)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from tgm_b200 import _cabi
from tgm_b200.nn import DyGFormer
DEV = torch.device('cuda', 0)
torch.manual_seed(0); rng = np.random.default_rng(0)
N, B, L, dN, dE, dT, C, out = 100_000, 400, 32, 128, 16, 100, 50, 172
m = DyGFormer(dN, dE, dT, C, output_dim=out, patch_size=1, num_layers=2, num_heads=2, max_input_sequence_length=L).to(DEV).eval()
k = L - 1
node_x = torch.randn(N, dN, device=DEV)
B2 = B // 2
src, dst = rng.integers(0, N, B2), rng.integers(0, N, B2)
t = rng.integers(10_000, 2_000_000, B2)
nbrs = rng.integers(0, N, (2 * B2, k)).astype(np.int32)
nt = np.sort(np.clip(np.tile(t, 2)[:, None] - rng.integers(1, 9000, (2 * B2, k)), 0, None), 1)
ef = rng.standard_normal((2 * B2, k, dE)).astype(np.float32)
dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
args = (node_x, dv(np.stack([src, dst])), dv(t), dv(nbrs), dv(nt), dv(ef))
def run(n):
    with torch.no_grad():
        for _ in range(n): y = m(*args)
    return y
names = {0: 'CUTLASS FastF32 collective for every token linear', 2: 'hand-written kernel for the GELU-fused linear, collective for the rest', 1: 'hand-written tcgen05 kernel for every token linear (default)'}
ref = None
for rep in range(2):
    for mode in (1, 2, 0):
        _cabi.check(_cabi.lib.tgm_set_option(b'tc_linear', mode))
        y = run(5); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(50); e1.record(); torch.cuda.synchronize()
        y0 = y[0] if isinstance(y, (tuple, list)) else y
        if ref is None: ref = y0.clone()
        print(f'rep {rep} tc_linear={mode} ({names[mode]}): {e0.elapsed_time(e1) / 50:.4f} ms/forward, max diff vs first {float((y0 - ref).abs().max()):.2e}', flush=True)
_cabi.check(_cabi.lib.tgm_set_option(b'tc_linear', 1))
