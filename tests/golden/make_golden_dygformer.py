"""Golden vectors for DyGFormer from the UNMODIFIED reference module (tgm/nn/encoder/dygformer.py)
in eval() mode with seeded weights:  python tests/golden/make_golden_dygformer.py
The reference's own test is shape-only (test/unit/test_nn/test_dygformer.py:7-29).
Writes tests/golden/nn_dygformer_*.npz."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_shim import import_reference  # noqa: E402

import_reference()
from tgm.nn import DyGFormer  # noqa: E402


def run(name, N, B, L, dN, dE, dT, C, out_dim, P, layers, heads, seed, bias):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = DyGFormer(node_feat_dim=dN, edge_x_dim=dE, time_feat_dim=dT, channel_embedding_dim=C,
                      output_dim=out_dim, patch_size=P, num_layers=layers, num_heads=heads,
                      max_input_sequence_length=L).eval()
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if 'norm_layers' in n and n.endswith('weight'):
                prm.copy_(1 + 0.2 * torch.randn(prm.shape, generator=g))
            elif prm.ndim == 1 and (bias or 'time_encoder' not in n):
                prm.copy_(0.2 * torch.randn(prm.shape, generator=g))
    k = L - 1
    node_x = torch.randn(N, dN, generator=g)
    src = torch.randint(0, N, (B,), generator=g)
    dst = torch.randint(0, N, (B,), generator=g)
    t = torch.randint(1000, 2_000_000, (B,), generator=g)
    nbrs = torch.randint(0, max(4, N // 3), (2 * B, k), generator=g)  # few ids: co-occurrences
    nt = torch.sort((t.repeat(2)[:, None] - torch.randint(1, 900, (2 * B, k), generator=g)).clamp(min=0), 1)[0]
    ef = torch.randn(2 * B, k, dE, generator=g)
    npad = torch.randint(0, k + 1, (2 * B,), generator=g)
    npad[0], npad[1] = k, 0
    pad = torch.arange(k)[None, :] < npad[:, None]
    nbrs[pad], nt[pad], ef[pad] = -1, 0, 0.0
    with torch.no_grad():
        zs, zd = model(node_x, torch.stack([src, dst]), t, nbrs, nt, ef)
    sd = {'p.' + k_: v.numpy() for k_, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, f'nn_dygformer_{name}.npz'), node_x=node_x.numpy(),
                        src=src.numpy(), dst=dst.numpy(), t=t.numpy(), nbrs=nbrs.numpy().astype(np.int32),
                        nt=nt.numpy(), ef=ef.numpy(), z_src=zs.numpy(), z_dst=zd.numpy(),
                        patch_size=np.int64(P), num_layers=np.int64(layers), num_heads=np.int64(heads),
                        **sd)
    print(name, zs.shape, float(zs.abs().max()))


def run_grad(name, N, B, L, dN, dE, dT, C, out_dim, P, layers, heads, seed):
    """Gradients of every parameter from the reference's autograd: loss = sum(z_src * G_src) +
    sum(z_dst * G_dst), eval() mode (dropout off: a dropout mask cannot be reproduced across
    implementations).  Writes tests/golden/nn_dyggrad_*.npz."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = DyGFormer(node_feat_dim=dN, edge_x_dim=dE, time_feat_dim=dT, channel_embedding_dim=C,
                      output_dim=out_dim, patch_size=P, num_layers=layers, num_heads=heads,
                      max_input_sequence_length=L).eval()
    with torch.no_grad():
        for n, prm in model.named_parameters():
            if 'norm_layers' in n and n.endswith('weight'):
                prm.copy_(1 + 0.2 * torch.randn(prm.shape, generator=g))
            elif prm.ndim == 1:
                prm.copy_(0.2 * torch.randn(prm.shape, generator=g))
    k = L - 1
    node_x = torch.randn(N, dN, generator=g)
    src = torch.randint(0, N, (B,), generator=g)
    dst = torch.randint(0, N, (B,), generator=g)
    t = torch.randint(1000, 20_000, (B,), generator=g)
    nbrs = torch.randint(0, max(4, N // 3), (2 * B, k), generator=g)
    nt = torch.sort((t.repeat(2)[:, None] - torch.randint(1, 900, (2 * B, k), generator=g)).clamp(min=0), 1)[0]
    ef = torch.randn(2 * B, k, dE, generator=g)
    npad = torch.randint(0, k + 1, (2 * B,), generator=g)
    npad[0], npad[1] = k, 0
    pad = torch.arange(k)[None, :] < npad[:, None]
    nbrs[pad], nt[pad], ef[pad] = -1, 0, 0.0
    zs, zd = model(node_x, torch.stack([src, dst]), t, nbrs, nt, ef)
    Gs, Gd = torch.randn(zs.shape, generator=g), torch.randn(zd.shape, generator=g)
    ((zs * Gs).sum() + (zd * Gd).sum()).backward()
    sd = {'p.' + k_: v.detach().numpy() for k_, v in model.state_dict().items()}
    gr = {'g.' + n: prm.grad.numpy() for n, prm in model.named_parameters()}
    np.savez_compressed(os.path.join(HERE, f'nn_dyggrad_{name}.npz'), node_x=node_x.numpy(),
                        src=src.numpy(), dst=dst.numpy(), t=t.numpy(), nbrs=nbrs.numpy().astype(np.int32),
                        nt=nt.numpy(), ef=ef.numpy(), z_src=zs.detach().numpy(), z_dst=zd.detach().numpy(),
                        G_src=Gs.numpy(), G_dst=Gd.numpy(), patch_size=np.int64(P),
                        num_layers=np.int64(layers), num_heads=np.int64(heads), **sd, **gr)
    print('grad', name, {k_: float(np.abs(v).max()) for k_, v in list(gr.items())[:3]})


def main():
    # name, N, B, L, dN, dE, dT, C, out, P, layers, heads, seed, t2v bias
    run('small', 30, 6, 8, 3, 4, 6, 4, 5, 1, 2, 2, 1, False)
    run('patch2', 25, 5, 8, 2, 3, 4, 6, 7, 2, 1, 2, 2, True)
    run('seq32', 60, 4, 32, 4, 16, 10, 8, 12, 1, 2, 2, 3, False)
    run('seq32_patch4', 60, 4, 32, 4, 16, 10, 8, 12, 4, 2, 4, 4, False)
    run_grad('small', 30, 5, 8, 3, 4, 6, 4, 5, 1, 2, 2, 11)
    run_grad('patch2', 25, 4, 8, 2, 3, 4, 6, 7, 2, 1, 2, 12)
    run_grad('seq32', 60, 3, 32, 4, 16, 10, 8, 12, 1, 2, 2, 13)


if __name__ == '__main__':
    main()
