"""Seeded synthetic "omics-shaped" data generated directly on the device in feature-major layout.

Model: Z ~ N(0,1) n x r, r = K + 5; per block X_b = Z diag(s) L_b + noise E_b with
s_j = decay**j; Y = Z C + y_noise E_Y.  The feature axis is cut into fixed chunks whose RNG streams are
keyed by the *global* chunk index, so any sharding of the features over GPUs sees identical data.
This is the analogue of the reference's ``mbpls/data/get_data.py:orthogonal_data`` demo generator, not
part of the fit path.
"""
from __future__ import annotations

import torch

CHUNK = 4096


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def latent_factors(n, K, device, seed):
    r = K + 5
    Z = torch.randn((n, r), dtype=torch.float64, device=device, generator=_gen(device, seed))
    return Z


def response(n, q, K, device, seed, decay=0.7, y_noise=0.05):
    Z = latent_factors(n, K, device, seed)
    r = Z.shape[1]
    g = _gen(device, seed + 7)
    s = decay ** torch.arange(r, dtype=torch.float64, device=device)
    C = torch.randn((r, q), dtype=torch.float64, device=device, generator=g) * s[:, None]
    Y = Z @ C + y_noise * torch.randn((n, q), dtype=torch.float64, device=device, generator=g)
    return Y  # n x q (row-major), small


def fill_feature_major(Xt: torch.Tensor, n: int, g_lo: int, g_hi: int, K: int, seed: int, noise=0.1, decay=0.7,
                       nan_frac=0.0):
    """Fill rows [0, g_hi-g_lo) of ``Xt`` (feature-major, ld >= n) with global features [g_lo, g_hi)."""
    device = Xt.device
    Z = latent_factors(n, K, device, seed)
    r = Z.shape[1]
    s = decay ** torch.arange(r, dtype=torch.float64, device=device)
    Zs_t = (Z * s).t().contiguous()  # r x n
    # features per generator chunk: 4096, fewer for very long features so that a chunk's temporaries stay near 1 GB
    chunk = CHUNK if n <= 32768 else max(64, CHUNK * 32768 // n)
    c = (g_lo // chunk) * chunk
    while c < g_hi:
        g = _gen(device, seed * 1000003 + 17 + c // chunk)
        L = torch.randn((chunk, r), dtype=torch.float64, device=device, generator=g)
        E = torch.randn((chunk, n), dtype=torch.float64, device=device, generator=g)
        blk = torch.addmm(E, L, Zs_t, beta=noise)  # chunk x n
        if nan_frac > 0:
            m = torch.rand((chunk, n), dtype=torch.float32, device=device, generator=g) < nan_frac
            blk[m] = float("nan")
        a, b = max(c, g_lo), min(c + chunk, g_hi)
        Xt[a - g_lo:b - g_lo, :n] = blk[a - c:b - c]
        c += chunk
    if Xt.shape[1] > n:
        Xt[:g_hi - g_lo, n:] = 0.0
