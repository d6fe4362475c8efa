"""Cross-check the vectorised permutohedral restatement against the independent scalar
transcription (both in oracle/permuto.py; PARITY UNPINNED vs the third-party original)."""
import numpy as np
import torch

from oracle import permuto as P

KW = dict(pos_dim=3, log2_hashmap_size=12, nr_levels=16, nr_feat_per_level=2,
          coarsest_scale=1.0, finest_scale=1e-4, init_scale=1e-5)


def test_vectorised_matches_scalar():
    g = torch.Generator().manual_seed(3)
    kw = dict(KW, init_scale=1.0)
    p = P.init_params(kw, g)
    tab, sh = p["_encoding.lattice_values"], p["_encoding.random_shift_per_level"]
    sc = P.scale_factors(kw)
    x = torch.rand(40, 3, generator=g)
    out = P.encode(x, tab, sh, sc)
    assert out.shape == (40, 32)
    for i in range(40):
        ref = P.encode_scalar(x[i].numpy(), tab.numpy(), sh.numpy(), sc.numpy())
        np.testing.assert_allclose(out[i].numpy(), ref, rtol=1e-5, atol=1e-6)


def test_partition_of_unity_and_continuity():
    """Barycentric weights sum to 1: a constant table must encode to that constant."""
    g = torch.Generator().manual_seed(4)
    p = P.init_params(KW, g)
    tab = torch.full_like(p["_encoding.lattice_values"], 0.25)
    x = torch.rand(500, 3, generator=g)
    out = P.encode(x, tab, p["_encoding.random_shift_per_level"], P.scale_factors(KW))
    assert torch.allclose(out, torch.full_like(out, 0.25), atol=1e-4)
    # coarse levels are continuous: a tiny step changes the coarse features only slightly
    p2 = P.init_params(dict(KW, init_scale=1.0), g)
    sc = P.scale_factors(KW)
    a = P.encode(x, p2["_encoding.lattice_values"], p2["_encoding.random_shift_per_level"], sc)
    b = P.encode(x + 1e-5, p2["_encoding.lattice_values"], p2["_encoding.random_shift_per_level"], sc)
    assert (a[:, :4] - b[:, :4]).abs().max() < 1e-2


def test_concat_points_width():
    g = torch.Generator().manual_seed(5)
    p = P.init_params(KW, g)
    x = torch.rand(7, 3, generator=g)
    out = P.encode(x, p["_encoding.lattice_values"], p["_encoding.random_shift_per_level"],
                   P.scale_factors(KW), concat_points=True, concat_points_scaling=2.0)
    assert out.shape == (7, 35) and torch.allclose(out[:, 32:], x * 2.0)
