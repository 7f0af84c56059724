"""Seeded synthetic inputs shared by tests, bench.py and tests/golden/make_golden.py.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All generators are CPU + seeded so that the
GPU box regenerates exactly the inputs the golden fixtures were made from.
"""
import numpy as np
import torch


def synth_input(channels: int, n_class: int, B: int, H: int, W: int, seed: int, occupancy: float = 0.15,
                labelled: float = 0.3):
    """Sparse one-hot page tensor [B,C,H,W] fp32 + label map [B,H,W] int64 (0 = ignore)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, channels, (B, H, W), generator=g)
    occ = torch.rand((B, H, W), generator=g) < occupancy
    x = torch.zeros(B, channels, H, W)
    x.scatter_(1, ids[:, None], occ[:, None].float())
    labels = torch.randint(0, n_class, (B, H, W), generator=g)
    labels = labels * (torch.rand((B, H, W), generator=g) < labelled)
    labels[:, 0, 0] = 1  # every page keeps at least one pixel
    return x, labels.long()


def class_map(seed: int, H: int, W: int, n_rect: int = 200, n_class: int = 5) -> np.ndarray:
    """uint8 class map: ``n_rect`` random rectangles painted in order + 3 % salt noise, so that
    closing has 1-px gaps to fill and the labelling sees many tiny components."""
    rng = np.random.RandomState(seed)
    m = np.zeros((H, W), np.uint8)
    for _ in range(n_rect):
        h, w = rng.randint(1, max(2, H // 8)), rng.randint(1, max(2, W // 5))
        y, x = rng.randint(0, H), rng.randint(0, W)
        m[y:y + h, x:x + w] = rng.randint(0, n_class)
    noise = rng.rand(H, W) < 0.03
    m[noise] = rng.randint(0, n_class, int(noise.sum()))
    return m
