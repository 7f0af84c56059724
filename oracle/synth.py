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


def kv_pred_mask(seed: int, shape, boxes, n_class: int, noise: float = 0.01) -> np.ndarray:
    """Synthetic soft-max output [H, W, n_class] fp32 for the ``_extract_value`` tests: background (class 0 / 1) everywhere,
    then a random subset of the text-line boxes ([x1, y1, x2, y2], already in grid coordinates) painted with a foreground
    class; about a third of those are split between TWO classes at a random column (so that a line is claimed by more
    than one field), some are dilated by a pixel so neighbouring lines merge, and ``noise`` (1 %) salt noise gives the (1, 3)
    closing and the labelling something to do (without it the top-most component of a class is a painted box, which is what
    the reference's multi-line branch needs to get past its area test)."""
    rng = np.random.RandomState(seed)
    H, W = shape
    cls = (rng.rand(H, W) < 0.3).astype(np.int64)               # classes 0 / 1
    for x1, y1, x2, y2 in boxes:
        if rng.rand() < 0.45:
            continue
        c = rng.randint(2, n_class)
        g = rng.randint(0, 2)
        ya, yb, xa, xb = max(y1 - g, 0), y2 + g, max(x1 - g, 0), x2 + g
        if rng.rand() < 0.35 and x2 - x1 > 4:
            xs = rng.randint(x1 + 2, x2 - 1)
            c2 = 2 + (c - 2 + rng.randint(1, n_class - 2)) % (n_class - 2)
            cls[ya:yb, xa:xs] = c
            cls[ya:yb, xs:xb] = c2
        else:
            cls[ya:yb, xa:xb] = c
    salt = rng.rand(H, W) < noise
    cls[salt] = rng.randint(0, n_class, int(salt.sum()))
    p = rng.rand(H, W, n_class).astype(np.float32) * 0.2
    p[np.arange(H)[:, None], np.arange(W)[None, :], cls] += 1.0
    return (p / p.sum(-1, keepdims=True)).astype(np.float32)
