"""Synthetic FUNSD-shaped page records for bench.py: seeded, CPU-only, no model code.

Two 1-character anchor cells pin the chargrid to exactly ``gh x gw`` cells of ``unit`` pixels; ``n_words`` random words of 1-9
characters follow; the text-line cells are the same boxes with labels idx mod 4.  tests/test_bench_inputs.py checks that this
generator and the one the CPU baseline uses (oracle/raster.py) produce the same pages, so both arms of bench.py time the same
workload."""
import numpy as np


def synth_page(seed: int, gh: int = 512, gw: int = 512, n_words: int = 198, unit: int = 8, n_chars_vocab: int = 94):
    """-> (words, lines): dicts of fp64 x / y / w / h arrays, ``chars`` (int32 feature rows per word, offset by the two reserved
    ids) and ``label`` (int32 per line)."""
    rng = np.random.RandomState(seed)
    boxes = [(0, 0, unit, unit), ((gw - 1) * unit - 1, (gh - 1) * unit - 1, unit, unit)]
    chars = [rng.randint(0, n_chars_vocab, 1), None]
    chars[1] = rng.randint(0, n_chars_vocab, 1)
    for _ in range(n_words):
        n = rng.randint(1, 10)
        w = int(unit * n * rng.uniform(2, 4))
        h = int(unit * rng.uniform(2, 6))
        x = rng.randint(0, gw * unit - w)
        y = rng.randint(0, gh * unit - h)
        boxes.append((x, y, w, h))
        chars.append(rng.randint(0, n_chars_vocab, n))
    b = np.array(boxes, np.float64)
    words = dict(x=b[:, 0].copy(), y=b[:, 1].copy(), w=b[:, 2].copy(), h=b[:, 3].copy(), chars=[c.astype(np.int32) + 2 for c in chars])
    lines = dict(x=b[:, 0].copy(), y=b[:, 1].copy(), w=b[:, 2].copy(), h=b[:, 3].copy(),
                 label=(np.arange(len(boxes)) % 4).astype(np.int32))
    return words, lines


def class_map_rects(seed: int, H: int, W: int, n_rect: int = 200, n_class: int = 5) -> np.ndarray:
    """uint8 class map for the post-process timings (synthetic variant): ``n_rect`` random rectangles of classes
    0..n_class-1 painted in order + 3 % salt noise (1-px gaps for the closing, thousands of tiny components for the labelling)."""
    rng = np.random.RandomState(seed)
    m = np.zeros((H, W), np.uint8)
    for _ in range(n_rect):
        h, w = rng.randint(1, max(2, H // 8)), rng.randint(1, max(2, W // 5))
        y, x = rng.randint(0, H), rng.randint(0, W)
        m[y:y + h, x:x + w] = rng.randint(0, n_class)
    noise = rng.rand(H, W) < 0.03
    m[noise] = rng.randint(0, n_class, int(noise.sum()))
    return m
