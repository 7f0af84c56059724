"""numpy restatement of the three box rasterisers that feed MSAU.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Array-in / array-out versions of:

* R1 ``get_box_mask_box_label_word`` .... data_generator_funsd_bert.py:149-186
      (+ ``get_min_max_x_y_w_h`` :49-61) -- word-level chargrid, text-line label mask
* R2 ``get_box_mask_box_label`` ......... data_generator_funsd_bert.py:64-93 -- one D-vector per box
* R3 ``KVModel._generate_masks_from_label`` inference/kv_model.py:83-148 -- inference chargrid
      (char-id / line-id / char-index uint16 masks)
* ``to_categorical`` + NCHW transposes ... inference/generic_util.py:94-95, kv_model.py:274-278

Python/numpy semantics that matter and are kept (hard part 5): true division in
float64, ``int()`` truncation toward zero, sequential float64 ``sum`` for the mean scaling ratio,
in-order overwrite (last writer wins), numpy slice clipping at the array edge, uint16 wrap.

Pinned by tests/golden/raster_*.npz (generated from the unmodified reference).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np


def _grid_geometry(x, y, w, h):
    """get_min_max_x_y_w_h (dgfb.py:49-61) + grid extent (dgfb.py:72-73 / 154-155)."""
    x = [float(v) for v in x]
    y = [float(v) for v in y]
    w = [float(v) for v in w]
    h = [float(v) for v in h]
    min_h, min_w = min(h), min(w)
    max_y = max(a + b for a, b in zip(y, h))
    max_x = max(a + b for a, b in zip(x, w))
    min_x, min_y = min(x), min(y)
    wn = int((max_x - min_x) / min_w) + 1
    hn = int((max_y - min_y) / min_h) + 1
    return min_x, min_y, min_w, min_h, hn, wn


def raster_word_chargrid(words: Dict[str, Sequence], lines: Dict[str, Sequence], feat_table: np.ndarray):
    """R1.  ``words``: x,y,w,h (len n) and ``chars`` = list of int arrays (row index into
    ``feat_table`` [n_rows, D] for every character of the word; an empty array = empty ocr text).
    ``lines``: x,y,w,h,label.  Returns (grid float64 [D,Hn,Wn], label uint8 [Hn,Wn])."""
    wx, wy, ww, wh = (np.asarray(words[k], dtype=np.float64) for k in "xywh")
    chars = words["chars"]
    min_x, min_y, min_w, min_h, hn, wn = _grid_geometry(wx, wy, ww, wh)
    ratios = [float(ww[i]) / len(chars[i]) if len(chars[i]) != 0 else 0 for i in range(len(ww))]
    acc = 0
    for r in ratios:  # builtin sum(): sequential, starts from int 0
        acc = acc + r
    mean_ratio = acc / len(ratios)
    ratios = [r if r != 0 else mean_ratio for r in ratios]
    min_scale = min(ratios)
    D = feat_table.shape[-1]
    grid = np.zeros((D, hn, wn))
    label = np.zeros((hn, wn)).astype("uint8")
    for i in range(len(ww)):
        nx = int((float(wx[i]) - min_x) / min_scale)
        ny = int((float(wy[i]) - min_y) / min_h)
        nw = max(int(float(ww[i]) / min_scale), 1)
        nh = max(int(float(wh[i]) / min_h), 1)
        n = len(chars[i]) if len(chars[i]) != 0 else nw
        pcw = max(int(nw / n), 1)
        for j, cid in enumerate(chars[i]):
            grid[:, ny:ny + nh, nx + pcw * j:nx + pcw * (j + 1)] = feat_table[cid][:, None, None]
    lx, ly, lw, lh = (np.asarray(lines[k], dtype=np.float64) for k in "xywh")
    for i in range(len(lw)):
        nx = int((float(lx[i]) - min_x) / min_w)
        ny = int((float(ly[i]) - min_y) / min_h)
        nw = max(int(float(lw[i]) / min_w), 1)
        nh = max(int(float(lh[i]) / min_h), 1)
        label[ny:ny + nh, nx:nx + nw] = int(lines["label"][i]) + 1
    return grid, label


def raster_box_grid(cells: Dict[str, Sequence], feats: np.ndarray):
    """R2.  ``cells``: x,y,w,h,label; ``feats`` [n, D].  Returns (grid float64 [D,Hn,Wn], label u8)."""
    cx, cy, cw, ch = (np.asarray(cells[k], dtype=np.float64) for k in "xywh")
    min_x, min_y, min_w, min_h, hn, wn = _grid_geometry(cx, cy, cw, ch)
    grid = np.zeros((feats.shape[-1], hn, wn))
    label = np.zeros((hn, wn)).astype("uint8")
    for i in range(len(cw)):
        nx = int((float(cx[i]) - min_x) / min_w)
        ny = int((float(cy[i]) - min_y) / min_h)
        nw = max(int(float(cw[i]) / min_w), 1)
        nh = max(int(float(ch[i]) / min_h), 1)
        grid[:, ny:ny + nh, nx:nx + nw] = feats[i][:, None, None]
        label[ny:ny + nh, nx:nx + nw] = int(cells["label"][i]) + 1
    return grid, label


def raster_kv_chargrid(boxes: np.ndarray, char_ids: Sequence[np.ndarray]):
    """R3.  ``boxes`` [n,4] = x1,y1,x2,y2 (page pixels); ``char_ids[i]`` = token ids of line i's text
    AFTER the host-side digit->'0' folding and tok_to_id lookup (kv_model.py:126,140).
    Returns dict(input_mask, line_id_mask, character_id_mask uint16 [H,W], scaled_boxes int64 [n,4],
    scale, bg_pad, bbox)."""
    boxes = np.asarray(boxes, dtype=np.float64)
    heights = boxes[:, 3] - boxes[:, 1]
    min_x, min_y = float(boxes[:, 0].min()), float(boxes[:, 1].min())
    max_x, max_y = float(boxes[:, 2].max()), float(boxes[:, 3].max())
    bbox = (min_x, min_y, max_x, max_y)
    median_h = float(np.median(heights))
    bg_pad = int(median_h * 3)
    min_x, min_y = min_x - bg_pad, min_y - bg_pad
    max_x, max_y = max_x + bg_pad, max_y + bg_pad
    scale = 3.0 / median_h
    w, h = max_x - min_x, max_y - min_y
    shape = [int(h * scale * 1.0), int(w * scale * 1.0)]
    input_mask = np.zeros(shape, dtype="uint16")
    line_mask = np.zeros(shape, dtype="uint16")
    char_mask = np.zeros(shape, dtype="uint16")
    scaled = np.zeros((len(boxes), 4), dtype=np.int64)
    for li in range(len(boxes)):
        x1, y1, x2, y2 = (float(v) for v in boxes[li])
        x1, y1, x2, y2 = x1 - min_x, y1 - min_y, x2 - min_x, y2 - min_y
        x1, y1, x2, y2 = int(x1 * scale * 1.0), int(y1 * scale * 1.0), int(x2 * scale * 1.0), int(y2 * scale * 1.0)
        scaled[li] = (x1, y1, x2, y2)
        ids = char_ids[li]
        if len(ids) > 0:
            cfw = max(1.0 * (x2 - x1) / len(ids), 1.0)
            cw = max(0.9 * cfw, 1.0)
            cw = min(cw, int((y2 - y1) * 1.2))
            line_mask[y1:y2, x1:x2] = li + 1
            for idx, cid in enumerate(ids):
                off = x1 + idx * cfw
                sx, ex = int(off), int(off + cw)
                input_mask[y1:y2, sx:ex] = cid
                line_mask[y1:y2, sx:ex] = li + 1
                char_mask[y1:y2, sx:ex] = idx + 1
    return dict(input_mask=input_mask, line_id_mask=line_mask, character_id_mask=char_mask,
                scaled_boxes=scaled, scale=scale, bg_pad=bg_pad, bbox=bbox)


def one_hot_nchw(ids: np.ndarray, n_token: int) -> np.ndarray:
    """generic_util.py:94-95 + kv_model.py:274-278: np.eye(n)[ids] -> [1, n_token, H, W] float32."""
    oh = np.eye(n_token, dtype="B")[ids]          # [H, W, n]
    return np.ascontiguousarray(oh.transpose(2, 0, 1))[None].astype(np.float32)


# ----------------------------------------------------------------------------- synthetic pages
def synth_page(seed: int, gh: int = 512, gw: int = 512, n_words: int = 198, unit: int = 8, n_chars_vocab: int = 94):
    """SURVEY.md section 8(d) c4 page generator: two 1-char anchor cells pin the grid to exactly
    gh x gw; ``n_words`` random words.  Returns (words, lines) dicts for R1 / R2 / R3."""
    rng = np.random.RandomState(seed)
    u = unit
    xs, ys, ws, hs, chars = [0], [0], [u], [u], [rng.randint(0, n_chars_vocab, 1)]
    xs.append((gw - 1) * u - 1); ys.append((gh - 1) * u - 1); ws.append(u); hs.append(u)
    chars.append(rng.randint(0, n_chars_vocab, 1))
    for _ in range(n_words):
        n = rng.randint(1, 10)
        w = int(u * n * rng.uniform(2, 4))
        h = int(u * rng.uniform(2, 6))
        x = rng.randint(0, gw * u - w)
        y = rng.randint(0, gh * u - h)
        xs.append(x); ys.append(y); ws.append(w); hs.append(h)
        chars.append(rng.randint(0, n_chars_vocab, n))
    words = dict(x=np.array(xs, np.float64), y=np.array(ys, np.float64), w=np.array(ws, np.float64),
                 h=np.array(hs, np.float64), chars=[c.astype(np.int32) + 2 for c in chars])
    lines = dict(x=words["x"].copy(), y=words["y"].copy(), w=words["w"].copy(), h=words["h"].copy(),
                 label=(np.arange(len(xs)) % 4).astype(np.int32))
    return words, lines
