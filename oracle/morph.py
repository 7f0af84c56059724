"""numpy restatement of the rectangular morphology + connected-component post-process.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows inference/morph_util.py:13-22,65-84
and the head of ``KVModel._extract_value`` (inference/kv_model.py:162-177).

The arithmetic lives in a third-party dependency that is NOT under /root/reference:
**SciPy ``ndimage``** (``maximum_filter`` / ``minimum_filter`` / ``label`` / ``find_objects``;
requirements.txt is unpinned, the oracle container has scipy 1.18.1).  Its published behaviour,
restated here without calling it:

* ``maximum_filter(img, size=(sh,sw), origin=(oh,ow), mode='constant', cval=0)``: output[i,j] is the
  max over rows ``i - sh//2 + oh ... i - sh//2 + oh + sh - 1``... SciPy's origin shifts the window
  toward *lower* indices for positive origin: window rows = [i - sh//2 - oh, i - sh//2 - oh + sh - 1];
  out-of-image samples are 0.  ``minimum_filter`` likewise with min (so a border pixel whose
  window leaves the image always becomes 0).
* ``label(binary)`` default structure = 4-connectivity; int32 labels numbered 1.. in order of each
  component's first pixel in C-order raster scan.
* ``find_objects(labels)``: per label the half-open bounding slices (rows, cols).

``tests/test_oracle_morph.py`` checks this restatement against SciPy itself on random maps (SciPy is
in the image on both boxes) and against tests/golden/morph_*.npz made through the reference module.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def _rect_filter(img: np.ndarray, size, origin, is_max: bool) -> np.ndarray:
    sh, sw = (size, size) if np.isscalar(size) else size
    oh, ow = (origin, origin) if np.isscalar(origin) else origin
    H, W = img.shape
    r0 = -(sh // 2) - oh
    c0 = -(sw // 2) - ow
    pad_t, pad_b = max(-r0, 0), max(r0 + sh - 1, 0)
    pad_l, pad_r = max(-c0, 0), max(c0 + sw - 1, 0)
    p = np.zeros((H + pad_t + pad_b, W + pad_l + pad_r), dtype=img.dtype)
    p[pad_t:pad_t + H, pad_l:pad_l + W] = img
    out = None
    for dr in range(sh):
        for dc in range(sw):
            rs = pad_t + r0 + dr
            cs = pad_l + c0 + dc
            v = p[rs:rs + H, cs:cs + W]
            if out is None:
                out = v.copy()
            else:
                out = np.maximum(out, v) if is_max else np.minimum(out, v)
    return out


def r_dilation(image, size, origin=0):
    return _rect_filter(np.asarray(image), size, origin, True)


def r_erosion(image, size, origin=0):
    return _rect_filter(np.asarray(image), size, origin, False)


def r_opening(image, size, origin=0):
    return r_dilation(r_erosion(image, size, origin), size, origin)


def r_closing(image, size, origin=0):
    # morph_util.py:81-84 ignores its ``origin`` argument (passes 0 to both filters)
    return r_erosion(r_dilation(image, size, 0), size, 0)


def label4(binary: np.ndarray) -> Tuple[np.ndarray, int]:
    """4-connected labelling, labels numbered by first raster appearance (int32)."""
    fg = np.asarray(binary) != 0
    H, W = fg.shape
    parent = np.arange(H * W, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for i in range(H):
        row = fg[i]
        for j in range(W):
            if not row[j]:
                continue
            a = i * W + j
            if j > 0 and row[j - 1]:
                ra, rb = find(a), find(a - 1)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
            if i > 0 and fg[i - 1, j]:
                ra, rb = find(a), find(a - W)
                if ra != rb:
                    parent[max(ra, rb)] = min(ra, rb)
    labels = np.zeros(H * W, dtype=np.int32)
    nxt = 0
    root_label = {}
    flat = fg.ravel()
    for a in range(H * W):
        if flat[a]:
            r = find(a)
            if r not in root_label:
                nxt += 1
                root_label[r] = nxt
            labels[a] = root_label[r]
    return labels.reshape(H, W), nxt


def find_objects(labels: np.ndarray) -> List[Tuple[slice, slice]]:
    n = int(labels.max()) if labels.size else 0
    out = []
    for k in range(1, n + 1):
        ys, xs = np.nonzero(labels == k)
        out.append((slice(int(ys.min()), int(ys.max()) + 1), slice(int(xs.min()), int(xs.max()) + 1)))
    return out


def connected_components(image, thres=0):
    binary = image > thres if thres > 0 else image
    labels, _ = label4(binary)
    return labels, find_objects(labels)


def postprocess_page(pred_class: np.ndarray, n_class: int, size=(1, 3)):
    """kv_model.py:174-177 for every foreground class c in 2..n_class-1:
    closing -> label -> find_objects.  Returns {c: (closed bool map, labels int32, bboxes [n,4] y0,y1,x0,x1)}."""
    res = {}
    for c in range(2, n_class):
        closed = r_closing(pred_class == c, size)
        labels, objs = connected_components(closed)
        bb = np.array([[o[0].start, o[0].stop, o[1].start, o[1].stop] for o in objs], dtype=np.int32).reshape(-1, 4)
        res[c] = (closed, labels, bb)
    return res
