"""numpy restatement of ``KVModel._extract_value`` (inference/kv_model.py:151-261) and the helpers it calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned by tests/golden/kv_extract.npz, which holds the outputs of the
unmodified reference on the same seeded inputs (tests/golden/make_golden.py ``golden_kv``).

Helpers restated: ``area`` / ``ycenter`` (inference/morph_util.py:33-34, :55-56), ``union_boxes`` / ``intersect_boxes``
(:86-104), ``sort_box_reading_order`` (inference/generic_util.py:51-91).  ``np.argsort`` / ``np.unique`` / ``set`` are used
exactly where the reference uses them, because the result depends on their tie-breaking and iteration order.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import morph

MULTIPLE_LINES_FIELDS = (5, 11)      # kv_model.py:156


def _area(o) -> int:
    return (o[1].stop - o[1].start) * (o[0].stop - o[0].start)


def _ycenter(o):
    return np.mean([o[0].stop, o[0].start])


def union_boxes(boxes):
    if not boxes:
        return None
    x1, y1, x2, y2 = boxes[0]
    for b in boxes[1:]:
        x1, y1, x2, y2 = min(x1, b[0]), min(y1, b[1]), max(x2, b[2]), max(y2, b[3])
    return [x1, y1, x2, y2]


def intersect_boxes(boxes):
    if not boxes:
        return None
    x1, y1, x2, y2 = boxes[0]
    for b in boxes[1:]:
        x1, y1, x2, y2 = max(x1, b[0]), max(y1, b[1]), min(x2, b[2]), min(y2, b[3])
    return [x1, y1, x2, y2]


def reading_order(lines: List[Dict]) -> List[Dict]:
    """generic_util.py:51-91: repeatedly pull out the "top-left-most" remaining line.  Starting from the first remaining line,
    a candidate replaces the current pick when its centre lies at least half its own height above the pick's centre, or when
    its centre is left of the pick's right edge and above the pick's bottom edge."""
    rest = list(lines)
    out = []
    if not rest:
        return rest
    while len(rest) > 1:
        pick = rest[0]
        for cand in rest[1:]:
            px1, py1, px2, py2 = pick["box"]
            pcy = (py1 + py2) / 2
            x1, y1, x2, y2 = cand["box"]
            cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
            if cy <= pcy - (y2 - y1) / 2:
                pick = cand
                continue
            if cx < px2 and cy < py2:
                pick = cand
        out.append(pick)
        rest.remove(pick)
    out.append(rest[0])
    return out


def extract_value(line_mask, char_mask, label_lines, pred_mask, num_classes):
    """-> (values, new_pred_mask) exactly as kv_model.py:151-261."""
    pred_mask = np.asarray(pred_mask)
    n_class = pred_mask.shape[2]
    values = [("", None, None, None)] * n_class
    pred_class = np.argmax(pred_mask, axis=-1)
    new_pred_mask = np.zeros(pred_mask.shape)
    new_pred_mask[:, :, 0] = pred_mask[:, :, 0]
    used = [0] * (len(label_lines) + 1)
    line_ids_for_field = [[] for _ in range(num_classes + 1)]
    boxes_for_field = [[] for _ in range(num_classes + 1)]
    for i, l in enumerate(label_lines):
        l["id"] = i + 1
    for c in range(2, n_class):
        closed = morph.r_closing(pred_class == c, (1, 3))
        labels, objects = morph.connected_components(closed)
        if not objects:
            continue
        multi = c in MULTIPLE_LINES_FIELDS
        order = np.argsort([-_ycenter(o) for o in objects]) if multi else np.argsort([_area(o) for o in objects])
        best = order[-1]
        alts = []
        if _area(objects[best]) < 5:
            continue
        if multi and len(order) > 1:
            for k in order[:-1]:
                if _area(objects[k]) > 5:
                    alts.append(k)
                    o = objects[k]
                    boxes_for_field[c].append([o[1].start, o[0].start, o[1].stop, o[0].stop])
        o = objects[best]
        boxes_for_field[c].append([o[1].start, o[0].start, o[1].stop, o[0].stop])
        line_ids = [i for i in np.unique(line_mask[labels == best + 1]) if i > 0]
        for k in alts:
            line_ids += [i for i in np.unique(line_mask[labels == k + 1]) if i > 0]
            new_pred_mask[:, :, c][labels == k + 1] = 1
        line_ids_for_field[c] = list(set(line_ids))
        for i in line_ids:
            used[i] += 1
        new_pred_mask[:, :, c][labels == best + 1] = 1
    for c in range(2, n_class):
        line_ids = line_ids_for_field[c]
        if not line_ids:
            continue
        value = ""
        lines = reading_order([label_lines[i - 1] for i in line_ids if i > 0])
        line_boxes = []
        for line in lines:
            line_boxes.append(line["box"])
            if used[line["id"]] <= 1:
                value += line["text"]
            else:
                x1, y1, x2, y2 = line["box"]
                sel = set(np.unique(char_mask[y1:y2, x1:x2][new_pred_mask[:, :, c][y1:y2, x1:x2] > 0]))
                sel.discard(0)
                if not sel:
                    continue
                lo, hi = min(sel), max(sel)
                if hi > len(line["text"]) - 3:
                    hi = len(line["text"]) + 1
                value += line["text"][lo - 2 if lo >= 2 else 0: hi - 1]
            if c in MULTIPLE_LINES_FIELDS:
                value += "\n"
        if value and value[-1] == "\n":
            value = value[:-1]
        merged = union_boxes(line_boxes)
        values[c] = (value, [boxes_for_field[c][-1]], intersect_boxes(boxes_for_field[c] + [merged]),
                     union_boxes(boxes_for_field[c] + [merged]))
    return values, new_pred_mask
