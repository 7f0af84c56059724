"""Shared builder of the ``_extract_value`` test inputs (the same seeded inputs tests/golden/make_golden.py ``golden_kv`` fed
to the reference): R3 masks from the numpy oracle rasteriser (itself pinned bit-exact by tests/golden/raster.npz), label lines
with the reference's scaled integer boxes, and the synthetic soft-max map."""
import json
import os

import numpy as np

from oracle import raster as orr
from oracle.synth import kv_pred_mask


def load_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "kv_extract.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta["cases"], meta["charset"]


def build_case(case, charset):
    words, _ = orr.synth_page(case["seed"], case["gh"], case["gw"], case["n_words"])
    boxes = np.stack([words["x"], words["y"], words["x"] + words["w"], words["y"] + words["h"]], 1)
    texts = ["".join(charset[c - 2] for c in ch) for ch in words["chars"]]
    tok = {t: i for i, t in enumerate(" " + "$" + charset)}
    ids = [np.array([tok.get(c, 1) for c in "".join(ch if not ch.isdigit() else "0" for ch in t)], np.int32) for t in texts]
    r3 = orr.raster_kv_chargrid(np.array([[int(v) for v in b] for b in boxes], np.float64), ids)
    label_lines = [dict(box=[int(v) for v in sb], text=t, type=0, value=0) for sb, t in zip(r3["scaled_boxes"], texts)]
    pm = kv_pred_mask(case["pred_seed"], r3["input_mask"].shape, [l["box"] for l in label_lines], case["n_class"], case["noise"])
    return r3["line_id_mask"], r3["character_id_mask"], label_lines, pm


def plain(v):
    if v is None or isinstance(v, str):
        return v
    if isinstance(v, (list, tuple)):
        return [plain(e) for e in v]
    return int(v)
