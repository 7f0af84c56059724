"""oracle/raster.py replays the reference rasterisers' outputs (tests/golden/raster.npz)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import raster as orr


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold(golden_dir):
    z = np.load(os.path.join(golden_dir, "raster.npz"))
    return z, json.loads(str(z["meta"]))


def page(meta_page):
    words, lines = orr.synth_page(meta_page["seed"], meta_page["gh"], meta_page["gw"], meta_page["n_words"])
    if meta_page["tag"] == "odd":
        words["chars"][3] = np.zeros(0, np.int32)
        words["chars"][7] = np.zeros(0, np.int32)
    return words, lines


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_r1_r2_r3_bit_exact(gold, idx):
    z, meta = gold
    mp = meta["pages"][idx]
    tag = mp["tag"]
    words, lines = page(mp)
    grid, label = orr.raster_word_chargrid(words, lines, np.eye(96))
    g1 = torch.Tensor(grid).numpy()
    assert g1.shape[1:] == (mp["gh"], mp["gw"])
    assert sha(g1) == str(z[f"{tag}::r1_sha"])
    assert sha(label) == str(z[f"{tag}::r1_label_sha"])
    assert (label == z[f"{tag}::r1_label"]).all()
    feats = np.random.RandomState(mp["seed"] + 100).randn(len(lines["x"]), mp["feat_dim"])
    grid2, label2 = orr.raster_box_grid(lines, feats)
    assert sha(torch.Tensor(grid2).numpy()) == str(z[f"{tag}::r2_sha"])
    assert sha(label2) == str(z[f"{tag}::r2_label_sha"])
    boxes = np.stack([words["x"], words["y"], words["x"] + words["w"], words["y"] + words["h"]], 1)
    charset = meta["charset"]
    tok = {t: i for i, t in enumerate(" $" + charset)}
    ids = []
    for ch in words["chars"]:
        text = "".join(charset[c - 2] for c in ch)
        text = "".join(c if not c.isdigit() else "0" for c in text)
        ids.append(np.array([tok.get(c, 1) for c in text], np.int32))
    r3 = orr.raster_kv_chargrid(boxes, ids)
    assert tuple(z[f"{tag}::r3_shape"]) == r3["input_mask"].shape
    assert sha(r3["input_mask"]) == str(z[f"{tag}::r3_input_sha"])
    assert sha(r3["line_id_mask"]) == str(z[f"{tag}::r3_line_sha"])
    assert sha(r3["character_id_mask"]) == str(z[f"{tag}::r3_char_sha"])
    assert (r3["scaled_boxes"] == z[f"{tag}::r3_boxes"]).all()
    assert r3["scale"] == z[f"{tag}::r3_scale_pad"][0] and r3["bg_pad"] == z[f"{tag}::r3_scale_pad"][1]
    if tag == "small":
        oh = orr.one_hot_nchw(r3["input_mask"], int(z["small::r3_n_token"]))
        assert sha(oh) == str(z["small::r3_onehot_sha"])
