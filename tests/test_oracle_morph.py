"""oracle/morph.py replays the reference's morph_util outputs (tests/golden/morph.npz) and agrees
with SciPy (the third-party library the reference delegates to) on random maps."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import morph as omo
from oracle.synth import class_map


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold(golden_dir):
    z = np.load(os.path.join(golden_dir, "morph.npz"))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("idx", [0, 1])
def test_closing_ccl_bbox_bit_exact(gold, idx):
    z, meta = gold
    mp = meta[idx]
    tag, H, W = mp["tag"], mp["H"], mp["W"]
    m = class_map(mp["seed"], H, W)
    res = omo.postprocess_page(m, 5)
    for c in range(2, 5):
        closed, labels, bb = res[c]
        want = np.unpackbits(z[f"{tag}::{c}::closed"])[:H * W].reshape(H, W).astype(bool)
        assert (closed == want).all()
        assert labels.dtype == np.int32
        assert (labels == z[f"{tag}::{c}::labels"]).all()
        assert sha(labels) == str(z[f"{tag}::{c}::labels_sha"])
        assert (bb == z[f"{tag}::{c}::bboxes"]).all()


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_rect_filters_bit_exact(gold, idx):
    z, meta = gold
    mp = meta[idx]
    tag, H, W = mp["tag"], mp["H"], mp["W"]
    m = class_map(mp["seed"], H, W) > 2
    fns = dict(dil=omo.r_dilation, ero=omo.r_erosion, open=omo.r_opening)
    n = 0
    for key in z.files:
        parts = key.split("::")
        if parts[0] != tag or parts[1] not in fns:
            continue
        sh, sw = (int(v) for v in parts[2].split("x"))
        origin = eval(parts[3])
        got = fns[parts[1]](m, (sh, sw), origin)
        want = np.unpackbits(z[key])[:H * W].reshape(H, W).astype(bool)
        assert (got == want).all(), key
        n += 1
    assert n == 15


def test_against_scipy_random():
    from scipy import ndimage
    rng = np.random.RandomState(0)
    for H, W, p in ((17, 23, 0.5), (64, 64, 0.6), (50, 31, 0.3)):
        img = rng.rand(H, W) < p
        lab, n = omo.label4(img)
        want, nw = ndimage.label(img)
        assert n == nw and (lab == want).all()
        got = omo.find_objects(lab)
        assert got == [o for o in ndimage.find_objects(want)]
        for size, origin in (((1, 3), 0), ((3, 1), 0), ((4, 4), (1, -2)), ((5, 2), (-2, 0))):
            assert (omo.r_dilation(img, size, origin) == ndimage.maximum_filter(img, size, origin=origin, mode="constant")).all()
            assert (omo.r_erosion(img, size, origin) == ndimage.minimum_filter(img, size, origin=origin, mode="constant")).all()
