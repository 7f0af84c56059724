"""CPU-side checks of the round-2 host logic (no GPU, no compute calls): page bucketing, host grid geometry vs the oracle,
post_process_kv vs the reference's table, the alternative trainer's optimiser plumbing, per-plan options through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import msau_b200
from msau_b200 import _lib, kv_model, train, training
from oracle import raster as orr


def test_page_grid_shape_equals_oracle_geometry():
    """The feeder's host geometry (data_generator_funsd_bert.py:49-61,72-73) must give the grid the rasterisers produce."""
    for seed, (gh, gw) in enumerate([(40, 48), (56, 40), (24, 64), (512, 512), (37, 61)]):
        words, lines = orr.synth_page(300 + seed, gh, gw, 20)
        assert train.page_grid_shape(words) == (gh, gw)
        grid, label = orr.raster_word_chargrid(words, lines, np.eye(96))
        assert grid.shape[1:] == (gh, gw) and label.shape == (gh, gw)
    # a page whose extent is not a multiple of its smallest box: the same truncating division as the reference
    odd = dict(x=[0.0, 13.0, 50.5], y=[0.0, 7.25, 31.0], w=[7.0, 9.0, 11.5], h=[5.0, 6.5, 8.0])
    _, _, min_w, min_h, hn, wn = orr._grid_geometry(odd["x"], odd["y"], odd["w"], odd["h"])
    assert train.page_grid_shape(odd) == (hn, wn)


def test_bucketed_feeder_groups_by_exact_shape_without_gpu(monkeypatch):
    """Bucketing is pure host logic; HostBatch only needs pinned memory, which is replaced by plain memory here."""
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    shapes = [(40, 48), (56, 40), (40, 48), (40, 48), (24, 64), (56, 40), (40, 48)]
    wp, lp = zip(*[orr.synth_page(400 + i, h, w, 12) for i, (h, w) in enumerate(shapes)])
    f = train.BucketedPageFeeder(list(wp), list(lp), max_pages=2, seed=5)
    assert f.shapes() == sorted(set(shapes)) and f.n_pages == len(shapes)
    seen = []
    for shape, idx, hw, hl in f.batches:
        assert all(shapes[i] == shape for i in idx) and 1 <= len(idx) <= 2
        assert hw.n_pages == len(idx) == hl.n_pages and hw.with_chars and hl.with_labels
        seen += idx
    assert sorted(seen) == list(range(len(shapes)))
    assert sorted(len(b[1]) for b in f.batches) == [1, 2, 2, 2]        # (40,48) x4 -> 2 + 2, (56,40) x2 -> 2, (24,64) -> 1
    # deterministic, epoch-dependent order
    a = [tuple(b[1]) for b in f]
    b = [tuple(b[1]) for b in f]
    g = train.BucketedPageFeeder(list(wp), list(lp), max_pages=2, seed=5)
    assert [tuple(x[1]) for x in g] == a and sorted(a) == sorted(b)
    with pytest.raises(ValueError):
        train.BucketedPageFeeder(list(wp), list(lp)[:-1])


def test_post_process_kv_matches_reference_table():
    """inference/postprocess.py:8-15: odd class ids > 1 are the value classes, keyed by the class name without its 'v_' prefix."""
    values = [("", [], None, None)] * 17
    values = list(values)
    values[3] = ("Mizuho", [], None, None)
    values[7] = ("1234567", [], None, None)
    values[4] = ("ignored key class", [], None, None)
    got = kv_model.post_process_kv(values)
    # values[idx] with idx odd and > 1 is named by CLASS_NAMES[idx - 1] = 'v_<field>'
    assert got["bank_name"] == "Mizuho" and got["bank_branch_name"] == "" and got["account_number"] == "1234567"
    assert set(got) == {kv_model.CLASS_NAMES[i - 1][2:] for i in range(3, 17, 2)}
    assert "ignored key class" not in got.values()


def test_get_optimizer_matches_reference_choices():
    """model/training/optimizer.py:4-30: rmsprop by default, 'momentum' -> SGD(0.9), anything else -> Adam; lr 0.001."""
    m = msau_b200.MSAUWrapper(12, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    o = training.get_optimizer(m)
    assert o.name == "rmsprop" and o.param_groups[0]["lr"] == 0.001 and o.fused_args()["betas"][0] == 0.99 and o.fused_args()["max_norm"] == 0.0
    o = training.get_optimizer(m, dict(optimizer="momentum", learning_rate=0.01, momentum=0.8, lr_decay_rate=1e-4))
    fa = o.fused_args()
    assert o.name == "momentum" and fa["lr"] == 0.01 and fa["betas"][0] == 0.8 and fa["weight_decay"] == 1e-4
    o = training.get_optimizer(m, dict(optimizer="adam", learning_rate=None))
    assert o.name == "adam" and o.param_groups[0]["lr"] == 0.001
    t = training.Trainer(m)
    t._initialize(None)
    assert [t.adjust_lr(e) for e in (0, 9, 10, 25)] == [0.001, 0.001, 0.001 * 0.95, 0.001 * 0.95 ** 2]     # trainer.py:45-49
    with pytest.raises(NotImplementedError):
        training.UNetLoss({"cost_name": "dice"})


def test_plan_options_are_per_plan_through_the_abi():
    """msau_set_option edits the defaults a NEW plan copies; msau_plan_set_option edits one plan; unknown names are errors.
    (Plan creation does no GPU work beyond two small descriptor uploads, so this only runs where a device exists.)"""
    L = _lib.lib()
    assert L.msau_plan_set_option(None, b"pdl", 0) == -1
    assert L.msau_set_option(b"no_such_option", 1) == -1 and b"unknown option" in L.msau_last_error()
    assert L.msau_set_option(b"pdl", 1) == 0
    if not torch.cuda.is_available():
        return
    cfg = _lib.MsauConfig(12, 17, 3, 2, 8, 3, 2, 3)              # 17 classes: logits pitch 32
    h = C.c_void_p()
    assert L.msau_plan_create(C.byref(cfg), 1, 32, 40, C.byref(h)) == 0
    assert L.msau_plan_set_option(h, b"conv3_tma", 0) == 0 and L.msau_plan_set_option(h, b"nope", 0) == -1
    L.msau_plan_destroy(h)


def test_n_class_limits():
    L = _lib.lib()
    h = C.c_void_p()
    cfg = _lib.MsauConfig(12, 33, 3, 2, 8, 3, 2, 3)
    assert L.msau_plan_create(C.byref(cfg), 1, 32, 40, C.byref(h)) == -1 and b"n_class" in L.msau_last_error()
