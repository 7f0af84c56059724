"""Reference-shaped drivers on the GPU: KVModel inference path and the train()/evaluate() loop."""
import json
import types

import numpy as np
import pytest
import torch

import msau_b200
from msau_b200 import kv_model, train as mtrain
from oracle import model as om
from oracle import morph as omo
from oracle import raster as orr

pytestmark = pytest.mark.gpu
CHARSET = "".join(chr(c) for c in range(33, 127) if chr(c) != "$") + chr(161)


def make_kv(seed=2):
    cfg = om.MsauConfig(channels=96, n_class=5, scale_space_num=3, res_depth=2, feat_root=8)
    sd = om.init_state_dict(cfg, seed)
    net = msau_b200.MSAUWrapper(96, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    net.load_state_dict(sd)
    kv = kv_model.KVModel()
    kv.net = net
    kv.load(model_weight=None, charset=None, n_class=5)
    kv.set_charset(CHARSET)
    assert kv.n_token == 96
    return kv, cfg, sd


def test_kv_model_inference_path(tmp_path):
    kv, cfg, sd = make_kv()
    words, _ = orr.synth_page(5, 40, 44, 25)
    texts = ["".join(CHARSET[c - 2] for c in ch) for ch in words["chars"]]
    lines = [dict(box=[int(words["x"][i]), int(words["y"][i]), int(words["x"][i] + words["w"][i]), int(words["y"][i] + words["h"][i])],
                  text=texts[i], type=0, value=0) for i in range(len(texts))]
    p = tmp_path / "page.json"
    p.write_text(json.dumps(dict(lines=lines)))
    im, lm, cm, label_lines, scale, bg_pad, bbox = kv._generate_masks_from_label(str(p))
    boxes = np.array([l["box"] for l in lines], np.float64)
    ids = [np.array([kv.tok_to_id.get(c, 1) for c in "".join(ch if not ch.isdigit() else "0" for ch in t)], np.int32) for t in texts]
    want = orr.raster_kv_chargrid(boxes, ids)
    assert np.array_equal(im, want["input_mask"]) and np.array_equal(lm, want["line_id_mask"]) and np.array_equal(cm, want["character_id_mask"])
    assert scale == want["scale"] and bg_pad == want["bg_pad"]
    # network on the one-hot grid: class map vs the oracle model
    ids_dev = torch.from_numpy(im.view(np.int16)).cuda()[None]
    pred = kv.predict_maps(ids_dev)
    x = torch.from_numpy(orr.one_hot_nchw(im, kv.n_token))
    logits, _ = om.msau_forward(sd, cfg, x)
    ref = logits.argmax(1).numpy().astype(np.uint8)
    assert (pred.cpu().numpy() == ref).mean() >= 0.999
    # post-process of the engine's own class map: bit-exact vs the oracle on the same map
    comps = kv.components(pred, 5)
    pm = pred[0].cpu().numpy()
    res = omo.postprocess_page(pm, 5)
    for c in range(2, 5):
        closed, labels, n_lab, bboxes = comps[c]
        assert np.array_equal(closed[0].cpu().numpy().astype(bool), res[c][0])
        assert np.array_equal(labels[0].cpu().numpy(), res[c][1])
        assert np.array_equal(bboxes[0, :int(n_lab[0])].cpu().numpy(), res[c][2])
    kv_results, dbg = kv.predict((str(p), None))
    assert isinstance(kv_results, dict) and dbg is None


@pytest.mark.parametrize("idx", range(6))
def test_extract_value_matches_reference(golden_dir, idx):
    """KVModel._extract_value (device closing / labelling / line and character look-ups + host text assembly) against the
    outputs of the unmodified reference on the same inputs (tests/golden/kv_extract.npz): field texts, boxes and the
    float64 new_pred_mask bit for bit."""
    import hashlib
    from kv_cases import build_case, load_golden, plain
    z, cases, charset = load_golden(golden_dir)
    case = cases[idx]
    lm, cm, label_lines, pm = build_case(case, charset)
    values, new_mask = kv_model.KVModel._extract_value(lm, cm, label_lines, pm, case["n_class"])
    key = case["key"]
    assert [plain(v) for v in values] == json.loads(str(z[key + "::values"]))
    assert new_mask.dtype == np.float64 and tuple(new_mask.shape) == tuple(z[key + "::shape"])
    assert hashlib.sha256(np.ascontiguousarray(new_mask).tobytes()).hexdigest() == str(z[key + "::new_mask_sha"])


@pytest.mark.parametrize("seed,n_class,noise", [(21, 5, 0.01), (22, 7, 0.0), (23, 12, 0.0), (24, 6, 0.02)])
def test_extract_value_matches_oracle(seed, n_class, noise):
    """Other pages / class counts (12 classes reach the reference's second multi-line field, 11) against the numpy oracle."""
    from kv_cases import build_case, plain
    from oracle import kv as okv
    charset = "".join(chr(c) for c in range(33, 127) if chr(c) != "$") + chr(161)
    case = dict(seed=seed, gh=56, gw=44, n_words=36, n_class=n_class, noise=noise, pred_seed=seed + 100)
    lm, cm, lines_a, pm = build_case(case, charset)
    _, _, lines_b, _ = build_case(case, charset)
    want, want_mask = okv.extract_value(lm, cm, lines_a, pm, n_class)
    got, got_mask = kv_model.KVModel._extract_value(lm, cm, lines_b, pm, n_class)
    assert [plain(v) for v in got] == [plain(v) for v in want]
    assert np.array_equal(got_mask, want_mask)
    assert any(v[0] for v in want)


def test_extract_value_degenerate_inputs():
    """No foreground pixel at all, only specks below the area threshold, and a 2-class map (no field classes)."""
    H, W = 40, 56
    lm = np.zeros((H, W), np.uint16); lm[5:9, 4:30] = 1
    cm = np.zeros((H, W), np.uint16); cm[5:9, 4:30] = np.arange(1, 27)[None, :]
    lines = [dict(box=[4, 5, 30, 9], text="abcdefghijklmnopqrstuvwxyz", type=0, value=0)]
    pm = np.zeros((H, W, 5), np.float32); pm[:, :, 0] = 1.0
    values, mask = kv_model.KVModel._extract_value(lm, cm, [dict(l) for l in lines], pm, 5)
    assert values == [("", None, None, None)] * 5 and mask.shape == (H, W, 5) and mask[:, :, 1:].sum() == 0
    pm2 = pm.copy(); pm2[6, 10, 0] = 0.0; pm2[6, 10, 3] = 1.0; pm2[6, 11, 0] = 0.0; pm2[6, 11, 3] = 1.0     # area 2 < 5
    values, mask = kv_model.KVModel._extract_value(lm, cm, [dict(l) for l in lines], pm2, 5)
    assert all(v[0] == "" for v in values) and mask[:, :, 1:].sum() == 0
    pm3 = pm.copy(); pm3[5:9, 4:30, 0] = 0.0; pm3[5:9, 4:30, 3] = 1.0                                        # one clean field
    values, mask = kv_model.KVModel._extract_value(lm, cm, [dict(l) for l in lines], pm3, 5)
    assert values[3][0] == lines[0]["text"] and values[3][1] == [[4, 5, 30, 9]] and mask[:, :, 3].sum() == 4 * 26
    values, mask = kv_model.KVModel._extract_value(lm, cm, [dict(l) for l in lines], pm[:, :, :2], 2)
    assert values == [("", None, None, None)] * 2 and mask.shape == (H, W, 2)


def test_extract_value_many_components():
    """A salt-and-pepper class map (what untrained weights predict) has more components than the default box table holds:
    the labelling is repeated with a larger table and the result still equals the oracle's."""
    from oracle import kv as okv
    rng = np.random.RandomState(3)
    H, W = 96, 120
    cls_map = rng.randint(0, 5, (H, W))
    pm = np.zeros((H, W, 5), np.float32)
    pm[np.arange(H)[:, None], np.arange(W)[None, :], cls_map] = 1.0
    lm = np.zeros((H, W), np.uint16); lm[10:14, 5:100] = 1; lm[30:34, 5:100] = 2
    cm = np.zeros((H, W), np.uint16); cm[10:14, 5:100] = np.arange(1, 96)[None, :]; cm[30:34, 5:100] = np.arange(1, 96)[None, :]
    lines = [dict(box=[5, 10, 100, 14], text="a" * 95, type=0, value=0), dict(box=[5, 30, 100, 34], text="b" * 95, type=0, value=0)]
    want, want_mask = okv.extract_value(lm, cm, [dict(l) for l in lines], pm, 5)
    labels, n_lab, bb, cap = kv_model.KVModel._dev_components(torch.from_numpy(cls_map.astype(np.uint8)).cuda(), 5, 64)
    assert int(n_lab.max()) > 64 and cap >= int(n_lab.max()) and bb.shape[1] == int(n_lab.max())
    got, got_mask = kv_model.KVModel._extract_value(lm, cm, [dict(l) for l in lines], pm, 5)
    assert [tuple(v) for v in got] == [tuple(v) for v in want] and np.array_equal(got_mask, want_mask)


def test_train_and_evaluate_loop():
    cfg = om.MsauConfig(channels=96, n_class=5, scale_space_num=3, res_depth=2, feat_root=8)
    model = msau_b200.MSAUWrapper(96, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    model.load_state_dict(om.init_state_dict(cfg, 4))
    dataset = []
    for s in range(3):
        words, lines = orr.synth_page(20 + s, 32, 40, 16)
        grid, label = orr.raster_word_chargrid(words, lines, np.eye(96))
        dataset.append(dict(mask=torch.Tensor(grid).unsqueeze(0), label=torch.Tensor(label).unsqueeze(0)))
    args = types.SimpleNamespace(num_epochs=2, clip=True, batch_size=1)
    before = model.flat_params.clone()
    model, val_accs = mtrain.train(dataset, model, args, val_dataset=dataset)
    assert len(val_accs) == 2 and all(0.0 <= a <= 1.0 for a in val_accs)
    assert not torch.equal(before.cuda(), model.flat_params)
    r = mtrain.evaluate(dataset, model, args, name="Train", max_num_examples=100)
    assert {"prec", "recall", "acc"} <= set(r)
