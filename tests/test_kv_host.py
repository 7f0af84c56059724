"""Host logic of msau_b200.kv_model.KVModel.extract_value_device (component choice, line bookkeeping, reading order, text
assembly -- inference/kv_model.py:151-261) checked WITHOUT a GPU: the three device steps are replaced by numpy stand-ins built
on the oracle's morphology, and the result must equal the reference's own output (tests/golden/kv_extract.npz).  The device
steps themselves are checked on the GPU in tests/test_drivers_gpu.py."""
import json

import numpy as np
import pytest

from msau_b200.kv_model import KVModel
from oracle import morph as omo

from kv_cases import build_case, load_golden, plain


class HostKV(KVModel):
    @staticmethod
    def _dev_components(pred_class, n_class, max_labels):
        res = omo.postprocess_page(pred_class, n_class)
        labels = np.stack([res[c][1] for c in range(2, n_class)])
        n_lab = np.array([len(res[c][2]) for c in range(2, n_class)], np.int32)
        bb = np.zeros((n_class - 2, max(int(n_lab.max()), 1), 4), np.int32)
        for c in range(2, n_class):
            bb[c - 2, :n_lab[c - 2]] = res[c][2]
        return labels, n_lab, bb, max_labels

    @staticmethod
    def _dev_select(labels, line_mask, slot_of, n_slots, num_lines):
        slots = np.take_along_axis(slot_of, labels.reshape(labels.shape[0], -1), 1).reshape(labels.shape)
        slots[labels == 0] = -1
        presence = np.zeros((n_slots, num_lines + 1), np.uint8)
        for m in range(labels.shape[0]):
            sel = slots[m] >= 0
            presence[slots[m][sel], line_mask[sel]] = 1
        return presence, (slots >= 0).astype(np.uint8)

    @staticmethod
    def _dev_char_ranges(char_mask, new_mask, queries):
        out = []
        for m, x1, y1, x2, y2 in queries:
            v = char_mask[y1:y2, x1:x2][(new_mask[m, y1:y2, x1:x2] > 0) & (char_mask[y1:y2, x1:x2] > 0)]
            out.append((int(v.min()), int(v.max())) if v.size else (2 ** 31 - 1, 0))
        return np.array(out, np.int64).reshape(-1, 2)


@pytest.mark.parametrize("idx", range(6))
def test_host_logic_matches_reference(golden_dir, idx):
    z, cases, charset = load_golden(golden_dir)
    case = cases[idx]
    lm, cm, label_lines, pm = build_case(case, charset)
    pred_class = np.argmax(pm, axis=-1).astype(np.uint8)
    values, new_mask = HostKV.extract_value_device(lm, cm, label_lines, pred_class, case["n_class"], case["n_class"])
    key = case["key"]
    assert [plain(v) for v in values] == json.loads(str(z[key + "::values"]))
    H, W, C = (int(v) for v in z[key + "::shape"])
    want_fg = np.unpackbits(z[key + "::new_mask_fg"])[:H * W * (C - 1)].reshape(H, W, C - 1)
    got = np.zeros((H, W, C - 1), np.uint8)
    if new_mask is not None:
        got[:, :, 1:] = new_mask.transpose(1, 2, 0)
    assert np.array_equal(got, want_fg)
