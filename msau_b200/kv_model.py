"""Inference driver over the CUDA engine: JSON lines -> char-id grid (R3) -> one-hot -> MSAU -> per-class closing +
connected components -> field values.

Mirrors inference/kv_model.py: ``KVModel.load`` (:37-57), ``_generate_masks_from_label`` (:83-148), ``predict``
(:264-338) and ``_extract_value`` (:151-261) keep their names and argument meaning.  On the hot path the grid
rasterisation, the one-hot expansion, the network, the argmax, the (1,3) closing, the labelling and the line / character
look-ups of ``_extract_value`` all run on the device; only the per-class component summaries (bounding boxes, touched line
ids, character ranges) come back for the text assembly, which stays host-side.
Visualisation (PIL / cv2 debug images) is out of scope: ``predict`` returns ``(kv_results, None)``.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, morph, raster

# inference/postprocess.py:2-5 -- the reference's (project-specific) class table; assign another list to use other field names
CLASS_NAMES = ['NUL', 'k_bank_name', 'v_bank_name', 'k_bank_branch_name', 'v_bank_branch_name', 'k_account_number',
               'v_account_number', 'k_account_type', 'v_account_type', 'k_account_name', 'v_account_name',
               'k_account_name_kana', 'v_account_name_kana', 'k_branch', 'v_branch', 'k_financial_institution',
               'v_financial_institution']


def post_process_kv(values):
    """inference/postprocess.py:8-15: the value classes (odd class ids > 1) keyed by their field name."""
    results = {}
    for idx, v in enumerate(values):
        if idx % 2 == 1 and idx > 1:
            field_name = CLASS_NAMES[idx - 1][2:] if len(CLASS_NAMES) > idx - 1 else str(idx - 1)
            results[field_name] = v[0]
    return results


class KVModel:
    default_config = {"scale": 3.0, "charset": "", "model_kv": "", "n_class": 0}

    def __init__(self):
        self.net = None              # assign an msau_b200.MSAUWrapper before load(), as with the reference (:29,:38)
        self.scale = None
        self.tok_to_id, self.id_to_tok = None, None
        self.blank_idx = 1
        self.n_token = 1
        self.charset = ""
        self.n_class = 1

    def load(self, **config):
        self.net = self.net.cuda()
        if config.get("model_weight"):
            self.net.load_weights(config["model_weight"])
        self.net.eval()
        self.scale = self.default_config["scale"]
        path_charset = config.get("charset")
        if path_charset is not None:
            with open(path_charset, "r") as f:
                self.set_charset(f.read())
        else:
            self.charset = None
        self.n_class = config["n_class"]

    def set_charset(self, chars: str):
        """' ' + '$' + file contents -> token table (kv_model.py:46-53)."""
        self.charset = " " + "$" + chars
        self.blank_idx = 1
        self.tok_to_id = {tok: idx for idx, tok in enumerate(self.charset)}
        self.id_to_tok = {idx: tok for tok, idx in self.tok_to_id.items()}
        self.n_token = len(self.tok_to_id)

    @staticmethod
    def _read_json_layout_ocr(json_path):
        with open(json_path, "r") as f:
            return json.load(f)

    def _encode_lines(self, label_lines):
        boxes = np.array([l["box"] for l in label_lines], np.float64).reshape(-1, 4)
        ids = []
        for l in label_lines:
            text = "".join(c if not c.isdigit() else "0" for c in l["text"])          # kv_model.py:126
            ids.append(np.array([self.tok_to_id.get(c, self.blank_idx) for c in text], np.int32))
        return boxes, ids

    def _generate_masks_from_label(self, label_path, as_numpy: bool = True):
        """-> (input_mask, line_id_mask, character_id_mask, label_lines, scale, bg_pad, bounding_box), kv_model.py:83-148.
        Rasterised on the device; ``as_numpy=False`` keeps the three masks there (int16 storage of the uint16 values)."""
        json_dict = self._read_json_layout_ocr(label_path) if isinstance(label_path, str) else label_path
        label_lines = json_dict["lines"]
        boxes, ids = self._encode_lines(label_lines)
        bbox = (min(l["box"][0] for l in label_lines), min(l["box"][1] for l in label_lines),
                max(l["box"][2] for l in label_lines), max(l["box"][3] for l in label_lines))
        r = raster.rasterize_kv([boxes], [ids])
        g = r["geom3"].cpu().numpy()[0]
        scaled = r["scaled_boxes"].cpu().numpy()
        for l, sb in zip(label_lines, scaled):
            l["box"] = [int(v) for v in sb]                                            # kv_model.py:125
        masks = [r[k][0] for k in ("input_mask", "line_id_mask", "character_id_mask")]
        if as_numpy:
            masks = [m.cpu().numpy().view(np.uint16) for m in masks]
        return masks[0], masks[1], masks[2], label_lines, float(g[2]), int(g[3]), bbox

    # ------------------------------------------------------------------ device hot path
    def predict_maps(self, input_masks: torch.Tensor):
        """uint16 char-id grids [n,H,W] (CUDA) -> uint8 class maps [n,H,W]: one-hot (generic_util.py:94-95) + network + argmax."""
        x = raster.one_hot(input_masks, self.n_token, layout="nhwc")
        return self.net.predict_classes(x, layout=1)

    @staticmethod
    def components(pred_class: torch.Tensor, n_class: int, size=(1, 3), max_labels: int = 4096):
        """kv_model.py:174-177 for every class c in 2..n_class-1 of every page, batched on the device.
        Returns {c: (closed uint8 [n,H,W], labels int32 [n,H,W], n_labels int32 [n], bboxes int32 [n,max_labels,4])}."""
        out = {}
        for c in range(2, n_class):
            closed = morph.class_closing_batch(pred_class, c, size)
            labels, n_labels, bboxes = morph.ccl_batch(closed, max_labels)
            out[c] = (closed, labels, n_labels, bboxes)
        return out

    # ------------------------------------------------------------------ _extract_value (kv_model.py:151-261)
    multiple_lines_fields = (5, 11)          # kv_model.py:156 (the reference's other two lists are empty)

    @staticmethod
    def _reading_order(lines):
        """sort_box_reading_order (inference/generic_util.py:51-91): repeatedly take the top-left-most remaining line."""
        rest, out = list(lines), []
        if not rest:
            return rest
        while len(rest) > 1:
            pick = rest[0]
            for cand in rest[1:]:
                _, py1, px2, py2 = pick["box"]
                x1, y1, x2, y2 = cand["box"]
                cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
                if cy <= (py1 + py2) / 2 - (y2 - y1) / 2 or (cx < px2 and cy < py2):
                    pick = cand
            out.append(pick)
            rest.remove(pick)
        return out + [rest[0]]

    @staticmethod
    def _merge_boxes(boxes, outer: bool):
        """union_boxes / intersect_boxes (inference/morph_util.py:86-104)."""
        if not boxes:
            return None
        lo, hi = (min, max) if outer else (max, min)
        x1, y1, x2, y2 = boxes[0]
        for b in boxes[1:]:
            x1, y1, x2, y2 = lo(x1, b[0]), lo(y1, b[1]), hi(x2, b[2]), hi(y2, b[3])
        return [x1, y1, x2, y2]

    # The three device steps of ``extract_value_device``; tests/test_kv_host.py swaps in numpy stand-ins to check the host logic
    # (component choice, reading order, text assembly) without a GPU.
    @staticmethod
    def _dev_components(pred_class: torch.Tensor, n_class: int, max_labels: int):
        """closing + labelling of every foreground class map in one batch (kv_model.py:174-177) ->
        (labels int32 [n_maps,H,W] on the device, component counts [n_maps] and bounding boxes [n_maps, n, 4] on the host, the
        label capacity actually used: the labelling is repeated with a larger box table when a map has more components)."""
        maps = torch.stack([morph.class_equals(pred_class[None], c)[0] for c in range(2, n_class)])
        closed = morph.closing_batch(maps, (1, 3))
        labels, n_lab, bboxes = morph.ccl_batch(closed, max_labels)
        n_lab_h = n_lab.cpu().numpy()
        if int(n_lab_h.max()) > max_labels:          # a noisy map (untrained weights): boxes of every component are needed
            max_labels = 1 << (int(n_lab_h.max()) - 1).bit_length()
            labels, n_lab, bboxes = morph.ccl_batch(closed, max_labels)
        return labels, n_lab_h, bboxes[:, :max(int(n_lab_h.max()), 1)].cpu().numpy(), max_labels

    @staticmethod
    def _dev_select(labels, line_mask, slot_of: np.ndarray, n_slots: int, num_lines: int):
        """-> (presence uint8 [n_slots, num_lines+1] on the host, new_mask uint8 [n_maps,H,W] on the device)."""
        from . import _lib
        dev = labels.device
        n_maps, H, W = labels.shape
        slot_d = torch.from_numpy(slot_of).to(dev)
        presence = torch.empty((n_slots, num_lines + 1), dtype=torch.uint8, device=dev)
        new_mask = torch.empty((n_maps, H, W), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().msau_kv_select_components(labels.data_ptr(), line_mask.data_ptr(), n_maps, H, W, slot_d.data_ptr(),
                                                            slot_of.shape[1] - 1, n_slots, num_lines, presence.data_ptr(),
                                                            new_mask.data_ptr(), _lib.current_stream()))
        return presence.cpu().numpy(), new_mask

    @staticmethod
    def _dev_char_ranges(char_mask, new_mask, queries):
        """queries [(map, x1, y1, x2, y2)] -> int array [nq, 2] = (min, max) character index, (INT_MAX, 0) if none."""
        from . import _lib
        dev = new_mask.device
        _, H, W = new_mask.shape
        q_d = torch.tensor(queries, dtype=torch.int32, device=dev)
        r_d = torch.empty((len(queries), 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().msau_kv_char_range(char_mask.data_ptr(), new_mask.data_ptr(), H, W, q_d.data_ptr(), len(queries),
                                                     r_d.data_ptr(), _lib.current_stream()))
        return r_d.cpu().numpy()

    @classmethod
    def extract_value_device(cls, line_mask, char_mask, label_lines, pred_class, n_class: int, num_classes: int,
                             max_labels: int = 4096):
        """The reference's ``_extract_value`` with every full-size map kept on the device.  ``line_mask`` / ``char_mask``: int16
        [H,W] CUDA (uint16 values, as ``_generate_masks_from_label(as_numpy=False)`` returns them), ``pred_class``: uint8 [H,W]
        CUDA arg-max map.  Returns (values, new_mask) with ``new_mask`` uint8 [n_class-2, H, W] on the device = the planes
        2.. of the reference's ``new_pred_mask`` (None when no component was picked: all zeros).  Host<->device traffic: the
        component bounding boxes (16 B each), one slot table per class, one byte per (picked component, line) and 8 B per
        doubly-claimed line."""
        H, W = pred_class.shape
        n_maps = n_class - 2
        num_lines = len(label_lines)
        values = [("", None, None, None)] * n_class
        used = [0] * (num_lines + 1)
        line_ids_for_field = [[] for _ in range(num_classes + 1)]
        boxes_for_field = [[] for _ in range(num_classes + 1)]
        for i, l in enumerate(label_lines):
            l["id"] = i + 1
        if n_maps <= 0:
            return values, None
        labels, n_lab_h, bb_h, max_labels = cls._dev_components(pred_class, n_class, max_labels)
        # ---- pick components per class from the bounding boxes (kv_model.py:181-205)
        slot_of = np.full((n_maps, max_labels + 1), -1, np.int32)
        picked = {}                                       # c -> [component ids, best first then alternatives]
        n_slots = 0
        for c in range(2, n_class):
            n = int(n_lab_h[c - 2])
            if n == 0:
                continue
            bb = bb_h[c - 2, :n].astype(np.int64)             # y0, y1, x0, x1 (half-open), one row per component
            area = (bb[:, 3] - bb[:, 2]) * (bb[:, 1] - bb[:, 0])           # morph_util.area, same integers as the reference's list
            multi = c in cls.multiple_lines_fields
            # np.argsort on the same values in the same order as the reference's Python lists (ties break identically);
            # ycenter = np.mean([stop, start]) = (y0 + y1) / 2 exactly
            order = np.argsort(-((bb[:, 1] + bb[:, 0]) / 2.0)) if multi else np.argsort(area)
            best = int(order[-1])
            if area[best] < 5:
                continue
            alts = [int(k) for k in order[:-1] if area[int(k)] > 5] if (multi and n > 1) else []
            for k in alts + [best]:
                y0, y1, x0, x1 = (int(v) for v in bb[k])
                boxes_for_field[c].append([x0, y0, x1, y1])
            picked[c] = [best] + alts
            for k in picked[c]:
                slot_of[c - 2, k + 1] = n_slots
                n_slots += 1
        new_mask = None
        if n_slots:
            pres_h, new_mask = cls._dev_select(labels, line_mask, slot_of, n_slots, num_lines)
            for c, comps in picked.items():
                line_ids = []
                for k in comps:                           # np.unique order = ascending ids; 0 (no line) dropped (:208, :212)
                    line_ids += [int(i) for i in np.nonzero(pres_h[slot_of[c - 2, k + 1]])[0] if i > 0]
                line_ids_for_field[c] = list(set(line_ids))
                for i in line_ids:
                    used[i] += 1
        # ---- character ranges of the lines claimed by more than one field (kv_model.py:236-241)
        ordered, queries = {}, []
        for c in range(2, n_class):
            if not line_ids_for_field[c]:
                continue
            ordered[c] = cls._reading_order([label_lines[i - 1] for i in line_ids_for_field[c] if i > 0])
            for line in ordered[c]:
                if used[line["id"]] > 1:
                    x1, y1, x2, y2 = line["box"]
                    ys, ye, _ = slice(y1, y2).indices(H)       # numpy slice semantics (clipping, negative wrap)
                    xs, xe, _ = slice(x1, x2).indices(W)
                    queries.append((c - 2, xs, ys, xe, ye))
        ranges = cls._dev_char_ranges(char_mask, new_mask, queries) if queries else []
        # ---- text assembly (kv_model.py:223-255)
        qi = iter(ranges)
        for c in range(2, n_class):
            if c not in ordered:
                continue
            value, line_boxes = "", []
            for line in ordered[c]:
                line_boxes.append(line["box"])
                if used[line["id"]] <= 1:
                    value += line["text"]
                else:
                    lo, hi = (int(v) for v in next(qi))
                    if hi == 0:                           # no character of this line under the field's mask
                        continue
                    if hi > len(line["text"]) - 3:
                        hi = len(line["text"]) + 1
                    value += line["text"][lo - 2 if lo >= 2 else 0: hi - 1]
                if c in cls.multiple_lines_fields:
                    value += "\n"
            if value and value[-1] == "\n":
                value = value[:-1]
            merged = cls._merge_boxes(line_boxes, True)
            values[c] = (value, [boxes_for_field[c][-1]], cls._merge_boxes(boxes_for_field[c] + [merged], False),
                         cls._merge_boxes(boxes_for_field[c] + [merged], True))
        return values, new_mask

    @staticmethod
    def _extract_value(line_mask, char_mask, label_lines, pred_mask, num_classes):
        """kv_model.py:151-261, same arguments and return value: ``pred_mask`` [H,W,C] probabilities (numpy or CUDA tensor),
        returns ``(values, new_pred_mask)`` with ``values[c] = (text, [box], intersect_box, union_box)`` and ``new_pred_mask``
        a float64 [H,W,C] numpy array.  Everything between the arg-max and the text assembly runs on the device
        (``extract_value_device``); only this compatibility wrapper brings the mask planes back to build the reference's
        return value -- ``predict`` does not."""
        pm = pred_mask if torch.is_tensor(pred_mask) else torch.from_numpy(np.ascontiguousarray(pred_mask))
        pm = pm.cuda()
        n_class = pm.shape[2]

        def dev16(m):
            if torch.is_tensor(m):
                return m.cuda().contiguous()
            return torch.from_numpy(np.ascontiguousarray(np.asarray(m).astype(np.uint16)).view(np.int16)).cuda()

        # np.argmax(pred_mask, axis=-1) (:162, first maximum wins) with the engine's arg-max kernel on the channels-last map
        pm = pm.to(torch.float32).contiguous()
        pred_class = torch.empty(tuple(pm.shape[:2]), dtype=torch.uint8, device=pm.device)
        with torch.cuda.device(pm.device):
            _lib.check(_lib.lib().msau_onehot_argmax(pm.data_ptr(), 2, 1, n_class, pm.shape[0] * pm.shape[1], 1, pred_class.data_ptr(),
                                                     _lib.current_stream()))
        values, new_mask = KVModel.extract_value_device(dev16(line_mask), dev16(char_mask), label_lines, pred_class, n_class,
                                                        num_classes)
        new_pred_mask = np.zeros(tuple(pm.shape))
        new_pred_mask[:, :, 0] = pm[:, :, 0].cpu().numpy()
        if new_mask is not None:
            new_pred_mask[:, :, 2:] = new_mask.permute(1, 2, 0).cpu().numpy()
        return values, new_pred_mask

    def predict(self, data, debug_info=None, label_path=None, eval_results=None):
        """kv_model.py:264-338 -> (kv_results, debug_im).  ``data`` = (json_path, image); the image is only used by the
        reference for its debug rendering and is ignored here (debug_im is None).  The network's arg-max map, the closing,
        the labelling and the line / character look-ups stay on the device; a few hundred bytes per field come back."""
        json_path, _ = data
        input_im, line_mask, char_mask, label_lines, scale, bg_pad, bbox = self._generate_masks_from_label(json_path, as_numpy=False)
        with torch.no_grad():
            pred_class = self.predict_maps(input_im[None])[0]
        values, _ = self.extract_value_device(line_mask, char_mask, label_lines, pred_class, self.n_class, self.n_class)
        kv_results = post_process_kv(values)                  # kv_model.py:315
        return kv_results, None
