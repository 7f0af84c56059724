"""Inference driver over the CUDA engine: JSON lines -> char-id grid (R3) -> one-hot -> MSAU -> per-class closing +
connected components -> field values.

Mirrors inference/kv_model.py: ``KVModel.load`` (:37-57), ``_generate_masks_from_label`` (:83-148), ``predict``
(:264-338) and ``_extract_value`` (:151-261) keep their names and argument meaning.  On the hot path the grid
rasterisation, the one-hot expansion, the network, the argmax, the (1,3) closing and the labelling all run on the device;
only the small per-class component summaries come back for the text assembly, which stays host-side.
Visualisation (PIL / cv2 debug images) is out of scope: ``predict`` returns ``(kv_results, None)``.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import morph, raster

CLASS_NAMES = None   # inference/postprocess.py:2-5 is a project-specific table; callers pass their own


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
            closed = morph.closing_batch(morph.class_equals(pred_class, c), size)
            labels, n_labels, bboxes = morph.ccl_batch(closed, max_labels)
            out[c] = (closed, labels, n_labels, bboxes)
        return out

    @staticmethod
    def _extract_value(line_mask, char_mask, label_lines, pred_mask, num_classes):
        """kv_model.py:151-261 with the argmax / closing / labelling on the device.  ``pred_mask`` is [H,W,C] probabilities
        (numpy or CUDA tensor).  Returns (values, new_pred_mask) like the reference."""
        pm = pred_mask if torch.is_tensor(pred_mask) else torch.from_numpy(np.asarray(pred_mask))
        pm = pm.cuda()
        n_class = pm.shape[2]
        pred_class = pm.argmax(dim=-1).to(torch.uint8)[None].contiguous()
        comps = KVModel.components(pred_class, n_class)
        line_mask = np.asarray(line_mask)
        char_mask = np.asarray(char_mask)
        num_lines = len(label_lines)
        values = [("", None, None, None)] * n_class
        new_pred_mask = np.zeros(tuple(pm.shape))
        new_pred_mask[:, :, 0] = pm[:, :, 0].cpu().numpy()
        line_used_count = [0] * (num_lines + 1)
        line_ids_for_field = [[] for _ in range(num_classes + 1)]
        boxes_for_field = [[] for _ in range(num_classes + 1)]
        for idx, l in enumerate(label_lines):
            l["id"] = idx + 1
        for c in range(2, n_class):
            _, labels_d, n_lab, bboxes = comps[c]
            n = int(n_lab[0])
            if n == 0:
                continue
            objects = morph.objects_from_bboxes(n, bboxes[0].cpu().numpy())
            areas = [(o[1].stop - o[1].start) * (o[0].stop - o[0].start) for o in objects]
            best = int(np.argsort(areas)[-1])
            if areas[best] < 5:
                continue
            labels = labels_d[0].cpu().numpy()
            box = objects[best]
            boxes_for_field[c].append([box[1].start, box[0].start, box[1].stop, box[0].stop])
            line_ids = [int(i) for i in np.unique(line_mask[labels == best + 1]) if i > 0]
            line_ids_for_field[c] = list(set(line_ids))
            for i in line_ids:
                line_used_count[i] += 1
            new_pred_mask[:, :, c][labels == best + 1] = 1
        for c in range(2, n_class):
            line_ids = line_ids_for_field[c]
            if not line_ids:
                continue
            lines = sorted((label_lines[i - 1] for i in line_ids), key=lambda l: (l["box"][1], l["box"][0]))
            value, line_boxes = "", []
            for line in lines:
                line_boxes.append(line["box"])
                if line_used_count[line["id"]] <= 1:
                    value += line["text"]
                    continue
                x1, y1, x2, y2 = line["box"]
                sel = set(np.unique(char_mask[y1:y2, x1:x2][new_pred_mask[:, :, c][y1:y2, x1:x2] > 0])) - {0}
                if not sel:
                    continue
                lo, hi = min(sel), max(sel)
                if hi > len(line["text"]) - 3:
                    hi = len(line["text"]) + 1
                value += line["text"][lo - 2 if lo >= 2 else 0: hi - 1]
            xs1, ys1 = min(b[0] for b in line_boxes), min(b[1] for b in line_boxes)
            xs2, ys2 = max(b[2] for b in line_boxes), max(b[3] for b in line_boxes)
            values[c] = (value, [boxes_for_field[c][-1]], None, [xs1, ys1, xs2, ys2])
        return values, new_pred_mask

    def predict(self, data, debug_info=None, label_path=None, eval_results=None):
        """kv_model.py:264-338 -> (kv_results, debug_im).  ``data`` = (json_path, image); the image is only used by the
        reference for its debug rendering and is ignored here (debug_im is None)."""
        json_path, _ = data
        input_im, line_mask, char_mask, label_lines, scale, bg_pad, bbox = self._generate_masks_from_label(json_path, as_numpy=False)
        x = raster.one_hot(input_im[None], self.n_token, layout="nhwc")
        with torch.no_grad():
            probs = self.net._run_forward(x, 1, False, True)[4]                   # softmax probabilities [1,C,H,W]
        a_pred = probs[0].permute(1, 2, 0)                                           # [H,W,C]  (kv_model.py:307)
        values, _ = self._extract_value(line_mask.cpu().numpy().view(np.uint16), char_mask.cpu().numpy().view(np.uint16),
                                        label_lines, a_pred, self.n_class)
        kv_results = {i: v[0] for i, v in enumerate(values) if v[0]}
        return kv_results, None
