"""Chargrid / BERT-grid box rasterisation on the GPU, bit-exact with the reference's NumPy loops.

Mirrors data_generator_funsd_bert.py (R1 ``get_box_mask_box_label_word`` :149-186, R2
``get_box_mask_box_label`` :64-93, ``FUNSDMaskDataLoader.getitem`` :216-222) and the rasteriser half of
inference/kv_model.py (R3 ``_generate_masks_from_label`` :83-148, one-hot :274-278).

Pages travel to the device as CSR arrays (a few KB per page) and are expanded there; the dense grid
(100 MB / page at 96 x 512 x 512 fp32) never crosses PCIe.  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MsauError("msau_b200.raster needs a CUDA device (no CPU fallback)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())




def _h2d(a: np.ndarray, device) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.pin_memory().to(device, non_blocking=True)


class HostBatch:
    """Page records packed once into ONE pinned host buffer (CSR arrays back to back): what a data loader hands to the
    training loop.  ``BoxBatch.from_host`` uploads it with a single asynchronous copy."""

    def __init__(self, pages: Sequence[Dict], with_chars: bool = True, with_labels: bool = False):
        self.n_pages = len(pages)
        self.with_chars, self.with_labels = with_chars, with_labels
        ptr = np.zeros(len(pages) + 1, np.int32)
        ptr[1:] = np.cumsum([len(pg["x"]) for pg in pages])
        self.n_boxes = int(ptr[-1])
        arrays = [np.concatenate([np.asarray(pg[k], np.float64) for pg in pages]) for k in "xywh"] + [ptr]
        if with_chars:
            lens = np.fromiter((len(ch) for pg in pages for ch in pg["chars"]), np.int32, self.n_boxes)
            cptr = np.zeros(self.n_boxes + 1, np.int32)
            cptr[1:] = np.cumsum(lens)
            allc = [np.asarray(ch, np.int32) for pg in pages for ch in pg["chars"] if len(ch)]
            cfeat = np.concatenate(allc) if allc else np.zeros(1, np.int32)
            arrays += [lens, cptr, cfeat]
        if with_labels:
            arrays.append(np.concatenate([np.asarray(pg["label"], np.int32) for pg in pages]))
        self.layout, total = [], 0
        for a in arrays:
            self.layout.append((total, a.nbytes, str(a.dtype), a.shape))
            total += (a.nbytes + 15) // 16 * 16
        self.nbytes = max(total, 16)
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
        hv = self.buf.numpy()
        for a, (o, nb, _, _) in zip(arrays, self.layout):
            hv[o:o + nb] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)


class BoxBatch:
    """CSR batch of pages of boxes on the device.  ``chars`` (list of int arrays per box) is optional."""

    def __init__(self, pages: Sequence[Dict], device=None, with_chars: bool = True, with_labels: bool = False):
        self._bind(HostBatch(pages, with_chars, with_labels), _dev(device))

    @classmethod
    def from_host(cls, host: HostBatch, device=None) -> "BoxBatch":
        self = cls.__new__(cls)
        self._bind(host, _dev(device))
        return self

    def _bind(self, host: HostBatch, dev: torch.device):
        self.n_pages, self.n_boxes, self.h_bytes = host.n_pages, host.n_boxes, host.nbytes
        dev_buf = torch.empty(host.nbytes, dtype=torch.uint8, device=dev)
        dev_buf.copy_(host.buf, non_blocking=True)          # one pinned -> device copy
        self._keep = host                                   # the pinned buffer must outlive the asynchronous copy
        dv = [dev_buf[o:o + nb].view(getattr(torch, dt)).reshape(shape) for o, nb, dt, shape in host.layout]
        self.x, self.y, self.w, self.h, self.page_ptr = dv[:5]
        k = 5
        self.n_chars = self.char_ptr = self.char_feat = self.labels = None
        if host.with_chars:
            self.n_chars, self.char_ptr, self.char_feat = dv[5:8]
            k = 8
        if host.with_labels:
            self.labels = dv[k]
        self.device = dev

    def geometry(self) -> torch.Tensor:
        """[n_pages, 8] fp64 = min_x, min_y, min_w, min_h, min_scale, Hn, Wn, 0 (dgfb.py:49-61,154-160)."""
        geom = torch.empty((self.n_pages, 8), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().msau_raster_geometry(self.x.data_ptr(), self.y.data_ptr(), self.w.data_ptr(), self.h.data_ptr(),
                                                       _lib.ptr(self.n_chars), self.page_ptr.data_ptr(), self.n_pages,
                                                       geom.data_ptr(), _lib.current_stream()))
        return geom


def raster_features(boxes: BoxBatch, geom: torch.Tensor, feat_table: torch.Tensor, out_hw: Tuple[int, int], use_chars: bool,
                    layout: str = "nchw", feat_row: Optional[torch.Tensor] = None, owner: Optional[torch.Tensor] = None) -> torch.Tensor:
    """R1 (use_chars) / R2 feature grid.  feat_table: fp64 CUDA [rows, D].  Returns fp32 [n,D,H,W] ("nchw") or [n,H,W,Dp]
    ("nhwc"), or -- layout "ids" -- the int16 [n,H,W] map of feature-table ROW indices (-1 = background): for a one-hot
    table with row r = e_r (``np.eye``, the chargrid) that is the channel id map ``MSAUWrapper.train_step(..., layout=2)``
    consumes without the dense grid ever being written."""
    H, W = out_hw
    D = feat_table.shape[1]
    dev = boxes.device
    n = boxes.n_pages
    lay = {"nchw": 0, "nhwc": 1, "ids": 2}[layout]
    if lay == 2:
        grid = torch.empty((n, H, W), dtype=torch.int16, device=dev)
    else:
        grid = torch.empty((n, D, H, W) if lay == 0 else (n, H, W, (D + 3) // 4 * 4), dtype=torch.float32, device=dev)
    if owner is None:
        owner = torch.empty((n * H * W,), dtype=torch.int32, device=dev)
    if not use_chars and feat_row is None:
        feat_row = torch.arange(boxes.n_boxes, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().msau_raster_features(
            boxes.x.data_ptr(), boxes.y.data_ptr(), boxes.w.data_ptr(), boxes.h.data_ptr(), boxes.page_ptr.data_ptr(), n,
            boxes.n_boxes, _lib.ptr(boxes.char_ptr) if use_chars else 0, _lib.ptr(boxes.char_feat) if use_chars else 0,
            0 if use_chars else feat_row.data_ptr(), feat_table.data_ptr(), D, geom.data_ptr(), int(use_chars), H, W, lay,
            grid.data_ptr(), owner.data_ptr(), _lib.current_stream()))
    return grid


def raster_labels(boxes: BoxBatch, geom: torch.Tensor, out_hw: Tuple[int, int], owner: Optional[torch.Tensor] = None) -> torch.Tensor:
    """label_mask[ny:ny+nh, nx:nx+nw] = label + 1 (uint8), dgfb.py:88-89 / :176-182."""
    H, W = out_hw
    dev = boxes.device
    n = boxes.n_pages
    out = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    if owner is None:
        owner = torch.empty((n * H * W,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().msau_raster_labels(boxes.x.data_ptr(), boxes.y.data_ptr(), boxes.w.data_ptr(), boxes.h.data_ptr(),
                                                 boxes.labels.data_ptr(), boxes.page_ptr.data_ptr(), n, boxes.n_boxes,
                                                 geom.data_ptr(), H, W, out.data_ptr(), owner.data_ptr(), _lib.current_stream()))
    return out


def _grid_hw(geom: torch.Tensor) -> Tuple[int, int]:
    g = geom.cpu().numpy()   # synchronises: only used when the caller did not fix the page size
    return int(g[:, 5].max()), int(g[:, 6].max())


def rasterize_word_chargrid(word_pages: Sequence[Dict], line_pages: Sequence[Dict], feat_table, out_hw=None, layout="nchw",
                            device=None):
    """Batched R1: returns (grid fp32, label uint8 [n,H,W], geom fp64 [n,8]).  ``feat_table`` [rows, D] float64."""
    dev = _dev(device)
    words = BoxBatch(word_pages, dev, with_chars=True)
    lines = BoxBatch(line_pages, dev, with_chars=False, with_labels=True)
    geom = words.geometry()
    if out_hw is None:
        out_hw = _grid_hw(geom)
    table = feat_table if torch.is_tensor(feat_table) else _h2d(np.asarray(feat_table, np.float64), dev)
    owner = torch.empty((len(word_pages) * out_hw[0] * out_hw[1],), dtype=torch.int32, device=dev)
    grid = raster_features(words, geom, table, out_hw, True, layout, owner=owner)
    label = raster_labels(lines, geom, out_hw, owner=owner)
    return grid, label, geom


def rasterize_box_grid(cell_pages: Sequence[Dict], feats: Sequence[np.ndarray], out_hw=None, layout="nchw", device=None):
    """Batched R2: ``feats[p]`` is [n_cells_p, D] float64.  Returns (grid, label, geom)."""
    dev = _dev(device)
    cells = BoxBatch(cell_pages, dev, with_chars=False, with_labels=True)
    geom = cells.geometry()
    if out_hw is None:
        out_hw = _grid_hw(geom)
    table = _h2d(np.concatenate([np.asarray(f, np.float64) for f in feats]), dev)
    owner = torch.empty((len(cell_pages) * out_hw[0] * out_hw[1],), dtype=torch.int32, device=dev)
    grid = raster_features(cells, geom, table, out_hw, False, layout, owner=owner)
    label = raster_labels(cells, geom, out_hw, owner=owner)
    return grid, label, geom


# ----------------------------------------------------------------------------- R3 (inference chargrid)
def rasterize_kv(box_pages: Sequence[np.ndarray], char_id_pages: Sequence[Sequence[np.ndarray]], out_hw=None, device=None):
    """Batched R3.  box_pages[p]: [n_lines,4] x1,y1,x2,y2; char_id_pages[p][i]: token ids of line i.
    Returns dict(input_mask, line_id_mask, character_id_mask uint16 [n,H,W], scaled_boxes int32 [n_lines,4], geom3 [n,8])."""
    dev = _dev(device)
    n = len(box_pages)
    ptr, cptr, ids = [0], [0], []
    for bp, cp in zip(box_pages, char_id_pages):
        ptr.append(ptr[-1] + len(bp))
        for c in cp:
            cptr.append(cptr[-1] + len(c)); ids.append(np.asarray(c, np.int32))
    boxes = _h2d(np.concatenate([np.asarray(b, np.float64).reshape(-1, 4) for b in box_pages]), dev)
    page_ptr = _h2d(np.asarray(ptr, np.int32), dev)
    char_ptr = _h2d(np.asarray(cptr, np.int32), dev)
    char_ids = _h2d(np.concatenate(ids) if ids else np.zeros(0, np.int32), dev)
    if char_ids.numel() == 0:
        char_ids = torch.zeros(1, dtype=torch.int32, device=dev)
    geom3 = torch.empty((n, 8), dtype=torch.float64, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        _lib.check(L.msau_raster_kv_geometry(boxes.data_ptr(), page_ptr.data_ptr(), n, geom3.data_ptr(), _lib.current_stream()))
        if out_hw is None:
            g = geom3.cpu().numpy()
            out_hw = (int(g[:, 4].max()), int(g[:, 5].max()))
        H, W = out_hw
        masks = torch.empty((3, n, H, W), dtype=torch.int16, device=dev)
        scaled = torch.empty((ptr[-1], 4), dtype=torch.int32, device=dev)
        owner = torch.empty((2 * n * H * W,), dtype=torch.int32, device=dev)
        _lib.check(L.msau_raster_kv(boxes.data_ptr(), page_ptr.data_ptr(), n, ptr[-1], char_ptr.data_ptr(), char_ids.data_ptr(),
                                    geom3.data_ptr(), H, W, masks[0].data_ptr(), masks[1].data_ptr(), masks[2].data_ptr(),
                                    scaled.data_ptr(), owner.data_ptr(), _lib.current_stream()))
    return dict(input_mask=masks[0], line_id_mask=masks[1], character_id_mask=masks[2], scaled_boxes=scaled, geom3=geom3,
                page_ptr=ptr)


def one_hot(ids: torch.Tensor, n_token: int, layout: str = "nchw") -> torch.Tensor:
    """to_categorical + transposes (generic_util.py:94-95, kv_model.py:274-278): uint16 ids [n,H,W] -> fp32 one-hot."""
    assert ids.is_cuda and ids.dtype in (torch.int16, torch.uint16)
    ids = ids.contiguous()
    n, H, W = ids.shape
    lay = 0 if layout == "nchw" else 1
    out = torch.empty((n, n_token, H, W) if lay == 0 else (n, H, W, (n_token + 3) // 4 * 4), dtype=torch.float32, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.check(_lib.lib().msau_one_hot(ids.data_ptr(), n, H, W, n_token, lay, out.data_ptr(), _lib.current_stream()))
    return out


# ----------------------------------------------------------------------------- reference-shaped per-page API
def _cells_to_page(cells, with_chars_from=None) -> Dict:
    pg = dict(x=[c.x for c in cells], y=[c.y for c in cells], w=[c.w for c in cells], h=[c.h for c in cells])
    return pg


def _feature_table(char_feats: List[np.ndarray]):
    """charset_feature[word][j] rows -> (table [rows, D] float64, per-word row-index arrays).  One-hot rows (the
    chargrid case, funsd_preprocessing_word_level.py:50-57) collapse to an identity table."""
    flat = [np.asarray(v, np.float64).reshape(len(v), -1) if len(v) else np.zeros((0, 0)) for v in char_feats]
    D = max((f.shape[1] for f in flat if f.size), default=0)
    stack = np.concatenate([f for f in flat if f.size]) if D else np.zeros((0, 0))
    one_hot_rows = D > 0 and bool(np.all((stack == 0) | (stack == 1)) and np.all(stack.sum(1) == 1))
    idx, k = [], 0
    for f in flat:
        n = f.shape[0] if f.size else 0
        idx.append(stack[k:k + n].argmax(1).astype(np.int32) if one_hot_rows else np.arange(k, k + n, dtype=np.int32))
        k += n
    return (np.eye(D) if one_hot_rows else stack), idx


def get_box_mask_box_label_word(dataset_instance, idx, device=None):
    """R1 with the reference's signature (dgfb.py:149-186): returns {"ocr_values", "mask" fp32 [D,H,W] CUDA,
    "label" uint8 [H,W] CUDA}."""
    inst = dataset_instance.inp_list[idx]
    cw, cl = inst["cells_word"], inst["cells"]
    table, rows = _feature_table(inst["charset_feature"])
    # the reference iterates ``for j, char in enumerate(cell.ocr_value)``: one rectangle per character of the text
    words = _cells_to_page(cw)
    words["chars"] = [rows[i][:len(c.ocr_value)] for i, c in enumerate(cw)]
    lines = _cells_to_page(cl)
    lines["label"] = np.asarray(inst["labels"], np.int32)
    grid, label, _ = rasterize_word_chargrid([words], [lines], table, device=device)
    return {"ocr_values": [c.ocr_value for c in cw], "mask": grid[0], "label": label[0]}


def get_box_mask_box_label(dataset_instance, idx, device=None):
    """R2 with the reference's signature (dgfb.py:64-93)."""
    inst = dataset_instance.inp_list[idx]
    box = dataset_instance.getitem_box(dataset_instance, idx)
    cells = _cells_to_page(inst["cells"])
    cells["label"] = np.asarray(box["label"], np.int32)
    grid, label, _ = rasterize_box_grid([cells], [np.asarray(box["feats"], np.float64)], device=device)
    return {"ocr_values": box["ocr_values"], "mask": grid[0], "label": label[0]}
