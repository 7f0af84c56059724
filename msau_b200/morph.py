"""Rectangular morphology + connected components on the GPU, bit-exact with the reference's SciPy calls.

Mirrors inference/morph_util.py:13-22,65-84 (same names, argument meaning and quirks):

    r_dilation / r_erosion / r_opening / r_closing (image, size, origin=0) -> ndarray, same dtype
    connected_components(image, thres=0) -> (labels int32 ndarray, objects list[(slice, slice)])

plus batched device-resident versions used by the inference driver (``*_batch``).  ``r_closing`` ignores its
``origin`` argument exactly like the reference (morph_util.py:81-84 passes 0).  Images must be bool or
uint8; there is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import _lib


def _pair(v):
    if np.isscalar(v):
        return int(v), int(v)
    a, b = v
    return int(a), int(b)


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MsauError("msau_b200.morph needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def rect_filter_batch(maps: torch.Tensor, size, origin=0, is_max: bool = True) -> torch.Tensor:
    """maps: uint8 CUDA tensor [n, H, W] -> filtered uint8 [n, H, W] (SciPy maximum/minimum_filter, mode='constant')."""
    assert maps.is_cuda and maps.dtype == torch.uint8 and maps.dim() == 3
    maps = maps.contiguous()
    sh, sw = _pair(size)
    oh, ow = _pair(origin)
    out = torch.empty_like(maps)
    n, H, W = maps.shape
    with torch.cuda.device(maps.device):
        _lib.check(_lib.lib().msau_rect_filter(maps.data_ptr(), out.data_ptr(), n, H, W, sh, sw, oh, ow, int(is_max),
                                               _lib.current_stream()))
    return out


def closing_batch(maps: torch.Tensor, size) -> torch.Tensor:
    return rect_filter_batch(rect_filter_batch(maps, size, 0, True), size, 0, False)


def class_closing_batch(class_map: torch.Tensor, cls: int, size) -> torch.Tensor:
    """r_closing(pred_class == cls, size) of kv_model.py:175-176 for a batch of uint8 class maps [n, H, W].  Row windows (1, k <= 4)
    on maps whose width is a multiple of 16 run as ONE kernel that reads the class map once; anything else falls back to
    class_equals + dilation + erosion."""
    assert class_map.is_cuda and class_map.dtype == torch.uint8 and class_map.dim() == 3
    sh, sw = _pair(size)
    n, H, W = class_map.shape
    class_map = class_map.contiguous()
    if sh != 1 or sw > 4 or W % 16 or class_map.data_ptr() % 16:
        return closing_batch(class_equals(class_map, cls), size)
    out = torch.empty_like(class_map)
    with torch.cuda.device(class_map.device):
        _lib.check(_lib.lib().msau_class_closing_row(class_map.data_ptr(), out.data_ptr(), n, H, W, int(cls), sw, _lib.current_stream()))
    return out


def class_equals(class_map: torch.Tensor, cls: int) -> torch.Tensor:
    """(class_map == cls) as uint8 0/1 -- kv_model.py:175."""
    assert class_map.is_cuda and class_map.dtype == torch.uint8
    class_map = class_map.contiguous()
    out = torch.empty_like(class_map)
    with torch.cuda.device(class_map.device):
        _lib.check(_lib.lib().msau_class_equals(class_map.data_ptr(), out.data_ptr(), class_map.numel(), int(cls),
                                                _lib.current_stream()))
    return out


def ccl_batch(binary: torch.Tensor, max_labels: int = 4096):
    """binary uint8 CUDA [n,H,W] -> (labels int32 [n,H,W], n_labels int32 [n], bboxes int32 [n,max_labels,4] = y0,y1,x0,x1)."""
    assert binary.is_cuda and binary.dtype == torch.uint8 and binary.dim() == 3
    binary = binary.contiguous()
    n, H, W = binary.shape
    dev = binary.device
    labels = torch.empty((n, H, W), dtype=torch.int32, device=dev)
    n_labels = torch.empty((n,), dtype=torch.int32, device=dev)
    bboxes = torch.empty((n, max_labels, 4), dtype=torch.int32, device=dev)
    wp = (W + 31) // 32
    scratch = torch.empty((n * H * W + n * H * wp + n * ((H * wp + 31) // 32 + 2),), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().msau_ccl4(binary.data_ptr(), n, H, W, labels.data_ptr(), n_labels.data_ptr(), bboxes.data_ptr(),
                                        max_labels, scratch.data_ptr(), _lib.current_stream()))
    return labels, n_labels, bboxes


# ----------------------------------------------------------------------------- reference-named numpy API
def _to_dev(image) -> Tuple[torch.Tensor, np.dtype]:
    a = np.asarray(image)
    if a.dtype not in (np.bool_, np.uint8):
        raise TypeError(f"msau_b200.morph supports bool/uint8 images, got {a.dtype}")
    if a.ndim != 2:
        raise ValueError("expected a 2-D image")
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8)).to(_dev())
    return t[None], a.dtype


def r_dilation(image, size, origin=0):
    """Dilation with rectangular structuring element using maximum_filter (morph_util.py:65-67)."""
    t, dt = _to_dev(image)
    return rect_filter_batch(t, size, origin, True)[0].cpu().numpy().view(dt)


def r_erosion(image, size, origin=0):
    """Erosion with rectangular structuring element using minimum_filter (morph_util.py:70-72)."""
    t, dt = _to_dev(image)
    return rect_filter_batch(t, size, origin, False)[0].cpu().numpy().view(dt)


def r_opening(image, size, origin=0):
    t, dt = _to_dev(image)
    t = rect_filter_batch(rect_filter_batch(t, size, origin, False), size, origin, True)
    return t[0].cpu().numpy().view(dt)


def r_closing(image, size, origin=0):
    t, dt = _to_dev(image)
    return closing_batch(t, size)[0].cpu().numpy().view(dt)


def objects_from_bboxes(n: int, bboxes: np.ndarray) -> List[Tuple[slice, slice]]:
    return [(slice(int(b[0]), int(b[1])), slice(int(b[2]), int(b[3]))) for b in bboxes[:n]]


def connected_components(image, thres=0):
    """scipy.ndimage.label (4-connectivity) + find_objects (morph_util.py:13-22)."""
    a = np.asarray(image)
    binary = a > thres if thres > 0 else a
    t = torch.from_numpy(np.ascontiguousarray(binary != 0).view(np.uint8)).to(_dev())[None]
    cap = 4096
    while True:
        labels, n_labels, bboxes = ccl_batch(t, cap)
        n = int(n_labels[0])
        if n <= cap:
            break
        cap = n
    return labels[0].cpu().numpy(), objects_from_bboxes(n, bboxes[0].cpu().numpy())
