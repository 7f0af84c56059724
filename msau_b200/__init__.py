"""msau_b200 -- B200-native (sm_100a) engine for the datvo06/MSAU hot path.

Host-side mirror of the reference's interfaces over the C ABI in ``include/msau_b200.h``:

    msau_b200.model      MSAUWrapper                      (model/model.py)
    msau_b200.raster     chargrid / BERT-grid rasterisers (data_generator_funsd_bert.py, inference/kv_model.py)
    msau_b200.morph      r_dilation ... connected_components (inference/morph_util.py)
    msau_b200.kv_model   KVModel inference driver          (inference/kv_model.py)
    msau_b200.train      train / evaluate + data-parallel step (train_chargrid_funsd_msau.py)

There is no CPU fallback anywhere in this package.
"""
from ._lib import MsauError, lib, launch_count  # noqa: F401
from .model import MSAUWrapper, MSAU  # noqa: F401

__all__ = ["MSAUWrapper", "MSAU", "MsauError", "lib", "launch_count"]
