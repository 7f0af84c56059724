"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol the header declares;
host-side logic (schema, state_dict round trip, loud failure without a GPU)."""
import os
import re

import pytest
import torch

import msau_b200
from msau_b200 import _lib
from oracle import model as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "msau_b200.h")).read()
    declared = set(re.findall(r"\b(msau_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/msau_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS)
    assert L.msau_version() >= 100


def test_plan_rejects_bad_config_without_gpu_work():
    import ctypes as C
    L = _lib.lib()
    cfg = _lib.MsauConfig(96, 5, 4, 2, 8, 5, 2, 3)       # filter_size 5 unsupported
    h = C.c_void_p()
    rc = L.msau_plan_create(C.byref(cfg), 1, 64, 64, C.byref(h))
    assert rc == -1 and b"filter_size" in L.msau_last_error()
    cfg = _lib.MsauConfig(96, 5, 6, 3, 16, 3, 2, 3)      # S6 with featRoot 16: 512-channel levels are not covered
    rc = L.msau_plan_create(C.byref(cfg), 1, 64, 64, C.byref(h))
    assert rc == -3


@pytest.mark.parametrize("kw", [dict(featRoot=8, scale_space_num=4, res_depth=2), dict(featRoot=16, scale_space_num=2, res_depth=3)])
def test_state_dict_schema_matches_reference_order(kw):
    m = msau_b200.MSAUWrapper(96, 5, dict(final_act="softmax", **kw))
    cfg = om.MsauConfig(96, 5, kw["scale_space_num"], kw["res_depth"], kw["featRoot"])
    want = om.param_schema(cfg)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == want
    sd = om.init_state_dict(cfg, 3)
    m.load_state_dict(sd)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    # parameters are views of one flat buffer, in order
    off = 0
    for p in m.parameters():
        assert p.data_ptr() == m.flat_params.data_ptr() + 4 * off
        off += p.numel()
    assert off == m.flat_params.numel()


def test_save_load_roundtrip(tmp_path):
    m = msau_b200.MSAUWrapper(12, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    p = tmp_path / "w.pth"
    m.save(str(p))
    m2 = msau_b200.MSAUWrapper(12, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    m2.load_weights(str(p))
    assert torch.equal(m.flat_params, m2.flat_params)


def test_reference_quirks():
    with pytest.raises(TypeError):        # final_act default "sigmoid" -> torch.nn.Sigmoid(dim=1) raises (model.py:428-429)
        msau_b200.MSAUWrapper(1, 2, {})
    m = msau_b200.MSAUWrapper(12, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    with pytest.raises(msau_b200.MsauError):   # no CPU fallback
        m(torch.zeros(1, 12, 16, 16))


def test_no_cpu_fallback_anywhere():
    """Every public entry point fails loudly on a machine without a GPU instead of computing on the host."""
    import numpy as np
    from msau_b200 import kv_model, morph, raster
    m = msau_b200.MSAUWrapper(12, 5, dict(final_act="softmax", featRoot=8, scale_space_num=3, res_depth=2))
    x = torch.zeros(1, 12, 16, 16)
    lab = torch.ones(1, 16, 16, dtype=torch.int64)
    for call in (lambda: m.train_step(x, lab), lambda: m.train_step(x, lab, use_graph=True), lambda: m.predict_classes(x),
                 lambda: m.predict_classes_graph(x), lambda: m.set_feature_table(torch.zeros(3, 12)),
                 lambda: m.loss(x, x, lab)):
        with pytest.raises(msau_b200.MsauError):
            call()
    if not torch.cuda.is_available():
        with pytest.raises(msau_b200.MsauError):
            raster.rasterize_kv([np.array([[0, 0, 10, 10]], np.float64)], [[np.array([3], np.int32)]])
        with pytest.raises((msau_b200.MsauError, AssertionError, RuntimeError)):
            morph.r_closing(np.zeros((8, 8), bool), (1, 3))
        with pytest.raises((msau_b200.MsauError, AssertionError, RuntimeError)):
            kv_model.KVModel._extract_value(np.zeros((8, 8), np.uint16), np.zeros((8, 8), np.uint16), [], np.zeros((8, 8, 5), np.float32), 5)


def test_s6r3_defaults_are_plannable_and_counted():
    """The wrapper's own defaults (model/model.py:406-408: S=6, R=3, featRoot=8) build the reference's 13.08 M-parameter schema."""
    m = msau_b200.MSAUWrapper(96, 5, dict(final_act="softmax"))
    assert m.scale_space_num == 6 and m.res_depth == 3 and m.featRoot == 8
    assert m.flat_params.numel() == 13083687          # 13.08 M
