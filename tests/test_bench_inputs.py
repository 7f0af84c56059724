"""bench.py's native arm builds its pages with bench_inputs.synth_page (no oracle import on that arm); the CPU baseline arm
uses oracle.raster.synth_page.  Both must be the same workload."""
import numpy as np

import bench_inputs
from oracle import raster as orr


def test_bench_page_generator_equals_oracle_generator():
    for seed, gh, gw, n in ((0, 512, 512, 198), (7, 40, 48, 30), (1003, 512, 512, 198)):
        wa, la = bench_inputs.synth_page(seed, gh, gw, n)
        wb, lb = orr.synth_page(seed, gh, gw, n)
        for k in "xywh":
            assert wa[k].dtype == wb[k].dtype and np.array_equal(wa[k], wb[k]) and np.array_equal(la[k], lb[k])
        assert len(wa["chars"]) == len(wb["chars"]) and all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(wa["chars"], wb["chars"]))
        assert np.array_equal(la["label"], lb["label"]) and la["label"].dtype == lb["label"].dtype
