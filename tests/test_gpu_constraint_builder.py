"""ConstraintBuilder2D on the device (CudaExecutor -> csm_match2d_batch) must produce
the same Result as the same queue drained through the CPU oracle."""
import numpy as np
import pytest

from cartographer_b200 import constraint_builder as cb
from tests.test_constraint_builder_cpu import OracleExecutor, _fill, _small_queue

pytestmark = pytest.mark.gpu


def test_constraint_builder_2d_device_equals_oracle():
    opts, submaps, clouds, poses = _small_queue()
    ref = cb.ConstraintBuilder2D(opts, executor=OracleExecutor(opts))
    _fill(ref, submaps, clouds, poses)
    want = ref.WhenDone(lambda r: None)
    dev = cb.ConstraintBuilder2D(opts)           # CudaExecutor
    _fill(dev, submaps, clouds, poses)
    got = dev.WhenDone(lambda r: None)
    assert len(want) > 0 and len(got) == len(want)
    for a, b in zip(got, want):
        assert a.submap_id == b.submap_id and a.node_id == b.node_id
        assert np.float32(a.score) == np.float32(b.score)
        assert a.zbar_ij == b.zbar_ij
    assert dev.executor.stats["searched"] == 15
    for sid in list(dev.executor.matchers):
        dev.DeleteScanMatcher(sid)
