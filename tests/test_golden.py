"""Golden vectors produced by the reference itself (tests/golden/make_golden.py,
from oracle/_ref = unmodified swgl/src/gl.cc): the oracle must reproduce them
byte for byte on any box (CPU tier); the CUDA backend likewise (GPU tier)."""
import json
import os

import numpy as np
import pytest

from oracle.backends import OracleDevice
from webrender_b200 import scenes

from common import render

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(HERE, "index.json")))


def _check(device_cls, name):
    case = INDEX[name]
    frame = getattr(scenes, case["builder"])(**case["kwargs"])
    got = render(device_cls, frame, case["targets"])
    want = np.load(os.path.join(HERE, name + ".npz"))
    for t in case["targets"]:
        assert np.array_equal(got[t], want[t]), f"{name}/{t}: {(got[t] != want[t]).sum()} bytes differ"


@pytest.mark.parametrize("name", sorted(INDEX))
def test_oracle_matches_reference_golden(name):
    _check(OracleDevice, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(INDEX))
def test_cuda_matches_reference_golden(name):
    from webrender_b200.device import CudaDevice
    _check(CudaDevice, name)
