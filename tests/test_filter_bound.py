"""The error bound of the filtered projection against the oracle's fp32 chain, on the CPU:
tools/filter_proto.py evaluates the plane-induced homography in fp32, applies the bound, and compares every
pixel the bound decides with oracle/restated.py.  No decided pixel may differ, the observed error must stay
inside the bound, and the bound must decide nearly all pixels (else the GPU kernel would gain nothing)."""
import importlib.util
import os

import pytest

from tests import golden_util as gu

spec = importlib.util.spec_from_file_location("filter_proto", os.path.join(gu.ROOT, "tools", "filter_proto.py"))
filter_proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(filter_proto)


@pytest.mark.parametrize("mode", ["seq", "composed", "translate"])
def test_bound_decides_only_what_the_reference_chain_confirms(mode):
    st = filter_proto.run(11, mode, n_frames=10, frame_step=5, quiet=True)
    assert st["n"] > 1_000_000
    assert st["wrong"] == 0                               # every proven pixel equals the oracle's
    assert st["maxratio"] < 1.0                           # |q_cheap - q_ref| inside the bound wherever it matters
    assert st["unc"] / st["n"] < 0.08                     # incl. the candidates that map pixels onto themselves


@pytest.mark.parametrize("seed", [1, 2, 4, 5])
def test_bound_under_adversarial_geometry(seed):
    """Grazing / fronto-parallel planes, planes centimetres from the camera, pivots 100 m away, full-circle
    rotations, metres of translation, focal lengths of 40 and 6000 px: no proven pixel differs from the oracle's
    chain and the observed error stays inside the bound."""
    st = filter_proto.run_random(seed, quiet=True)
    assert st["n"] > 100_000 and st["wrong"] == 0 and st["maxratio"] < 1.0
