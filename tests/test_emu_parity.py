"""CPU-only parity tests of the KERNEL LOGIC: crispresso2_b200/csrc/c2b_core.cuh compiled with g++ against the
fiber warp emulator (tests/emu/), driven through the same C ABI and the same Python host code as the GPU build.
These prove the algorithm (wavefront DP, tagged tie-breaks, traceback window, row-space classification, count
block) against the reference goldens before any GPU time is spent; tests/test_gpu_parity.py repeats them on the
real sm_100a build."""
import numpy as np
import pytest

import golden_util as G
import parity_util as PU
from crispresso2_b200 import _lib, synth
from crispresso2_b200.engine import Engine
from oracle import oracle as O


@pytest.fixture(scope="module")
def emu():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    return Engine(lib_path=build_emu.build())


@pytest.mark.parametrize("case", ["fanc_cas9", "fanc_params", "synth_hdr"])
def test_golden_whole_path(emu, case, tmp_path):
    PU.check_golden_case(emu, case, tmp_path)


def test_golden_synth_single_subset(emu, tmp_path):
    PU.check_golden_case(emu, "synth_single", tmp_path, max_reads=300)


def test_align_vectors(emu):
    import gzip, json, os
    from crispresso2_b200 import align, resources
    with gzip.open(os.path.join(G.GOLD, "align_vectors.json.gz"), "rt") as fh:
        cases = json.load(fh)
    m = O.make_matrix()
    for c in cases[::3]:
        got = align.global_align(c["read"], c["ref"], m, np.array(c["gi"], dtype=np.int64), c["go"], c["ge"], engine=emu)
        assert got == (c["s1"], c["s2"], c["score"]), c
        a, ed = emu.classify_pair(c["s1"], c["s2"], c["inc"])
        p = resources.payload_from_device(a, ed, c["s1"], c["s2"])
        assert not G.payload_equal(c["payload"], p), (c, p.__dict__)
