"""Allele-level consumers on the CPU warp-emulator engine: crispresso2_b200/alleles.py against
the reference-generated fixtures of tests/golden/ (df_alleles, Alleles_frequency_table text, tables around the cut)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util as G  # noqa: E402
import parity_util as PU  # noqa: E402
from crispresso2_b200.engine import Engine  # noqa: E402


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    return Engine(lib_path=build_emu.build())


@pytest.mark.parametrize("case", G.CASES)
def test_allele_tables_equal_the_reference(emu, case, tmp_path):
    assert PU.check_alleles(emu, case, tmp_path) > 100


def test_foreign_dataframe_is_refused(emu):
    import pandas as pd
    from crispresso2_b200 import alleles
    with pytest.raises(TypeError):
        alleles.get_dataframe_around_cut_asymmetrical(pd.DataFrame({"Aligned_Sequence": ["A"]}), 1, 1, 1)
