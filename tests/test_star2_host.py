"""CPU check of the fused star-CTC lane arithmetic: csrc/star2_math.h is host + device code, tools/star2_host_check.cpp
steps the shipped functions lane by lane (J = 1, 2, 4 quads per lane) with the kernels' index formulas, and the result is
held to the float64 oracle (ha/star.py:65-163 restated) at the GPU tests' tolerances.  Runs without a GPU."""
import importlib.util
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_shipped_lane_arithmetic_matches_the_oracle_on_the_cpu(oracle):
    spec = importlib.util.spec_from_file_location("star2_host_check", os.path.join(ROOT, "tools", "star2_host_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main() == 0
