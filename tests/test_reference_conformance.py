"""The reference's own five hot-path test files against the drop-in env (tools/run_reference_tests.py).

Needs the reference checkout AND a B200 in one machine; this build has them in different boxes (see the tool's docstring), so:
* here (reference present, no GPU): the patching step is checked -- the reference modules really hand out the drop-in classes and
  every test file loads against them;
* with both present the whole suite runs.
"""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

REF = Path("/root/reference")
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))


@pytest.mark.skipif(not REF.exists(), reason="the reference checkout is not on this machine")
def test_reference_tests_load_against_the_drop_in_classes():
    import run_reference_tests as rrt
    from rl_brain_trainer_b200.env import ArmKinematicEnv

    mod = rrt.patch_reference(REF)
    assert mod.ArmKinematicEnv is ArmKinematicEnv
    suite = rrt.load_suite(REF)
    assert suite.countTestCases() >= 40        # env 4, eval, reward, approach_reward, split 37+


@pytest.mark.gpu
@pytest.mark.skipif(not REF.exists(), reason="the reference checkout is not on this machine (the GPU box has no /root/reference)")
def test_reference_tests_pass_against_the_gpu_adapter():
    import unittest

    import run_reference_tests as rrt

    rrt.patch_reference(REF)
    result = unittest.TextTestRunner(verbosity=0).run(rrt.load_suite(REF))
    assert result.wasSuccessful(), (result.failures[:3], result.errors[:3])
