"""CPU tests of the handoff-state buffer's host logic (JSON reader semantics of reset_samplers.py:131-166, row packing)."""

from __future__ import annotations

import json

import numpy as np
import pytest

from rl_brain_trainer_b200 import handoff
from rl_brain_trainer_b200 import _lib


def _states(n=6):
    rng = np.random.default_rng(0)
    return [{"episode_id": i, "step_index": 100 + i, "initial_q": rng.normal(size=7).tolist(), "goal_q": rng.normal(size=7).tolist(),
             "goal_pose6": rng.normal(size=6).tolist(), "initial_dq": rng.normal(size=7).tolist(),
             "initial_prev_action": rng.normal(size=7).tolist(), "position_error_norm": 0.002 * (i + 1),
             "orientation_error_norm": 0.02 * (i + 1), "action_l2": 0.1 * (i + 1), "dq_norm": 0.001} for i in range(n)]


def test_reader_filters_and_payload_shapes(tmp_path):
    st = _states()
    p = tmp_path / "buf.json"
    p.write_text(json.dumps({"states": st}))
    buf = handoff.load_handoff_states(p)
    assert len(buf) == 6 and buf.rows().shape == (6, _lib.define("KIN_HANDOFF_STATE_FLOATS")) and buf.rows().dtype == np.float32
    row = buf.rows()[2]
    assert np.allclose(row[:7], st[2]["initial_q"]) and np.allclose(row[7:14], st[2]["initial_dq"]) and np.allclose(row[14:21], st[2]["initial_prev_action"])
    assert np.allclose(row[21:28], st[2]["goal_q"]) and np.allclose(row[28:34], st[2]["goal_pose6"])
    # each filter drops the states ABOVE its bound (strict >), like the reference
    assert len(handoff.load_handoff_states(p, max_position_error_m=0.006)) == 3
    assert len(handoff.load_handoff_states(p, max_orientation_error_rad=0.04)) == 2
    assert len(handoff.load_handoff_states(p, max_action_l2=0.5)) == 5
    # a bare list is accepted; missing dq / prev_action default to zeros
    bare = [{k: v for k, v in s.items() if k not in ("initial_dq", "initial_prev_action")} for s in st]
    p.write_text(json.dumps(bare))
    b2 = handoff.load_handoff_states(p)
    assert len(b2) == 6 and not b2.initial_dq.any() and not b2.initial_prev_action.any()
    with pytest.raises(FileNotFoundError):
        handoff.load_handoff_states(tmp_path / "nope.json")


def test_states_round_trip_and_modes():
    buf = handoff.HandoffBuffer.from_states(_states(4))
    out = buf.to_states(dwell_count=3, source_checkpoint_name="ckpt.zip", handoff_mode="first_confirmed")
    assert out[1]["dwell_count"] == 3 and out[1]["handoff_mode"] == "first_confirmed" and out[1]["step_index"] == 101
    again = handoff.HandoffBuffer.from_states(out)
    assert np.array_equal(again.rows(), buf.rows()) and np.array_equal(again.episode_id, buf.episode_id)
    assert handoff.HANDOFF_MODES == ("final_settled", "first_confirmed", "final_always")
    with pytest.raises(ValueError):
        handoff.build_finisher_handoff_state_buffer(None, None, handoff_mode="sometimes")
