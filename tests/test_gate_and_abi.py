"""CPU tests: gate-score known answers from the reference, C-ABI symbol coverage, config mirror, checkpoint round trip."""

from __future__ import annotations

import ctypes
import json
import re

import numpy as np
import pytest

from rl_brain_trainer_b200 import _lib, config as kcfg, gate

from ._util import GOLD


def test_gate_score_matches_reference_known_answers():
    cases = json.loads((GOLD / "gate_cases.json").read_text())
    assert len(cases) >= 20
    seen_ok = set()
    for c in cases:
        metrics = {int(k): v for k, v in c["stage_metrics"].items()}
        out = gate.gated_score(metrics, c["current_stage"], gate.gate_config_from_dict(c["gate_config"]))
        exp = c["expected"]
        for k, v in exp.items():
            if isinstance(v, float):
                assert out[k] == pytest.approx(v, abs=1e-12), k
            else:
                assert out[k] == v, k
        seen_ok.add(bool(exp["retention_ok"]))
    assert seen_ok == {True, False}


def test_library_exports_every_declared_symbol():
    """The built .so must export exactly the functions include/kin_b200.h declares (no compute calls without a GPU)."""
    names = _lib.declared_functions()
    assert {"kin_env_step", "kin_env_reset", "kin_rollout_approach_finisher", "kin_route_step", "kin_ppo_grad", "kin_policy_act"} <= set(names)
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(L, n), f"libkin_b200.so lacks {n}"
    assert L.kin_abi_version() == _lib.define("KIN_ABI_VERSION")
    # argument validation happens before any CUDA call: a null handle is rejected with an error string
    L.kin_last_error_string.restype = ctypes.c_char_p
    L.kin_env_step.argtypes = _lib.lib().kin_env_step.argtypes
    rc = L.kin_env_step(None, None, 32, 1, 0, None, None, None, None, None, None, 0, 0, None, None)
    assert rc == _lib.define("KIN_ERR_INVALID_ARG") and b"bad handle" in L.kin_last_error_string()
    assert _lib.lib().kin_ppo_param_count(56) == 16143     # 8 270 actor (incl. log_std) + 7 873 critic


def test_library_is_built_from_the_sources_in_the_tree():
    """The library carries the sha256 of the sources it was compiled from; the loader refuses (after one rebuild) a stale one."""
    from rl_brain_trainer_b200 import build as kbuild

    want = kbuild.source_hash()
    assert len(want) == 64 and _lib.lib().kin_source_hash().decode() == want == _lib.library_source_hash()
    files = [f.name for f in kbuild.source_files()]
    assert "kin_b200.h" in files and "kin_step.cu" in files and "kin_rollout_tc16.cu" in files and "kin_ppo_tc.cu" in files


def test_header_structs_parse_and_params_fill():
    from rl_brain_trainer_b200 import params

    cfg = kcfg.load_preset("approach_dynamic_scale_big")
    p = params.env_params(cfg)
    assert p.ar_finisher_ready_pos_threshold_m == pytest.approx(0.005) and p.term_max_episode_steps == 128 and p.episode_length == 128
    assert p.dynamic_action_delta_scale_enabled == 1 and p.k_inv_delta_limit[0] == pytest.approx(1 / 0.08)
    assert abs(sum(x * x for x in p.fk_AT) - 3.0) < 1e-5      # a rotation matrix
    s = params.sampler_params(cfg, 9)
    assert s.n_stages == 12 and s.current_stage == 9 and s.stage_mix_enabled == 1 and s.current_stage_ratio == pytest.approx(0.70)
    assert s.previous_stage_min_index == 7 and s.old_workspace_max_stage_index == 5
    rs = params.sampler_params(kcfg.load_preset("randomstart_overnight"), 10)
    assert rs.random_start_enabled == 1 and rs.known_target_max_stage_index == 8 and rs.min_pair_joint_l2 == pytest.approx(0.04)
    fields = [f[0] for f in _lib.c_struct("KinEnvParams")._fields_]
    header = _lib.HEADER.read_text()
    assert all(re.search(r"\b%s\b" % f, header) for f in fields)


def test_config_mirror_defaults_and_yaml_semantics(tmp_path):
    d = kcfg.Phase1EnvConfig()
    assert d.episode_length == 75 and d.dwell_steps_target == 3 and d.reward_config.near_goal_bonus_decay == 0.5
    assert d.termination_config.success_pos_threshold_m == 0.06 and len(d.curriculum_config.stages) == 6
    fin = kcfg.load_preset("finisher_noop_ft")
    assert fin.mode_name == "dock" and fin.dwell_steps_target == 5 and not fin.curriculum_config.enabled
    assert fin.dock_dynamic_residual_action_limit_near == fin.dock_residual_action_limit == 1.0    # fallback chain of policy_config.py
    with pytest.raises(TypeError):
        kcfg.to_env_config({"env": {"reward": {"not_a_knob": 1.0}}})
    base = tmp_path / "base.yaml"
    base.write_text("env:\n  episode_length: 90\n  reward:\n    near_goal_bonus: 0.2\n")
    over = tmp_path / "over.yaml"
    over.write_text("base_config: base.yaml\nenv:\n  reward:\n    dwell_bonus: 0.3\n")
    cfg = kcfg.load_env_config_yaml(over)
    assert cfg.episode_length == 90 and cfg.reward_config.near_goal_bonus == 0.2 and cfg.reward_config.dwell_bonus == 0.3
    with pytest.raises(ValueError):
        kcfg.CurriculumStageConfig(name="x", start_q=(0,) * 6, goal_q=(0,) * 7)


def test_policy_checkpoint_round_trip(tmp_path):
    torch = pytest.importorskip("torch")
    from rl_brain_trainer_b200.policy import KEYS, PolicyWeights

    pol = PolicyWeights.preset("finisher", "cpu")
    z = tmp_path / "model.zip"
    pol.save_weights_zip(z)
    back = PolicyWeights.load(z, "cpu")
    for k in KEYS.values():
        assert torch.equal(pol.state_dict()[k], back.state_dict()[k])
    assert back.in_dim == 56 and back.has_value
    npz = tmp_path / "p.npz"
    pol.save_npz(npz)
    assert torch.equal(PolicyWeights.load(npz, "cpu").tensors["act_w"], pol.tensors["act_w"])
    assert np.isfinite(pol.tensors["log_std"].numpy()).all()
