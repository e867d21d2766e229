"""Does the reference's OWN bundled Approach checkpoint have a critic that matches its OWN env's returns?  (build container only)

Rolls the reference env (/root/reference) with the bundled checkpoint's deterministic policy on curriculum stages 5 and 11 and compares
the checkpoint critic's V(s) with the discounted return (gamma 0.98, TimeLimit bootstrap as SB3 does).  Result
(profiles/r2_reference_critic_check.txt): the value MSE is 3e4 .. 1e5 -- the critic was inherited through a chain of fine-tunes whose
reward scale differs from the final config's (success bonuses accrue every step once episodes no longer terminate on success).  A PPO
resume on that config therefore starts with a large value loss IN THE REFERENCE ITSELF; the on-device trainer's initial value_loss
(24 at gamma 0.98, on GAE targets) is not a trainer defect.
"""
import sys
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/hrl_ws/src/hrl_trainer")
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import numpy as np
import torch
from hrl_trainer.kinematic_phase1 import ArmKinematicEnv
from hrl_trainer.kinematic_phase1.training.policy_config import to_env_config
from rl_brain_trainer_b200 import config as kcfg

cfg = to_env_config(kcfg.preset_dict("approach_dynamic_scale_big"))
w = dict(np.load(kcfg.PRESET_DIR / "policies" / "approach_stage8_11.npz"))
T = lambda k: torch.as_tensor(w[k])  # noqa: E731


def mlp(x, net, head):
    h = torch.tanh(x @ T(f"mlp_extractor.{net}.0.weight").T + T(f"mlp_extractor.{net}.0.bias"))
    h = torch.tanh(h @ T(f"mlp_extractor.{net}.2.weight").T + T(f"mlp_extractor.{net}.2.bias"))
    return h @ T(f"{head}.weight").T + T(f"{head}.bias")


keys = None
for stage in (5, 11):
    env = ArmKinematicEnv(cfg)
    env.set_curriculum_stage(stage)
    errs, vals, rets, rew = [], [], [], []
    for ep in range(12):
        obs, _ = env.reset(seed=1000 + ep)
        rs, vs, done = [], [], False
        while not done:
            keys = keys or sorted(obs)
            x = torch.as_tensor(np.concatenate([np.asarray(obs[k], np.float32).ravel() for k in keys]))
            a = mlp(x, "policy_net", "action_net").clamp(-1, 1).numpy()
            vs.append(float(mlp(x, "value_net", "value_net")))
            obs, r, term, trunc, info = env.step(a)
            rs.append(r)
            done = term or trunc
        x = torch.as_tensor(np.concatenate([np.asarray(obs[k], np.float32).ravel() for k in keys]))
        G = float(mlp(x, "value_net", "value_net")) if trunc and not term else 0.0
        ret = []
        for r in reversed(rs):
            G = r + 0.98 * G
            ret.append(G)
        ret = ret[::-1]
        errs += [(v - g) ** 2 for v, g in zip(vs, ret)]
        vals += vs
        rets += ret
        rew += rs
    print(f"stage {stage}: reference env + bundled checkpoint: mean reward/step {np.mean(rew):.3f}; value MSE vs discounted return (gamma 0.98) = "
          f"{np.mean(errs):.1f}; mean V {np.mean(vals):.2f}, mean return {np.mean(rets):.2f}")
