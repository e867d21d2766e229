"""BASELINE.json configs[0] on the reference's OWN Python CPU path (only where /root/reference exists: the build container).

Stage-0 kinematic Approach env, single env, random-init 56-64-64-7 tanh MLP policy in the loop (torch, CPU), reset(seed=0), 1 000 steps,
auto-reset on done -- single process, then one process per host core (os.cpu_count() printed).  Writes
profiles/r2_config1_reference_python.json, which bench.py quotes beside the C port's live numbers (the GPU box has no /root/reference).
"""
import json, multiprocessing as mp, os, sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = "/root/reference/hrl_ws/src/hrl_trainer"
STEPS = 1000


def run(steps: int) -> tuple[int, float]:
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.path.insert(0, str(ROOT))
    import numpy as np
    import torch
    from hrl_trainer.kinematic_phase1 import ArmKinematicEnv
    from hrl_trainer.kinematic_phase1.training.policy_config import to_env_config
    from rl_brain_trainer_b200 import config as kcfg

    torch.set_num_threads(1)
    env = ArmKinematicEnv(to_env_config(kcfg.preset_dict("approach_dynamic_scale_big")))
    env.set_curriculum_stage(0)
    torch.manual_seed(0)
    mlp = torch.nn.Sequential(torch.nn.Linear(56, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(), torch.nn.Linear(64, 7))
    keys = sorted(env.observation_space.spaces.keys()) if hasattr(env.observation_space, "spaces") else None
    obs, _ = env.reset(seed=0)
    flat = lambda o: np.concatenate([np.asarray(o[k], dtype=np.float32).ravel() for k in (keys or sorted(o))])  # noqa: E731
    t0 = time.perf_counter()
    with torch.no_grad():
        for _ in range(steps):
            a = mlp(torch.from_numpy(flat(obs))).clamp(-1, 1).numpy()
            obs, _, term, trunc, _ = env.step(a)
            if term or trunc:
                obs, _ = env.reset()
    return steps, time.perf_counter() - t0


def _worker(q, steps):
    q.put(run(steps))


if __name__ == "__main__":
    if not Path(REF).exists():
        raise SystemExit("the reference tree is not present here")
    s, dt = run(STEPS)
    cores = os.cpu_count() or 1
    q = mp.Queue()
    procs = [mp.Process(target=_worker, args=(q, STEPS)) for _ in range(cores)]
    t0 = time.perf_counter()
    for p in procs:
        p.start()
    got = [q.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.perf_counter() - t0
    out = {"what": "reference Python env (hrl_trainer.kinematic_phase1.ArmKinematicEnv), Stage 0, random-init torch MLP in the loop, 1 000 steps, auto-reset",
           "where": "build container CPU (not the GPU box)", "os_cpu_count": cores,
           "single_process": {"env_steps_per_s": s / dt, "env_steps": s, "seconds": dt},
           "one_process_per_core": {"env_steps_per_s_sum": sum(n / t for n, t in got), "env_steps_per_s_wall": sum(n for n, _ in got) / wall,
                                    "processes": cores, "per_process": [n / t for n, t in got]}}
    (ROOT / "profiles" / "r2_config1_reference_python.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))
