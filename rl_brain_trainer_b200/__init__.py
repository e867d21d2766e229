"""B200-native kinematic Approach -> Finisher rollout (drop-in for RL_brain_trainer's CPU kinematic env).

Importing the package is cheap and CPU-safe (configs, presets).  Anything that computes goes through the
hand-written sm_100a kernels in ``csrc/`` behind the C-ABI in ``include/kin_b200.h``; there is no CPU or
PyTorch fallback -- constructing an env without the built library or without a CUDA device raises.
"""

from .config import (  # noqa: F401
    ApproachRewardConfig,
    CurriculumStageConfig,
    DockRewardConfig,
    JointSpec,
    Phase1EnvConfig,
    PointCurriculumConfig,
    TerminationConfig,
    load_env_config_yaml,
    load_preset,
    to_env_config,
)

__version__ = "0.1.0"
