"""Robot description of the 7-joint rack + 6R arm and host-side folding of its constant transforms.

The reference hard-codes the chain in ``v5_1/ee_fk.py:14-61`` (joint types, origin xyz / rpy and local
axes exported from the URDF); those numbers are the robot's geometry and are restated here as data.
``fold_chain`` pre-multiplies everything that does not depend on the joint values, in fp64, into the
``fk_*`` fields of ``KinEnvParams`` (include/kin_b200.h) so the device FK spends one sincos and ~48
flops per revolute joint:

    T = prod_i [O_i * J_i(q_i)],  J_0 = Trans(a_0 q_0),  J_i = Rot(a_i, q_i) = A_i Rz(q_i) A_i^T

with A_i a fixed rotation taking e_z to the joint axis.  Folding A_{i-1}^T O_i A_i into one constant
3x3 per joint leaves only the Rz(q_i) factors at run time.
"""

from __future__ import annotations

import math

import numpy as np

JOINT_TYPES = ("prismatic", "revolute", "revolute", "revolute", "revolute", "continuous", "revolute")

ORIGIN_XYZ = np.array([
    [0.00715921043213119, 0.0000809621375843506, -0.0635],
    [-0.021178, 0.0, 0.1868],
    [-0.0633967414837172, 0.000642782425827271, 0.0602000000000009],
    [-0.000134989688424625, 0.425, 0.0133123982251372],
    [-0.0000850456535865796, -0.39225, -0.0083864861805065],
    [0.0475482889721905, -0.000817137634885778, -0.0805958577476871],
    [0.0436977540622506, 0.000443046177049933, -0.0521517110277254],
])

ORIGIN_RPY = np.array([
    [0.0, 0.0, 0.0],
    [0.0, 0.0, 0.0],
    [1.5707963267949, 0.0, 1.5707963267949],
    [3.14159265358979, 0.0, 0.0],
    [3.14159265358979, 0.0, -1.5707963267949],
    [3.14159265358979, 1.5707963267949, 0.0],
    [-1.5707963267949, 0.0, -1.5707963267949],
])

AXES_LOCAL = np.array([
    [1.0, 0.0, 0.0],
    [0.0, 0.0, 1.0],
    [0.0101382310641698, 0.0, -0.999948606814815],
    [0.010138231064165, 0.0, 0.999948606814815],
    [0.0, -0.0101382310641647, -0.999948606814815],
    [0.0, 0.0, -1.0],
    [-0.0101384515502096, 0.0, 0.999948604579338],
])


def rpy_matrix(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """Rz(yaw) @ Ry(pitch) @ Rx(roll) -- the URDF convention used by ee_fk.py:64-71."""
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def axis_frame(axis: np.ndarray) -> np.ndarray:
    """A proper rotation A with A @ e_z == axis / |axis|, built by Gram-Schmidt (orthonormal to 1e-16).

    ee_fk.py:76 divides by (|axis| + 1e-12); that 1e-12 relative shortening of the axis is far below fp32
    resolution and is not reproduced.
    """
    a = np.asarray(axis, dtype=float)
    a = a / np.linalg.norm(a)
    ref = np.eye(3)[int(np.argmin(np.abs(a)))]
    u = np.cross(ref, a)
    u = u / np.linalg.norm(u)
    v = np.cross(a, u)
    return np.stack([u, v, a], axis=1)


def fold_chain() -> dict[str, np.ndarray]:
    """Constants of the folded chain: ``pbase[3], pq0[3], C[6,3,3], t[5,3], AT[3,3]`` (fp64)."""
    if JOINT_TYPES[0] != "prismatic" or any(t == "prismatic" for t in JOINT_TYPES[1:]):
        raise ValueError("fold_chain expects one leading prismatic joint followed by revolute joints")
    Ro = [rpy_matrix(*ORIGIN_RPY[i]) for i in range(7)]
    A = [np.eye(3)] + [axis_frame(AXES_LOCAL[i]) for i in range(1, 7)]
    # after joint 0: R = Ro0, p = po0 + Ro0 a0 q0
    R0 = Ro[0]
    pq0 = R0 @ AXES_LOCAL[0]
    pbase = ORIGIN_XYZ[0] + R0 @ ORIGIN_XYZ[1]
    C = np.zeros((6, 3, 3))
    t = np.zeros((5, 3))
    C[0] = R0 @ Ro[1] @ A[1]
    for j in range(2, 7):
        t[j - 2] = A[j - 1].T @ ORIGIN_XYZ[j]
        C[j - 1] = A[j - 1].T @ Ro[j] @ A[j]
    return {"pbase": pbase, "pq0": pq0, "C": C, "t": t, "AT": A[6].T}


def fk_pose6_folded(q: np.ndarray) -> np.ndarray:
    """fp64 evaluation of the folded chain (host-side check of ``fold_chain`` against the plain chain)."""
    k = fold_chain()
    q = np.asarray(q, dtype=float)

    def rz(a: float) -> np.ndarray:
        c, s = math.cos(a), math.sin(a)
        return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])

    p = k["pbase"] + k["pq0"] * q[0]
    R = k["C"][0] @ rz(q[1])
    for j in range(2, 7):
        p = p + R @ k["t"][j - 2]
        R = R @ k["C"][j - 1] @ rz(q[j])
    R = R @ k["AT"]
    return np.array([p[0], p[1], p[2], math.atan2(R[2, 1], R[2, 2]), math.atan2(-R[2, 0], math.hypot(R[0, 0], R[1, 0])),
                     math.atan2(R[1, 0], R[0, 0])])
