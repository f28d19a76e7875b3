"""Host-side kinematic bicycle helpers used by the tracker glue (same results as the reference's
/root/reference/src/control/vehicle_model.py:11-45).  The batched linearisation used by the solver is
the CUDA kernel behind ``MPCController.linearize_batch``; these NumPy versions serve single-vehicle
glue code (``TrajectoryTracker.track``'s integrator) and documentation of the contract."""
from __future__ import annotations

import numpy as np


def f_discrete(x, u, dt: float, wheelbase_px: float) -> np.ndarray:
    """One forward-Euler step of the bicycle model; state ``[x, y, yaw, v]``, input ``[a, delta]``."""
    px, py, yaw, v = x
    a, delta = u
    return np.array([px + dt * v * np.cos(yaw + 0.0),
                     py + dt * v * np.sin(yaw + 0.0),
                     yaw + dt * (v / wheelbase_px) * np.tan(delta),
                     v + dt * a], dtype=float)


def linearize(x, u, dt: float, wheelbase_px: float):
    """Jacobians ``A (4,4)``, ``B (4,2)`` about ``(x, u)`` and ``f(x, u)``; keeps the reference's
    ``sec^2 = 1 / (cos^2(delta) + 1e-9)`` regularisation."""
    yaw, v, delta = x[2], x[3], u[1]
    cy, sy = np.cos(yaw), np.sin(yaw)
    A = np.eye(4)
    A[0, 2], A[0, 3] = -dt * v * sy, dt * cy
    A[1, 2], A[1, 3] = dt * v * cy, dt * sy
    A[2, 3] = dt * (1.0 / wheelbase_px) * np.tan(delta)
    B = np.zeros((4, 2))
    B[3, 0] = dt
    B[2, 1] = dt * (v / wheelbase_px) * (1.0 / (np.cos(delta) ** 2 + 1e-9))
    return A, B, f_discrete(x, u, dt, wheelbase_px)
