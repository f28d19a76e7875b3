"""``MPCConfig`` mirror (defaults and ``to_parameters`` of /root/reference/src/config.py:66-92).

The tracker accepts the reference's own ``MPCConfig`` / ``VizConfig`` objects as well (duck typing);
this mirror exists so that the package is importable without cvxpy / matplotlib, which the
reference's ``src/__init__.py`` import chain requires."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

from .mpc_controller import MPCParameters


@dataclass
class MPCConfig:
    wheelbase_m: float = 2.8
    dt: float = 0.1
    horizon: int = 15
    v_px_s: float = 15.0
    sim_steps: int = 300
    q: Tuple[Tuple[float, ...], ...] = ((4.0, 0.0, 0.0, 0.0), (0.0, 4.0, 0.0, 0.0), (0.0, 0.0, 0.6, 0.0), (0.0, 0.0, 0.0, 0.1))
    r: Tuple[Tuple[float, ...], ...] = ((0.03, 0.0), (0.0, 0.25))
    q_terminal: Tuple[Tuple[float, ...], ...] = ((8.0, 0.0, 0.0, 0.0), (0.0, 8.0, 0.0, 0.0), (0.0, 0.0, 1.0, 0.0), (0.0, 0.0, 0.0, 0.2))
    u_bounds: Tuple[Tuple[float, float], Tuple[float, float]] = ((-35.0, 35.0), (-0.6, 0.6))
    v_bounds: Tuple[float, float] = (0.0, 90.0)
    du_bounds: Tuple[Tuple[float, float], Tuple[float, float]] = ((-12.0, 12.0), (-0.15, 0.15))

    def to_parameters(self, map_resolution: float) -> MPCParameters:
        return MPCParameters(
            wheelbase_px=self.wheelbase_m / map_resolution, dt=self.dt, horizon=self.horizon,
            q=np.array(self.q, dtype=float), r=np.array(self.r, dtype=float),
            q_terminal=np.array(self.q_terminal, dtype=float),
            u_bounds=self.u_bounds, v_bounds=self.v_bounds, du_bounds=self.du_bounds)


@dataclass
class VizConfig:
    backend: str = "auto"
    prediction_pause: float = 0.01
    animate_tree: bool = True
    record_frames: bool = False
    record_dir: str = "plots/frames"


def params_from_config(mpc, map_resolution: float) -> MPCParameters:
    """``mpc.to_parameters(map_resolution)`` for either this mirror or the reference's MPCConfig
    (whose result is the reference's own MPCParameters dataclass — converted field by field)."""
    p = mpc.to_parameters(map_resolution)
    if isinstance(p, MPCParameters):
        return p
    return MPCParameters(**{f: getattr(p, f) for f in MPCParameters.__dataclass_fields__})
