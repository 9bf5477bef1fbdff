"""Output side of the hot path: the reference's trajectory dump and its TUM
conversion, fed by the device-side spline interpolation (hb200_interpolate).

Reference behaviour restated here:
  * apps/hyperslam/main.cpp:56-83 -- on SIGUSR1 the optimizer's state is sampled over `state.range()` at 100 Hz
    and every sample is written as one line `stamp, qx, qy, qz, qw, px, py, pz` (value only, `root + sample`
    as the stamp, scientific notation, precision 20, ", " separated): file `estimation.hyper`;
  * evaluation/conversions.py:5-8 -- `estimation.hyper` -> TUM trajectory: columns [0, 5, 6, 7, 1, 2, 3, 4]
    (stamp, position, quaternion), space separated, '%.20e'.
`Range::sample(rate)` lives in HyperState (not in the tree): [INFERRED] stamps lower + i / rate for every i with
lower + i / rate < upper (half-open, the range of a spline is [t_left, t_right)).
"""
from __future__ import annotations

import numpy as np

RATE = 100        # reference main.cpp:69
PRECISION = 20    # reference main.cpp:57


def state_range(knot_stamps: np.ndarray, order: int) -> tuple[float, float]:
    """Valid span of a uniform B-spline of `order` on these knots: [t_left, t_right) with (order-1)//2 control
    points of padding on the left and order-1-(order-1)//2 on the right (reference optimizer.cpp:288-290)."""
    left = (order - 1) // 2
    right = order - 1 - left
    return float(knot_stamps[left]), float(knot_stamps[len(knot_stamps) - 1 - right])


def sample_range(lower: float, upper: float, rate: float = RATE) -> np.ndarray:
    n = int(np.ceil((upper - lower) * rate - 1e-12))
    stamps = lower + np.arange(max(n, 0)) / rate
    return stamps[stamps < upper]


def format_line(stamp: float, pose: np.ndarray) -> str:
    """`<< std::scientific << stamp << ", " << value.transpose().format({20, DontAlignCols, ", ", "\\n"})`"""
    return ", ".join(f"{float(v):.{PRECISION}e}" for v in (stamp, *pose))


def write_estimation(ctx, path: str, knot_stamps: np.ndarray, order: int, root: float = 0.0, rate: float = RATE) -> int:
    """Writes `estimation.hyper` from the context's current state; returns the number of samples."""
    lower, upper = state_range(np.asarray(knot_stamps, dtype=np.float64), order)
    stamps = sample_range(lower, upper, rate)
    pose, _, _, bad = ctx.interpolate(stamps, derivatives=False)
    if bad:
        raise ValueError(f"{bad} sample(s) fell outside the state's range")
    with open(path, "w") as f:
        for t, row in zip(stamps, pose):
            f.write(format_line(root + t, row) + "\n")
    return int(stamps.size)


def convert_hyper_to_tum_format(input_path: str, output_path: str) -> None:
    data = np.loadtxt(input_path, delimiter=",", ndmin=2)
    data = data[:, [0, 5, 6, 7, 1, 2, 3, 4]]
    np.savetxt(output_path, data, fmt="%.20e")
