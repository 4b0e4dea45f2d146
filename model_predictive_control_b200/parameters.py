"""Drop-in for the reference's ``session_4/parameters.py``: ``VehicleParameters`` with the same field
names and default values (/root/reference/session_4/parameters.py:4-54), so that code written
against the reference (``VehicleParameters()``, ``params.friction *= 0.8``, ``params.max_steer``)
runs unchanged.  Only the geometry, limits and the two kinematic-model parameters are used by the
GPU path; the tyre and motor coefficients are carried for interface compatibility."""
from dataclasses import dataclass

import numpy as np


@dataclass
class VehicleParameters:
    # geometry [m], mass [kg], inertia [kg m^2]
    length: float = 0.17
    axis_front: float = 0.047
    axis_rear: float = 0.05
    front: float = 0.08
    rear: float = 0.08
    width: float = 0.08
    height: float = 0.055
    mass: float = 0.1735
    inertia: float = 18.3e-5
    # input limits
    max_steer: float = 0.384
    max_drive: float = 1.0
    min_drive: float = -1.
    # state limits
    min_pos_x: float = -3.
    max_pos_x: float = 3.
    min_pos_y: float = -2.
    max_pos_y: float = 2.
    min_vel: float = -0.5
    max_vel: float = 0.5
    max_heading: float = 2 * np.pi
    min_heading: float = -2 * np.pi
    # Pacejka tyre coefficients (front / rear): stiffness, shape, peak
    bf: float = 3.1355
    cf: float = 2.1767
    df: float = 0.4399
    br: float = 2.8919
    cr: float = 2.4431
    dr: float = 0.6236
    # kinematic approximation
    friction: float = 1
    acceleration: float = 2
    # motor
    cm1: float = 0.3697
    cm2: float = 0.001295
    cr1: float = 0.1629
    cr2: float = 0.02133
