"""``VehicleParameters`` for the session-4 path.

Interface-compatible with the reference's ``session_4/parameters.py`` (a dataclass with these field
names and default values, /root/reference/session_4/parameters.py:4-54), so that code written
against the reference -- ``VehicleParameters()``, ``params.friction *= 0.8``, ``params.max_steer``
-- runs unchanged.  The class is generated from the table below.  The GPU path uses the geometry
(``axis_front``, ``axis_rear``), the input / state limits and the two kinematic-model parameters
(``friction``, ``acceleration``); tyre and motor coefficients are carried for compatibility only.
"""
from dataclasses import make_dataclass, field
import math

#            name           default     meaning
_FIELDS = [
    ("length",        0.17,       "car length [m]"),
    ("axis_front",    0.047,      "centre of gravity to front axis [m] (l_f of the kinematic model)"),
    ("axis_rear",     0.05,       "centre of gravity to rear axis [m] (l_r of the kinematic model)"),
    ("front",         0.08,       "centre of gravity to front end [m]"),
    ("rear",          0.08,       "centre of gravity to rear end [m]"),
    ("width",         0.08,       "car width [m]"),
    ("height",        0.055,      "car height [m]"),
    ("mass",          0.1735,     "mass [kg]"),
    ("inertia",       18.3e-5,    "yaw inertia [kg m^2]"),
    ("max_steer",     0.384,      "steering angle limit [rad] (input box, both signs)"),
    ("max_drive",     1.0,        "largest normalised drive command"),
    ("min_drive",     -1.0,       "largest reverse drive command"),
    ("min_pos_x",     -3.0,       "state box: p_x lower [m]"),
    ("max_pos_x",     3.0,        "state box: p_x upper [m]"),
    ("min_pos_y",     -2.0,       "state box: p_y lower [m]"),
    ("max_pos_y",     2.0,        "state box: p_y upper [m]"),
    ("min_vel",       -0.5,       "state box: v lower [m/s]"),
    ("max_vel",       0.5,        "state box: v upper [m/s]"),
    ("max_heading",   2 * math.pi,  "state box: psi upper [rad]"),
    ("min_heading",   -2 * math.pi, "state box: psi lower [rad]"),
    ("bf",            3.1355,     "Pacejka front stiffness"),
    ("cf",            2.1767,     "Pacejka front shape"),
    ("df",            0.4399,     "Pacejka front peak"),
    ("br",            2.8919,     "Pacejka rear stiffness"),
    ("cr",            2.4431,     "Pacejka rear shape"),
    ("dr",            0.6236,     "Pacejka rear peak"),
    ("friction",      1,          "kinematic model: v' = acceleration * a - friction * v"),
    ("acceleration",  2,          "kinematic model: drive-command gain"),
    ("cm1",           0.3697,     "motor coefficient"),
    ("cm2",           0.001295,   "motor coefficient"),
    ("cr1",           0.1629,     "rolling-resistance coefficient"),
    ("cr2",           0.02133,    "rolling-resistance coefficient"),
]

VehicleParameters = make_dataclass(
    "VehicleParameters", [(name, float, field(default=default)) for name, default, _ in _FIELDS])
VehicleParameters.__doc__ = "Vehicle geometry, limits and model coefficients:\n" + "\n".join(
    f"    {name}: {doc} (default {default})" for name, default, doc in _FIELDS)
VehicleParameters.__module__ = __name__
