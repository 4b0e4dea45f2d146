"""Per-step controller log of sessions 2/3.

The reference's schema (/root/reference/session_2/log.py:8-12, a dataclass deriving from rcracers'
``BaseControllerLog``, which is not vendored) is three growing lists: ``solver_success``,
``state_prediction`` and ``input_prediction``.  Here the same three attributes are the contract;
a batched solve appends one array per control step to each of them (``LinearMPC.__call__``)."""
from dataclasses import make_dataclass, field

LOG_FIELDS = ("solver_success", "state_prediction", "input_prediction")

ControllerLog = make_dataclass("ControllerLog", [(name, list, field(default_factory=list)) for name in LOG_FIELDS])
ControllerLog.__module__ = __name__
ControllerLog.__doc__ = "Lists, one entry per control step: " + ", ".join(LOG_FIELDS)


def n_steps(log) -> int:
    """Number of control steps recorded."""
    return len(log.solver_success)
