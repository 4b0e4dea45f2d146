"""Drop-in for the reference's ``session_3/problem.py``: the same ``Problem`` dataclass as
``session_2/problem.py`` with ``p_min = -120`` and ``v_min = -50`` (/root/reference/session_3/problem.py:8-36;
the two defaults that differ are lines 15 and 17).  ``from model_predictive_control_b200.problem3 import Problem``
replaces ``from problem import Problem`` of a session-3 script; the solver and the closed-loop driver are the ones of
:mod:`model_predictive_control_b200.problem`."""
from .problem import LinearMPC, closed_loop  # noqa: F401
from .problem import Problem3 as Problem

Problem.__doc__ = "Convenience class representing the problem data for session 3."
__all__ = ["Problem", "LinearMPC", "closed_loop"]
