"""Drop-in for the reference's ``session_2/log.py`` / ``session_3/log.py``: the per-step controller
log schema (/root/reference/session_2/log.py:8-12).  The reference derives from
``rcracers.simulator.core.BaseControllerLog`` (not vendored); the three list fields are the
contract, and a batched solve appends one array per control step to each of them."""
from dataclasses import dataclass, field


def new_list():
    return field(default_factory=list)


@dataclass
class ControllerLog:
    solver_success: list = new_list()
    state_prediction: list = new_list()
    input_prediction: list = new_list()
