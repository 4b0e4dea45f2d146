"""b200-mpc: batched finite-horizon MPC solves on NVIDIA B200 (sm_100a).

Reference-shaped modules (same names and call signatures as konnpaku-youmu/Model_Predictive_Control):
``FHC``, ``LinearSystem``, ``session1_sol``, ``problem``, ``log``, ``session4``.
Device-level operators: ``lq``, ``boxqp``.  All compute runs in libmpc_b200.so.
"""
__version__ = "0.1.0"
