"""CPU oracle for the Horn-Schunck hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package.  opticalflowhs_b200 never does (no CPU fallback in the product).
"""
from .oracle import *  # noqa: F401,F403
