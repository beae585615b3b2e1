"""TEST INFRASTRUCTURE — CPU oracle for the two-view pose hot path (see tv5_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (deep-sfm-revisited_b200/) never does.
"""
from .oracle import *  # noqa: F401,F403
