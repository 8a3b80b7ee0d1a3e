"""CPU oracle for the flexlmm FIT_MODEL hot path — TEST INFRASTRUCTURE ONLY (parity unpinned).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
