"""Parity oracle for the gVAMP bed hot path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (gvamp_b200/) never does.
"""
