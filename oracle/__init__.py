"""CPU oracle for the VisualRWKV hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (visualrwkv_b200/) never does.
"""
