"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (numpy / torch-CPU fp32) of the reference's algorithm for the hot path, pinned
against the real reference by tests/golden/* (generated in the build container by
oracle/gen_golden.py, which imports /root/reference).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product
(yolact_b200/) never does and has no CPU fallback.
"""
