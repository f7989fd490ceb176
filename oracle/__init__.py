"""ORACLE package — CPU restatements of the reference's hot path (test infrastructure only).

Import rules (enforced by tests/test_boundary_cpu.py): nothing under tokensgen_b200/ imports `oracle`; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
