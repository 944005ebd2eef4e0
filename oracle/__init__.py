"""ORACLE — test infrastructure only.

CPU / fp32 restatements of the reference's hot-path arithmetic, used exclusively by tests/,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs as the CHECKER.
Nothing under geo-deep-learning_b200/ imports this package.
"""
