"""TEST INFRASTRUCTURE ONLY: CPU oracle (plain-C restatement) and reference-build wrappers.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
