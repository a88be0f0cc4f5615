"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference's algorithm (the oracle) and loaders for the
reference's own compiled kernels (oracle/_ref).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package, and only as the checker."""
