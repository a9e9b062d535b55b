"""oracle/ -- CPU restatement of the reference hot path. TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) and nowhere else."""
