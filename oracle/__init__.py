"""TEST INFRASTRUCTURE ONLY.

oracle/ holds the parity checkers of the hot path: CPU restatements of the reference's algorithms and a loader
for the reference's own CUDA extensions (oracle/_ref, built from the unmodified sources under /root/reference by
oracle/build_ref.sh).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may
import it; nothing under raw_ngp_b200/ does.
"""
