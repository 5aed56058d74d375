"""TEST INFRASTRUCTURE ONLY: the CPU oracle for the FIRECODE embedding screen.

* ``oracle.prism_pruner``  numpy restatement of the absent third-party dependency (PARITY UNPINNED).
* ``oracle.loader``        runs the UNMODIFIED reference from /root/reference (this container only).
* ``oracle.port``          numpy restatement of the in-tree hot path on plain arrays; it is pinned
                           against ``oracle.loader`` runs here and travels to the GPU box.
* ``oracle/c``             plain-C restatement of the clash test used as the CPU baseline.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import it.
"""
