"""CPU oracle for the SPH/WVT hot path -- TEST INFRASTRUCTURE ONLY.

``oracle.ref``  : the reference's own sources compiled unmodified (oracle/_ref, git-ignored).
``oracle.port`` : a plain-C restatement of the same algorithm (oracle/toy_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``toycluster_b200``) never does.
"""
