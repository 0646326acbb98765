"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's free-energy hot path plus the harness that
runs the unmodified reference from /root/reference.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import anything from here; the product package
``meanfieldvargp_b200`` never does.
"""
