"""CPU oracle for the GeoSplatting splat + shade hot path.

TEST INFRASTRUCTURE ONLY: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The product package
``geosplatting_b200`` never does (tests/test_boundary.py greps for it).
"""
