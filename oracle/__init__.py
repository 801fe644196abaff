"""CPU oracle for the cluster-ICP hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker / the CPU arm.

PARITY UNPINNED at the third-party boundary (open3d 0.18.0 / pytorch3d 0.7.7 are
neither vendored in the reference nor installable here, and the reference holds no
golden vectors); see ``icp_oracle.c`` and DESIGN.md for what is pinned instead.
"""
