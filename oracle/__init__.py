"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's preshape hot path
(``embodiedscan/models/necks/preshape_norm_reverse_drop.py`` in
pqh22/ProxyTransformation), used as the checker for the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import, call, link or execute anything in this
directory.  Nothing under ``proxytransformation_b200/`` imports it, and the
product path raises if its CUDA library is missing rather than falling back.

Pinning status: the reference has NO tests or golden vectors for this path
(SURVEY.md §4), so the restatement is pinned against the reference ITSELF,
imported unmodified in the build container under three import shims
(``oracle/ref_shim.py``); the outputs are committed as fixtures under
``tests/golden/`` together with the generating script
(``tests/golden/make_golden.py``).  The pytorch3d boundary (ball query, FPS) is
un-vendored and unpinned upstream; its published CPU algorithm is restated in
``oracle/geom.c`` and cross-checked against the reference's in-tree copies
(``sample_farthest_points_naive`` :527-625, ``masked_gather`` :627-672).
"""
