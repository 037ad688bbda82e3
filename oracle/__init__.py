"""CPU oracle for the voxelise + MinkUNet hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + torch-CPU) of what the reference's hot path
computes.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.  Nothing under
``generalized-class-discovery-for-lidar-semantic-segmentation_b200/`` imports it: the product
path is the sm_100a C-ABI library and fails loudly when that library is missing.

PARITY PINNING STATUS
  * ``oracle.quantize.ravel_hash / sparse_quantize_np_unique / voxelize_minkunet`` follow
    pure-numpy/torch code that lives in the reference itself (``models/voxelizer.py:271-360``);
    ``tests/golden/make_golden.py`` runs the *reference's own functions* (extracted from the
    read-only checkout at generation time) and freezes their outputs, so this part is pinned.
  * ``utils/voxelizer.py`` (``get_transformation_matrix``) is imported from the reference by
    the same script and pinned the same way.
  * Everything that MinkowskiEngine computes (floor quantise + first-occurrence order, kernel
    maps, sparse conv, BN wrappers) is **parity unpinned**: MinkowskiEngine (un-vendored,
    unpinned, 0.5.x API) is absent from /root/reference and cannot be installed here, and the
    reference has no tests or golden vectors.  The restatement follows ME's published
    semantics as listed in SURVEY.md section 8(a) rows a6-a14 and is cross-checked against an
    independent dense ``torch.nn.functional.conv3d`` formulation (tests/test_oracle_dense_equiv.py).
"""
