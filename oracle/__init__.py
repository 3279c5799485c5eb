"""CPU oracle for BUSCA's per-frame association hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It restates, on the CPU, the algorithm of the reference
(lorenzovaquero/BUSCA @ 88a9ed7e) for the path in SURVEY.md section 8(a); every function cites
the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
or the timed CPU baseline - never as part of the product path (``busca_b200/`` fails loudly when
its CUDA library is missing).

Pinning: the reference is Python, so it was imported in the build container (with the shims under
tests/golden/shims) and its outputs on seeded inputs were committed as tests/golden/*.npz by
tests/golden/make_golden.py; tests/test_oracle_*.py check this package against those fixtures.
Third-party arithmetic not present under /root/reference:
  * positional_encodings==6.0.3  - PARITY UNPINNED (restated from the published algorithm)
  * cython_bbox.bbox_overlaps    - pinned on the reference's in-tree restatement
                                   (adapters/GHOST/src/tracking_utils.py:176-205)
  * cv2.resize (INTER_LINEAR, 8-bit) - pinned against cv2 itself through the reference's
                                   get_bbox_crop (fixtures crops.npz) and live in tests.

Integer/byte/index work is numpy (bit-exact target); the floating-point network (ReID ResNet-50,
Decision Transformer) is a functional fp32 PyTorch-CPU restatement driven by the state dict.
"""
