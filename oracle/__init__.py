"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Routeformer hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / CPU baseline.  The product
package (``routeformer_b200``) never imports from here and fails loudly when its
CUDA library is missing.

Parity status: PINNED.  ``oracle/routeformer_oracle.py`` is checked (a) directly
against the unmodified reference modules imported from ``/root/reference`` through
``oracle/reference_shim.py`` (tests/test_oracle_vs_reference.py, runs only where the
reference is mounted) and (b) against golden vectors generated from that reference
by ``oracle/make_golden.py`` and committed under ``tests/golden/``.
``oracle/area_resize.py`` (the loader's cv2.resize INTER_AREA scaling) restates
OpenCV's published algorithm and is pinned bit-exactly against cv2 itself and against
``tests/golden/area_resize.npz`` (generated with cv2 by ``oracle/make_golden_resize.py``).
The visual backbone (timm SwinV2, un-vendored, needs downloaded weights) is the one
boundary that is "parity unpinned": it is replaced by a build-defined patch-embed
encoder whose oracle is the plain-PyTorch twin in this package.
"""
