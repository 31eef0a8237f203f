"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / CPU baseline - never as the thing
measured as "ours" or shipped.  The product path (``wsi_hgnn_b200``) fails loudly when
its CUDA library is missing instead of falling back to this code.

Parity status (see DESIGN.md "Oracle"):
* The reference's arithmetic lives in DGL (un-vendored, un-pinned), nmslib (un-pinned) and
  scipy.  DGL/nmslib are absent from this image and the reference ships no tests or golden
  vectors, so the DGL-primitive semantics restated here are **parity unpinned** at that
  boundary.
* What IS pinned: (1) the reference's own model files (models/HEATNet4.py, HEATNet2.py,
  HGT.py, pooling/*.py) are executed UNMODIFIED on top of a minimal DGL shim
  (tests/dgl_shim.py, dense-adjacency formulation, independent of this code) by
  tools/make_golden.py, and the oracle must reproduce those outputs
  (tests/golden/*.pt); (2) scipy.stats.pearsonr - the very function the reference calls
  (construct_graph/graph_constructor.py:278-280) - pins the edge similarity.
"""
