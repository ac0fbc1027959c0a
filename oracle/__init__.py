"""CPU oracle for the DeepSphere ChebConv hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / scipy / torch-CPU autograd) of the algorithms in the reference
`deepsphere/models.py` and `deepsphere/utils.py`.  Every function cites the reference
file:line it follows.

It may be imported ONLY by `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`.  The product package `deepsphere_tf1_b200` never
imports it and has no CPU fallback: if the CUDA library is missing the product raises.

Parity pinning status (see DESIGN.md section "Oracle"):
  * graph construction (`healpix_weightmatrix`, `build_laplacian`, `rescale_L`,
    `build_laplacians`, `nside2indexes`): PINNED against the reference's own
    `deepsphere/utils.py`, executed in the build container through import shims
    (tests/golden/make_golden.py) -- fixtures committed under tests/golden/.
  * HEALPix nested geometry (healpy.pix2vec / get_all_neighbours; healpy 1.11.0 is a
    third-party dependency absent from /root/reference): restated from the published
    HEALPix algorithm; pinned only by known answers from outside this repository (both
    `get_all_neighbours` docstring examples and the `pix2ang` docstring angles of healpy,
    base-pixel centres and Nside-2 children from the ring formulas of Gorski et al. 2005:
    tests/test_graph_golden.py) and geometric invariants -> "parity unpinned" for the rest
    of the absolute pixel coordinates.
  * TF1 operators and models (chebyshev5, monomials, batch norm, pooling / unpooling, statistics,
    fc, `_inference`, `_decoder`, losses): PINNED to the reference's OWN graph-construction code.
    tests/golden/make_ops_golden.py imports the UNMODIFIED deepsphere/models.py with `tensorflow`
    replaced by an eager torch-CPU stand-in (tests/golden/tf_shim.py) and commits variables, logits,
    losses, every gradient and the batch-norm updates of 21 model cases (tests/golden/ops_golden.npz);
    tests/test_ops_golden.py holds this oracle to 2e-5 and the CUDA path to 1e-4 of them.  What the
    stand-in decides itself stays unpinned: the arithmetic INSIDE a TF op (summation order, the
    moments of tf.layers.batch_normalization), the random streams, and the optimizer update rules
    (oracle/optim_tf1.py: restated from the TF 1.10 sources, "parity unpinned").
  * icosahedral mesh: PINNED bit for bit to the `SphereIcosahedron` class source of
    notebooks/sphere_to_graph.ipynb (tests/golden/make_ico_golden.py).
"""
