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
    HEALPix algorithm; pinned only by the healpy docstring example and geometric
    invariants -> "parity unpinned" for the absolute pixel coordinates.
  * TF1 operators (chebyshev5, batch-norm, pooling, statistics, loss, Adam): the reference
    has no tests, golden vectors or fixtures and TensorFlow 1.10 cannot run here ->
    "parity unpinned"; anchored on line-by-line correspondence with models.py, mathematical
    identities (Chebyshev polynomials on eigenvectors, float64 finite differences) and the
    known answers printed in experiments/cosmo/demo.ipynb (parameter counts, level sizes,
    learning-rate end points, initial loss ln 2).
"""
