"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

The reference's deepsphere/utils.py and deepsphere/data.py are imported unmodified from
/root/reference.  utils.py imports healpy, matplotlib and a PyGSP fork at module level;
none is installable here, so they are replaced by import shims:
  * healpy  -> oracle.healpix (the scalar restatement of pix2vec / get_all_neighbours),
  * matplotlib.pyplot, pygsp.graphs -> empty stand-ins (never called on the new=False path).
Everything downstream of the two healpy calls (index filtering, float32 distances, kernel
width, Gaussian weights, CSR assembly, normalised Laplacian, ARPACK lmax, rescaling,
nside2indexes, ds_index, the dataset iterators) is therefore the reference's code.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from oracle import healpix as ohp  # noqa: E402


def _install_shims():
    hp = types.ModuleType('healpy')
    hp.pix2vec = ohp.pix2vec
    hp.nside2npix = ohp.nside2npix
    hp.npix2nside = lambda npix: int(round((npix / 12) ** 0.5))
    hp.pixelfunc = types.ModuleType('healpy.pixelfunc')
    hp.pixelfunc.get_all_neighbours = ohp.get_all_neighbours
    sys.modules['healpy'] = hp
    sys.modules['healpy.pixelfunc'] = hp.pixelfunc
    mpl = types.ModuleType('matplotlib')
    mpl.pyplot = types.ModuleType('matplotlib.pyplot')
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = mpl.pyplot
    pg = types.ModuleType('pygsp')
    pg.graphs = types.ModuleType('pygsp.graphs')
    for name in ('NNGraph', 'SphereIcosahedron', 'SphereEquiangular', 'SphereHealpix'):
        setattr(pg.graphs, name, type(name, (), {}))
    sys.modules['pygsp'] = pg
    sys.modules['pygsp.graphs'] = pg.graphs
    if 'builtins' not in sys.modules:
        import builtins  # noqa: F401


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _csr_parts(A, prefix):
    A = A.tocsr().copy()
    A.sum_duplicates()
    A.sort_indices()
    return {prefix + '_data': A.data, prefix + '_indices': A.indices.astype(np.int64),
            prefix + '_indptr': A.indptr.astype(np.int64), prefix + '_shape': np.array(A.shape)}


def main():
    _install_shims()
    utils = _load('ref_utils', os.path.join(REF, 'deepsphere', 'utils.py'))
    data = _load('ref_data', os.path.join(REF, 'deepsphere', 'data.py'))

    out = {}
    # full-sphere graphs: utils.healpix_weightmatrix + utils.build_laplacian
    for nside in (1, 2, 4, 8):
        W = utils.healpix_weightmatrix(nside=nside, nest=True, indexes=None, dtype=np.float32)
        L = utils.build_laplacian(W, lap_type='normalized')
        Lc = utils.build_laplacian(W, lap_type='combinatorial')
        out.update(_csr_parts(W, 'full%d_W' % nside))
        out.update(_csr_parts(L, 'full%d_Lnorm' % nside))
        out.update(_csr_parts(Lc, 'full%d_Lcomb' % nside))
    # partial sphere, consecutive indexes (fast path, the cosmo configuration at small size)
    nsides = [32, 16, 8, 4, 4]
    idx = utils.nside2indexes(nsides, 2)
    for i, n in enumerate(nsides):
        out['n2i_order2_%d' % i] = idx[i]
    idx1 = utils.nside2indexes(nsides, 1)
    idx0 = utils.nside2indexes(nsides, 0)
    out['n2i_order1_0'] = idx1[0]
    out['n2i_order0_4'] = idx0[4]
    W = utils.healpix_weightmatrix(nside=32, nest=True, indexes=idx[0], dtype=np.float32)
    out.update(_csr_parts(W, 'part32o2_W'))
    out.update(_csr_parts(utils.build_laplacian(W, 'normalized'), 'part32o2_Lnorm'))
    # partial sphere, non-consecutive indexes (slow path)
    sel = np.array([3, 4, 5, 6, 7, 40, 41, 42, 43, 200, 201, 203, 500, 501, 502, 503, 767])
    W = utils.healpix_weightmatrix(nside=8, nest=True, indexes=sel, dtype=np.float32)
    out['sel8_indexes'] = sel
    out.update(_csr_parts(W, 'sel8_W'))
    # explicit kernel width
    W = utils.healpix_weightmatrix(nside=4, nest=True, indexes=None, dtype=np.float32, std=0.05)
    out.update(_csr_parts(W, 'full4std_W'))
    # ds_index
    ds = utils.ds_index(np.arange(512, 1024), [16, 8, 4])
    for i, a in enumerate(ds):
        out['ds_index_%d' % i] = np.asarray(a)
    # build_laplacians: whole pipeline incl. ARPACK lmax and rescale (new=False)
    np.random.seed(0)
    Ls, p = utils.build_laplacians(nsides, indexes=idx, new=False)
    out['bl_p'] = np.array(p)
    out['bl_n'] = np.array(len(Ls))
    for i, L in enumerate(Ls):
        out.update(_csr_parts(L, 'bl_L%d' % i))
    Ls, p = utils.build_laplacians([4, 2, 1], new=False)
    out['blfull_p'] = np.array(p)
    for i, L in enumerate(Ls):
        out.update(_csr_parts(L, 'blfull_L%d' % i))
    # rescale_L with a fixed lmax (exact float32 arithmetic check)
    W = utils.healpix_weightmatrix(nside=4, nest=True, dtype=np.float32)
    L = utils.build_laplacian(W, 'normalized')
    Lr = utils.rescale_L(L.copy(), lmax=np.float32(1.93))
    out.update(_csr_parts(Lr, 'rescale4'))
    # 4-neighbour grid graph of a patch (use_4=True, utils.py:484-591)
    for nside, order in ((8, 2), (16, 1), (32, 2)):
        idx4 = utils.nside2indexes([nside], order)[0]
        W = utils.build_matrix_4_neighboors(nside, idx4, nest=True, dtype=np.float32)
        out.update(_csr_parts(W, 'four%d_%d_W' % (nside, order)))
    # equiangular (SOFT) 8-neighbour graph, the only equiangular construction that is fully in the reference's repository
    for bw in (2, 4, 8):
        W = utils.equiangular_weightmatrix(bw=bw, dtype=np.float32)
        out.update(_csr_parts(W, 'equi%d_W' % bw))
        out.update(_csr_parts(utils.build_laplacian(W, lap_type='normalized'), 'equi%d_Lnorm' % bw))
    np.savez_compressed(os.path.join(HERE, 'graph_golden.npz'), **out)

    # dataset protocol (deepsphere/data.py): deterministic (shuffle=False) iterators
    rs = np.random.RandomState(7)
    X = rs.randn(10, 6).astype(np.float64)
    lab = np.arange(10) % 3
    d = {'X': X, 'label': lab}
    ds_plain = data.LabeledDataset(X, lab, shuffle=False)
    it = ds_plain.iter(4)
    for i in range(4):
        bx, bl = next(it)
        d['plain_x%d' % i], d['plain_l%d' % i] = bx, bl
    ds_noise = data.LabeledDatasetWithNoise(X, lab, shuffle=False, start_level=0, end_level=2.0, nit=3)
    it = ds_noise.iter(4)
    for i in range(5):
        bx, bl = next(it)
        d['noise_x%d' % i], d['noise_l%d' % i] = bx, bl
    it1 = data.LabeledDataset(X, lab, shuffle=False).iter(1)
    bx, bl = next(it1)
    d['single_x'], d['single_l'] = bx, np.asarray(bl)
    ax, al = ds_plain.get_all_data()
    d['all_x'], d['all_l'] = ax, al
    np.savez_compressed(os.path.join(HERE, 'data_golden.npz'), **d)
    print('wrote', os.path.join(HERE, 'graph_golden.npz'), 'and data_golden.npz')


if __name__ == '__main__':
    main()
