"""Golden vectors of the TF1 operators, produced by the REFERENCE's own graph-construction code.

Run in the build container only (needs /root/reference):  python tests/golden/make_ops_golden.py

/root/reference/deepsphere/models.py is imported UNMODIFIED; `tensorflow` is replaced by tests/golden/tf_shim.py (an
eager torch-CPU stand-in, see its header for what it decides itself) and healpy / pygsp / matplotlib by the import
shims of make_golden.py.  Each case instantiates the reference's `deepsphere` (or `cgcnn`) class, which runs its
`build_graph` -> `_inference` / `_decoder` / `loss` / `training` on the fed batch, and records: every variable (name,
value), the logits, descriptor, loss, probabilities / prediction, d loss / d variable, and the batch-norm moving
statistics after the step.  tests/test_ops_golden.py checks the oracle (CPU) and the CUDA path (GPU) against them.
"""
import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np
import torch
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden  # noqa: E402
import tf_shim  # noqa: E402


def load_reference():
    make_golden._install_shims()
    tf_shim.install()
    sys.path.insert(0, make_golden.REF)
    warnings.simplefilter('ignore', SyntaxWarning)                  # `is 'mean'` comparisons (models.py:980-989)
    import deepsphere.models as ref_models
    import deepsphere.utils as ref_utils
    return ref_models, ref_utils


def random_graph(n, deg, seed):
    """A symmetric sparse matrix with spectrum roughly in [-1, 1]: stands for a rescaled Laplacian where only the
    operator code is exercised (the icosahedral graphs come from the absent PyGSP fork)."""
    rs = np.random.RandomState(seed)
    rows = np.repeat(np.arange(n), deg)
    cols = rs.randint(0, n, size=n * deg)
    vals = rs.randn(n * deg).astype(np.float32) / (2 * deg)
    A = sparse.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A = (A + A.T) * 0.5
    A.sum_duplicates()
    return sparse.coo_matrix(A, dtype=np.float32)


def run_case(ref_models, name, out, kw, x, labels, training, sphere=None, L=None, p=None, sampling='healpix',
             state_preset=None, var_preset=None, seed=0):
    """Instantiate the reference model (its constructor builds the graph = runs the step in the shim)."""
    tf_shim.reset(seed)
    tf_shim.feed(x, labels, training)
    presets = dict(state_preset or {})
    presets.update(var_preset or {})
    tf_shim.preset(presets)
    common = dict(num_epochs=1, scheduler=lambda step: 1e-3, optimizer=lambda lr: tf_shim.GradientRecorder(lr),
                  batch_size=x.shape[0], dir_name='_golden')
    softmax_in = []
    real_softmax = tf_shim.nn.softmax
    tf_shim.nn.softmax = lambda t: (softmax_in.append(t), real_softmax(t))[1]      # the logits (local in build_graph)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            if sphere is not None:
                model = ref_models.deepsphere(**sphere, **kw, **common)
            else:
                model = ref_models.cgcnn.__new__(ref_models.cgcnn)
                model.sampling = sampling
                model.__init__(L=L, p=p, **kw, **common)
    finally:
        tf_shim.nn.softmax = real_softmax
    rec = {}
    cfg = dict(kw)
    if 'mask' in cfg:
        cfg['mask'] = True                                                       # only the presence of the argument matters
    cfg.update(sampling=model.sampling if hasattr(model, 'sampling') else sampling, p=[int(v) for v in model.p],
               training=bool(training), n_levels=len(model.L), ratio=float(getattr(model, 'ratio', 1.)))
    rec['config'] = np.array(json.dumps(cfg))
    for i, Li in enumerate(model.L):
        Li = Li.tocoo()
        rec['L%d_row' % i], rec['L%d_col' % i] = Li.row.astype(np.int32), Li.col.astype(np.int32)
        rec['L%d_data' % i], rec['L%d_n' % i] = Li.data.astype(np.float32), np.array(Li.shape[0])
    rec['x'], rec['labels'] = np.asarray(x, np.float32), np.asarray(labels)
    rec['var_names'] = np.array(json.dumps(list(tf_shim.variables().keys())))
    for k, v in tf_shim.variables().items():
        rec['var/' + k] = v.detach().numpy()
        g = tf_shim.gradients().get(k)
        rec['grad/' + k] = (g if g is not None else torch.zeros_like(v)).detach().numpy()
    for k, v in tf_shim.state().items():
        if 'global_step' in k:
            continue
        rec['state/' + k] = v.detach().numpy()
        if k in tf_shim.updates():
            rec['update/' + k] = tf_shim.updates()[k].detach().numpy()
    rec['logits'] = softmax_in[-1].detach().numpy()
    rec['descriptor'] = model.op_descriptor.detach().numpy()
    rec['loss'] = np.array(float(model.op_loss))
    rec['probabilities'] = model.op_probabilities.detach().numpy()
    rec['prediction'] = model.op_prediction.detach().numpy()
    for i, m in enumerate(tf_shim.dropout_masks()):
        rec['dropout_mask%d' % i] = m.numpy()
    rec['n_dropout_masks'] = np.array(len(tf_shim.dropout_masks()))
    for k, v in rec.items():
        out['{}::{}'.format(name, k)] = v
    print('{:16s} loss {:.6f}  variables {}'.format(name, float(model.op_loss), list(tf_shim.variables().keys())))
    return model


def main():
    ref_models, ref_utils = load_reference()
    out = {}
    rs = np.random.RandomState(11)

    def moving(scopes_F):
        st = {}
        for sc, F in scopes_F:
            st[sc + '/batch_normalization/moving_mean'] = rs.randn(F).astype(np.float32) * 0.3
            st[sc + '/batch_normalization/moving_variance'] = (0.5 + rs.rand(F)).astype(np.float32)
        return st

    # A. the cosmo stack in miniature: partial sphere, K=5, BN, max pool, mean statistic, linear last conv
    nsides = [8, 4, 2, 2]
    sphere = dict(nsides=nsides, indexes=ref_utils.nside2indexes(nsides, 1), new=False)
    kw = dict(F=[4, 8, 2], K=[5, 5, 5], batch_norm=[True] * 3, M=[], statistics='mean')
    x = rs.randn(3, 64, 1).astype(np.float32)
    run_case(ref_models, 'cosmo_mini', out, kw, x, rs.randint(0, 2, 3), True, sphere=sphere,
             state_preset=moving([('conv1', 4), ('conv2', 8)]))
    run_case(ref_models, 'cosmo_mini_eval', out, kw, x, rs.randint(0, 2, 3), False, sphere=sphere,
             state_preset=moving([('conv1', 4), ('conv2', 8)]))

    # B. full sphere, average pooling, elu, fully connected head, L2 regularisation
    sphere = dict(nsides=[4, 2, 1], new=False)
    kw = dict(F=[3, 5], K=[3, 4], batch_norm=[True, False], M=[7, 4], pool='average', activation='elu',
              regularization=0.1, num_feat_in=2)
    x = rs.randn(2, 192, 2).astype(np.float32)
    run_case(ref_models, 'fc_head', out, kw, x, rs.randint(0, 4, 2), True, sphere=sphere)

    # C. statistics variants
    for stat in ('var', 'meanvar', 'max', 'histogram'):
        sphere = dict(nsides=[2, 1, 1], new=False)
        kw = dict(F=[3, 4], K=[2, 3], batch_norm=[False, True], M=[3], statistics=stat, regularization=0.01)
        x = rs.randn(2, 48, 1).astype(np.float32)
        run_case(ref_models, 'stat_' + stat, out, kw, x, rs.randint(0, 3, 2), True, sphere=sphere)

    # D. monomial filters with regularisation (the default stddev 0.1 of _weight_variable enters init AND the regulariser)
    sphere = dict(nsides=[4, 2, 2], new=False)
    kw = dict(F=[4, 3], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean', conv='monomials', regularization=0.5)
    x = rs.randn(2, 192, 1).astype(np.float32)
    run_case(ref_models, 'monomials', out, kw, x, rs.randint(0, 3, 2), True, sphere=sphere)

    # E. regression with the mean statistic
    sphere = dict(nsides=[4, 2, 2], new=False)
    kw = dict(F=[4, 1], K=[4, 4], batch_norm=[True, True], M=[], statistics='mean', regression=True)
    x = rs.randn(3, 192, 1).astype(np.float32)
    run_case(ref_models, 'regression', out, kw, x, rs.randn(3).astype(np.float32), True, sphere=sphere)

    # F. dense (segmentation) HEALPix encoder-decoder: max pool -> zero-fill unpool; average pool -> repeat unpool
    for pool, weighted in (('max', True), ('average', False)):
        sphere = dict(nsides=[4, 2, 1, 1], new=False)
        kw = dict(F=[4, 6, 5], K=[3, 3, 2], batch_norm=[True, True, False], M=[], pool=pool, dense=True, Fseg=3,
                  num_feat_in=2, weighted=weighted)
        x = rs.randn(2, 192, 2).astype(np.float32)
        run_case(ref_models, 'dense_' + pool, out, kw, x, rs.randint(0, 3, (2, 192)), True, sphere=sphere)

    # G. icosahedral conventions (prefix-slice pooling, pad-with-ones unpooling up to the hard-coded 10,242 vertices,
    #    models.py:1137-1138, 1297-1301, 1356-1358) on stand-in graphs of the level sizes 10242 / 10242 / 2562
    L5 = random_graph(10242, 3, 1)
    Ls = [L5, L5.copy(), random_graph(2562, 3, 2)]
    kw = dict(F=[3, 4, 4], K=[2, 2, 2], batch_norm=[True, False, True], M=[], pool='average', dense=True, Fseg=3,
              num_feat_in=2)
    x = rs.randn(1, 10242, 2).astype(np.float32)
    run_case(ref_models, 'ico_dense', out, kw, x, rs.randint(0, 3, (1, 10242)), True, L=Ls, p=[10242, 2562, 2562],
             sampling='icosahedron')

    # H. filter dropout and fully connected dropout (masks recorded in the order the graph draws them)
    sphere = dict(nsides=[2, 1, 1], new=False)
    kw = dict(F=[4, 4], K=[3, 3], batch_norm=[True, True], M=[6, 2], dropFilt=0.6, dropout=0.5)
    x = rs.randn(4, 48, 1).astype(np.float32)
    run_case(ref_models, 'dropouts', out, kw, x, rs.randint(0, 2, 4), True, sphere=sphere)

    # I. equiangular conventions: 2-D pooling of the (colatitude, longitude) grid with a sqrt(p) x sqrt(p) window and the
    #    repeat_elements unpooling (models.py:1131-1136, 1150-1155, 1348-1355), on the reference's own in-repository equiangular
    #    graph (utils.equiangular_weightmatrix + build_laplacian; the fork's SphereEquiangular is absent)
    def equi_L(bw):
        Lq = ref_utils.build_laplacian(ref_utils.equiangular_weightmatrix(bw=bw, dtype=np.float32), 'normalized')
        return sparse.coo_matrix(ref_utils.rescale_L(sparse.csr_matrix(Lq), lmax=np.float32(2.0)), dtype=np.float32)
    Ls = [equi_L(8), equi_L(4), equi_L(2)]
    kw = dict(F=[4, 6, 3], K=[3, 3, 2], batch_norm=[True, True, False], M=[], pool='max', statistics='mean', num_feat_in=2)
    x = rs.randn(2, 256, 2).astype(np.float32)

    # (run_case builds the model through cgcnn.__new__: give the class the attribute deepsphere.__init__ would set)
    ref_models.cgcnn.ratio = 1.
    run_case(ref_models, 'equi_max', out, kw, x, rs.randint(0, 3, 2), True, L=Ls, p=[4, 4, 1], sampling='equiangular')
    kw = dict(F=[4, 5, 3], K=[3, 2, 2], batch_norm=[True, False, True], M=[], pool='average', dense=True, Fseg=3, num_feat_in=2)
    Ls = [equi_L(8), equi_L(4), equi_L(4)]
    run_case(ref_models, 'equi_dense', out, kw, x, rs.randint(0, 3, (2, 256)), True, L=Ls, p=[4, 1, 1], sampling='equiangular')

    # J. rectangular equiangular grids (`ratio` = columns / rows, set by deepsphere.__init__ from a pair of bandwidths,
    #    models.py:1832-1836; the reshape of models.py:1134, 1153, 1353): an 8 x 12 grid (ratio 1.5, the aspect of the climate
    #    data's 768 x 1152 grid) pooled twice, and its encoder-decoder.  Only the operator code is exercised: random symmetric
    #    matrices stand for the Laplacians (the rectangular graph comes from the absent fork's SphereEquiangular)
    ref_models.cgcnn.ratio = 1.5
    Ls = [random_graph(96, 4, 31), random_graph(24, 4, 32), random_graph(6, 3, 33)]
    kw = dict(F=[4, 6, 3], K=[3, 3, 2], batch_norm=[True, True, False], M=[], pool='max', statistics='mean', num_feat_in=2)
    x = rs.randn(3, 96, 2).astype(np.float32)
    run_case(ref_models, 'equi_rect_max', out, kw, x, rs.randint(0, 3, 3), True, L=Ls, p=[4, 4, 1], sampling='equiangular')
    kw = dict(F=[4, 5, 3], K=[3, 2, 2], batch_norm=[True, False, True], M=[], pool='average', dense=True, Fseg=3, num_feat_in=2)
    Ls = [random_graph(96, 4, 34), random_graph(24, 4, 35), random_graph(24, 4, 36)]
    run_case(ref_models, 'equi_rect_dense', out, kw, x, rs.randint(0, 3, (3, 96)), True, L=Ls, p=[4, 1, 1], sampling='equiangular')
    ref_models.cgcnn.ratio = 1.

    # K. the GHCN stacks (experiments/ghcn/GHCN_train.py:7-60): the same irregular graph at every layer, no pooling (p = 1), the
    #    combinatorial Laplacian passed UN-rescaled, dense regression / global regression through a fully connected output,
    #    and the masked dense loss (`mask=[train, val]`: only its presence matters, models.py:785-787, 1030-1032 -- the last two
    #    input channels are the masks of the predictions and of the labels)
    def knn_like(n, deg, seed):
        g = np.random.RandomState(seed)
        rows = np.repeat(np.arange(n), deg)
        cols = (rows + 1 + g.randint(0, n - 1, size=n * deg)) % n             # no self loops
        W = sparse.coo_matrix((g.rand(n * deg).astype(np.float32), (rows, cols)), shape=(n, n)).tocsr()
        W = W.maximum(W.T)
        d = np.asarray(W.sum(axis=1)).ravel()
        return sparse.coo_matrix(sparse.diags(d) - W, dtype=np.float32)        # L = D - W, spectrum in [0, 2 max d]
    Lg = knn_like(60, 3, 41)
    Lg = sparse.coo_matrix(Lg / np.float32(4.0), dtype=np.float32)             # keep T_k(L) x of order one without rescale_L
    kw = dict(F=[5, 6, 6, 1], K=[3, 3, 3, 3], batch_norm=[True] * 4, M=[], statistics=None, regression=True, dense=True,
              num_feat_in=3)
    x = rs.randn(4, 60, 3).astype(np.float32)
    run_case(ref_models, 'ghcn_dense', out, kw, x, rs.randn(4, 60).astype(np.float32), True, L=[Lg] * 4, p=[1, 1, 1, 1])
    kw = dict(F=[5, 6, 6], K=[3, 3, 3], batch_norm=[True] * 3, M=[1], statistics='mean', regression=True, num_feat_in=3)
    run_case(ref_models, 'ghcn_global', out, kw, x, rs.randn(4).astype(np.float32), True, L=[Lg] * 3, p=[1, 1, 1])
    kw = dict(F=[5, 6, 1], K=[3, 3, 3], batch_norm=[True] * 3, M=[], statistics=None, regression=True, dense=True,
              num_feat_in=5, mask=[np.ones(60), np.ones(60)])
    xm = rs.randn(4, 60, 5).astype(np.float32)
    xm[..., -2:] = (rs.rand(4, 60, 2) < 0.7).astype(np.float32)
    run_case(ref_models, 'ghcn_masked', out, kw, xm, rs.randn(4, 60).astype(np.float32), True, L=[Lg] * 3, p=[1, 1, 1])

    # L. one layer pooling by 16 (Nside 4 -> 1)
    sphere = dict(nsides=[4, 1, 1], new=False)
    kw = dict(F=[3, 4], K=[3, 2], batch_norm=[True, True], M=[], statistics='mean', pool='max')
    x = rs.randn(2, 192, 1).astype(np.float32)
    run_case(ref_models, 'pool16', out, kw, x, rs.randint(0, 4, 2), True, sphere=sphere)

    # Cases that go through the reference's build_laplacians carry an ARPACK lmax (random start vector, utils.py:423): a
    # regeneration moves them by ~1e-7.  Keep the committed vectors of the cases that already exist and only add new ones
    # (--all rewrites everything).
    path = os.path.join(HERE, 'ops_golden.npz')
    if os.path.exists(path) and '--all' not in sys.argv:
        old = np.load(path, allow_pickle=False)
        have = {k.split('::')[0] for k in old.files}
        out = {k: v for k, v in out.items() if k.split('::')[0] not in have}
        print('kept', sorted(have), 'added', sorted({k.split('::')[0] for k in out}))
        out.update({k: old[k] for k in old.files})
    np.savez_compressed(os.path.join(HERE, 'ops_golden.npz'), **out)
    print('wrote', os.path.join(HERE, 'ops_golden.npz'), os.path.getsize(os.path.join(HERE, 'ops_golden.npz')), 'bytes')


if __name__ == '__main__':
    main()
