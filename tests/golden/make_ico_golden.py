"""Golden icosahedral meshes, produced by the REFERENCE's own class source.

Run in the build container only (needs /root/reference):  python tests/golden/make_ico_golden.py

The PyGSP fork that ships `SphereIcosahedron` is absent, but /root/reference/notebooks/sphere_to_graph.ipynb holds the class
as it was developed (and its helpers unique_rows / hashable_rows / float_to_int / decimal_to_digits).  The two notebook
cells are executed UNMODIFIED with `NNGraph` replaced by a stub that only records the constructor arguments (the kNN graph
itself is PyGSP code), and the vertices / faces / neighbour count of levels 0-4 are stored in ico_golden.npz.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NB = '/root/reference/notebooks/sphere_to_graph.ipynb'


class NNGraph(object):
    def __init__(self, coords, k=10, **kwargs):
        self.nn_coords, self.nn_k, self.nn_kwargs = np.array(coords), k, kwargs


def main():
    nb = json.load(open(NB))
    ns = {'np': np, 'NNGraph': NNGraph}
    for cell in nb['cells']:
        src = ''.join(cell['source'])
        if cell['cell_type'] == 'code' and ('def unique_rows(' in src or 'class SphereIcosahedron(NNGraph)' in src):
            exec(compile(src, NB, 'exec'), ns)
    out = {}
    for level in range(5):
        g = ns['SphereIcosahedron'](level)
        assert g.nn_coords.shape == (10 * 4 ** level + 2, 3)
        out['coords%d' % level] = g.nn_coords
        out['faces%d' % level] = np.asarray(g.faces)
        out['k%d' % level] = np.array(g.nn_k)
        out['kwargs%d' % level] = np.array(json.dumps(sorted(g.nn_kwargs)))
    np.savez_compressed(os.path.join(HERE, 'ico_golden.npz'), **out)
    print('wrote ico_golden.npz', {k: v.shape for k, v in out.items() if k.startswith('coords')})


if __name__ == '__main__':
    main()
