"""Dataset protocol vs fixtures produced by the reference's own deepsphere/data.py."""
import os

import numpy as np

from conftest import GOLDEN
from deepsphere_tf1_b200 import data

D = np.load(os.path.join(GOLDEN, 'data_golden.npz'))


def test_plain_batches_wrap_like_the_reference():
    ds = data.LabeledDataset(D['X'], D['label'], shuffle=False)
    assert ds.N == 10
    it = ds.iter(4)
    for i in range(4):
        bx, bl = next(it)
        np.testing.assert_array_equal(bx, D['plain_x%d' % i])
        np.testing.assert_array_equal(bl, D['plain_l%d' % i])
    ax, al = ds.get_all_data()
    np.testing.assert_array_equal(ax, D['all_x'])
    np.testing.assert_array_equal(al, D['all_l'])
    assert ax.dtype == np.float32


def test_single_sample_iteration():
    bx, bl = next(data.LabeledDataset(D['X'], D['label'], shuffle=False).iter(1))
    np.testing.assert_array_equal(bx, D['single_x'])
    assert int(bl) == int(D['single_l'])


def test_noise_ramp_matches_reference():
    ds = data.LabeledDatasetWithNoise(D['X'], D['label'], shuffle=False, start_level=0, end_level=2.0, nit=3)
    it = ds.iter(4)
    for i in range(5):
        bx, bl = next(it)
        np.testing.assert_array_equal(bx, D['noise_x%d' % i])
        np.testing.assert_array_equal(bl, D['noise_l%d' % i])


def test_shuffled_dataset_is_a_permutation():
    np.random.seed(3)
    ds = data.LabeledDataset(D['X'], np.arange(10), shuffle=True)
    it = ds.iter(5)
    seen = np.concatenate([next(it)[1], next(it)[1]])
    assert sorted(seen.tolist()) == list(range(10))
