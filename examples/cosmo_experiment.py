#!/usr/bin/env python3
"""The reference's cosmology driver (experiments/cosmo/experiment_deepsphere.py) switched over to the B200 path.

What changes for a user of the reference is the import block and the two lambdas of `get_params`;
the flow -- `get_params` -> datasets -> `models.deepsphere(**params)` -> `fit` -> classification error -- is the reference's
(experiment_deepsphere.py:24-117, 120-159).  The cosmological maps are private (data/README.md), so the script draws a synthetic
two-class problem of the same shape: patches of 1 / (12 order^2) of the sphere in nested order whose classes differ in their
large-scale power (white noise vs. white noise + a field that is constant over blocks of 64 pixels), standardised and trained
under the reference's additive-noise ramp (`LabeledDatasetWithNoise`, or `--device-data` for the HBM-resident pipeline).

    python examples/cosmo_experiment.py                      # Nside 128, order 2: seconds on one B200
    python examples/cosmo_experiment.py FCN --nside 1024     # the size of the reference's experiment
    torchrun --nproc-per-node 2 examples/cosmo_experiment.py # data parallel: every rank feeds its own shard
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from deepsphere_tf1_b200 import models, utils                                # noqa: E402  (was: from deepsphere import ...)
from deepsphere_tf1_b200 import optimizers as train                          # noqa: E402  (was: import tensorflow as tf)
from deepsphere_tf1_b200.data import DeviceDataset, LabeledDataset, LabeledDatasetWithNoise  # noqa: E402


def get_params(ntrain, EXP_NAME, order, Nside, architecture='FCN', num_epochs=80, verbose=True):
    """Hyper-parameters of the cosmology experiment (experiment_deepsphere.py:24-117); 'FCN' and 'CNN' architectures (the 2-D
    CNN baselines of the reference are a different model family)."""
    n_classes = 2
    params = dict(dir_name=EXP_NAME, conv='chebyshev5', pool='max', activation='relu', statistics='mean')
    params['F'] = [16, 32, 64, 64, 64, n_classes]
    params['K'] = [5] * 6
    params['batch_norm'] = [True] * 6
    params['M'] = []
    nsides = [Nside, Nside // 2, Nside // 4, Nside // 8, Nside // 16, Nside // 32, Nside // 32]
    params['nsides'] = nsides
    params['indexes'] = utils.nside2indexes(nsides, order)
    if architecture == 'CNN':                       # fully connected classifier on the same feature extractor
        for key in ('F', 'K', 'batch_norm', 'nsides', 'indexes'):
            params[key] = params[key][:-1]
        params['statistics'] = None
        params['M'] = [n_classes]
    elif architecture != 'FCN':
        raise ValueError('Unknown architecture {}.'.format(architecture))
    params['regularization'] = 0
    params['dropout'] = 1
    params['num_epochs'] = num_epochs
    params['batch_size'] = max(8 * order, 1)
    # the two lines that change: plain Python callables instead of lambdas that return TensorFlow objects
    params['scheduler'] = lambda step: train.exponential_decay(2e-4, step, decay_steps=1, decay_rate=0.999)
    params['optimizer'] = lambda lr: train.AdamOptimizer(lr, beta1=0.9, beta2=0.999, epsilon=1e-8)
    n_evaluations = min(80, num_epochs)
    params['eval_frequency'] = max(1, int(params['num_epochs'] * ntrain / params['batch_size'] / n_evaluations))
    if verbose:
        print('#sides: {}'.format(nsides))
        print('#samples per batch: {}'.format(params['batch_size']))
        n_steps = params['num_epochs'] * ntrain // params['batch_size']
        print('Learning rate will start at {:.1e} and finish at {:.1e}.'.format(params['scheduler'](0), params['scheduler'](n_steps)))
    return params


def synthetic_maps(n, npix, seed):
    """n patches of npix nested pixels, labels in {0, 1}: class 1 carries extra power at the scale of 64-pixel blocks."""
    rs = np.random.RandomState(seed)
    labels = rs.randint(0, 2, size=n)
    x = rs.randn(n, npix).astype(np.float32)
    block = min(64, npix)
    coarse = rs.randn(n, npix // block).astype(np.float32)
    x += 0.6 * labels[:, None].astype(np.float32) * np.repeat(coarse, block, axis=1)
    return x, labels


def classification_error(model, x, labels):
    """experiment_helper.model_error: the share of wrong predictions of the trained model."""
    pred = model.predict(x[:, :, np.newaxis], sess=model)
    return float(np.mean(pred != labels))


def single_experiment(order, sigma_noise, experiment_type, new, n_neighbors, Nside=128, ntrain=256, ntest=64, num_epochs=8,
                      device_data=False, verbose=True):
    EXP_NAME = ('cosmo' if new else 'oldgraph') + '_{}sides_{}noise_{}order_{}neighbor_{}_synthetic'.format(
        Nside, sigma_noise, order, n_neighbors, experiment_type)
    npix = 12 * Nside ** 2 // (12 * order ** 2) if order else 12 * Nside ** 2
    rank = int(os.environ.get('RANK', '0'))
    x_train, labels_train = synthetic_maps(ntrain, npix, 1 + 10 * rank)          # under torchrun every rank draws its own shard
    x_val, labels_val = synthetic_maps(ntest, npix, 2)
    x_test, labels_test = synthetic_maps(ntest, npix, 3)
    x_std = x_train.std()
    x_train, x_val, x_test = x_train / x_std, x_val / x_std, x_test / x_std
    rs = np.random.RandomState(4)
    x_val = x_val + sigma_noise * rs.randn(*x_val.shape).astype(np.float32)      # validation / test at the final noise level
    x_test = x_test + sigma_noise * rs.randn(*x_test.shape).astype(np.float32)

    params = get_params(ntrain, EXP_NAME, order, Nside, experiment_type, num_epochs, verbose)
    if 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        if not dist.is_initialized():
            dist.init_process_group('nccl')
    model = models.deepsphere(**params, new=new, n_neighbors=n_neighbors, verbose=verbose)
    if device_data:       # samples resident in HBM, batches and noise made by one kernel (data.DeviceDataset)
        training = DeviceDataset(x_train, labels_train, end_level=sigma_noise, nit=0)
    else:
        training = LabeledDatasetWithNoise(x_train, labels_train, end_level=sigma_noise)
    validation = LabeledDataset(x_val, labels_val)

    accuracy_validation, loss_validation, loss_training, t_step, t_batch = model.fit(training, validation, restore=False,
                                                                                     verbose=verbose)
    print('inference time: ', t_batch / params['batch_size'])
    error_validation = classification_error(model, x_val, labels_val)
    print('The validation error is {}%'.format(error_validation * 100), flush=True)
    error_test = classification_error(model, x_test, labels_test)
    print('The testing error is {}%'.format(error_test * 100), flush=True)
    return error_test, t_batch


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('experiment_type', nargs='?', default='FCN', choices=['FCN', 'CNN'])
    ap.add_argument('--order', type=int, default=2)
    ap.add_argument('--sigma-noise', type=float, default=0.5)
    ap.add_argument('--new', action='store_true', help='the kNN graph of the PyGSP fork instead of the in-repo 8-neighbour graph')
    ap.add_argument('--n-neighbors', type=int, default=8)
    ap.add_argument('--nside', type=int, default=128)
    ap.add_argument('--ntrain', type=int, default=256)
    ap.add_argument('--epochs', type=int, default=8)
    ap.add_argument('--device-data', action='store_true')
    a = ap.parse_args()
    np.random.seed(42)
    single_experiment(a.order, a.sigma_noise, a.experiment_type, a.new, a.n_neighbors, a.nside, a.ntrain, num_epochs=a.epochs,
                      device_data=a.device_data)
