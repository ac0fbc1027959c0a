"""Device-side input pipeline (deepsphere_tf1_b200/data.py: DeviceDataset, csrc/input.cu) against the host classes that restate
the reference's deepsphere/data.py: same sample order and labels, same noise-level ramp; the noise values come from another
generator (Philox on the device instead of numpy's Mersenne Twister), so they are compared as a distribution."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(n=24, m=300, seed=0):
    rs = np.random.RandomState(seed)
    return rs.randn(n, m).astype(np.float32), rs.randint(0, 5, size=n)


@pytest.mark.parametrize('shuffle', [False, True])
@pytest.mark.parametrize('bs', [1, 8, 7])
def test_batches_equal_the_host_dataset_without_noise(dev, shuffle, bs):
    from deepsphere_tf1_b200.data import DeviceDataset, LabeledDataset
    X, lab = _data()
    # the permutations are drawn from numpy's global generator in the constructor and at the first batch: same seed, same draws
    np.random.seed(3)
    host = LabeledDataset(X, lab, shuffle=shuffle)
    it_h = host.iter(bs)
    first_h = next(it_h)
    np.random.seed(3)
    devds = DeviceDataset(X, lab, shuffle=shuffle, device=dev)
    it_d = devds.iter(bs)
    first_d = next(it_d)
    assert devds.N == host.N
    for i in range(9):                                       # wraps around the end of the data
        xh, lh = first_h if i == 0 else next(it_h)
        xd, ld = first_d if i == 0 else next(it_d)
        assert xd.is_cuda and tuple(xd.shape) == tuple(np.shape(xh))
        np.testing.assert_array_equal(xd.cpu().numpy(), xh)
        np.testing.assert_array_equal(ld.cpu().numpy(), lh)


def test_noise_ramp_distribution_and_reproducibility(dev):
    from deepsphere_tf1_b200.data import DeviceDataset, LabeledDatasetWithNoise
    X = np.zeros((16, 20000), np.float32)
    lab = np.arange(16)
    kw = dict(shuffle=False, start_level=0.5, end_level=2.0, nit=4)
    host = LabeledDatasetWithNoise(X, lab, **kw)
    a = DeviceDataset(X, lab, device=dev, seed=11, **kw)
    it = a.iter(8)
    batches = [next(it)[0] for _ in range(6)]
    for i, b in enumerate(batches):
        level = host.noise_level(i)                           # the ramp of data.py:136-139
        v = b.double()
        assert abs(float(v.std()) - level) < 0.02 * level and abs(float(v.mean())) < 0.02 * level
        # Gaussian: excess kurtosis ~ 0, no correlation between neighbouring elements or between samples
        z = v / level
        assert abs(float((z ** 4).mean()) - 3.0) < 0.1
        assert abs(float((z[:, 1:] * z[:, :-1]).mean())) < 0.01 and abs(float((z[0] * z[1]).mean())) < 0.03
    assert not torch.equal(batches[4], batches[5])            # different calls draw different numbers
    b2 = DeviceDataset(X, lab, device=dev, seed=11, **kw).iter(8)
    assert torch.equal(next(b2)[0], batches[0])               # same seed, same call: same numbers
    c = DeviceDataset(X, lab, device=dev, seed=12, **kw).iter(8)
    assert not torch.equal(next(c)[0], batches[0])


def test_all_level_option_scales_the_last_quarter(dev):
    from deepsphere_tf1_b200.data import DeviceDataset
    X = np.zeros((8, 50000), np.float32)
    ds = DeviceDataset(X, np.zeros(8, int), shuffle=False, start_level=1.0, end_level=1.0, all_level=True, device=dev, seed=2)
    x, _ = next(ds.iter(8))
    std = x.double().std(dim=1).cpu().numpy()
    np.testing.assert_allclose(std[:6], 1.0, atol=0.02)       # GaussianNoise(all_level=True): the first 3/4 at full level
    assert np.all(std[6:] < 1.0 + 0.02) and np.all(std[6:] >= 0)


def test_fit_consumes_device_batches(dev):
    from deepsphere_tf1_b200 import models, optimizers
    from deepsphere_tf1_b200.data import DeviceDataset, LabeledDataset
    rs = np.random.RandomState(0)
    labels = rs.randint(0, 2, size=40)
    x = rs.randn(40, 192).astype(np.float32) + labels[:, None] * 1.5
    train = DeviceDataset(x[:32], labels[:32], start_level=0.1, end_level=0.1, device=dev)
    val = LabeledDataset(x[32:], labels[32:], shuffle=False)
    model = models.deepsphere(nsides=[4, 2, 2], new=False, F=[8, 2], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean',
                              num_epochs=12, batch_size=8, eval_frequency=16, seed=1, verbose=False,
                              scheduler=lambda step: 2e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr),
                              dir_name='_test_fit_device')
    acc, loss_val, loss_train, t_step, t_batch = model.fit(train, val, restore=False, verbose=False)
    assert len(acc) == 3 and loss_train[-1] < loss_train[0] and acc[-1] >= 75
    import shutil
    shutil.rmtree(model._get_path('checkpoints'), ignore_errors=True)
    shutil.rmtree(model._get_path('summaries'), ignore_errors=True)
