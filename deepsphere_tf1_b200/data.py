"""In-memory datasets feeding `fit` (the protocol of the reference's deepsphere/data.py).

`fit` consumes `.N`, `.iter(batch_size)` (an endless iterator of `(data, labels)` numpy
batches) and `.get_all_data()` (reference deepsphere/models.py:486-507).  Same classes,
arguments and batch contents as the reference (data.py:11-175); batches are produced by
modular index arithmetic over the sample order instead of itertools cycle / zip_longest, so
a batch is one fancy-indexing gather.  A batch that reaches the end of the data wraps
around to its start, exactly like the reference's cycle + grouper.
"""
import numpy as np


def grouper(iterable, n, fillvalue=None):
    """Chunk an iterable into tuples of length n, padding the last one (data.py:168-175)."""
    chunk = []
    for item in iterable:
        chunk.append(item)
        if len(chunk) == n:
            yield tuple(chunk)
            chunk = []
    if chunk:
        yield tuple(chunk + [fillvalue] * (n - len(chunk)))


class LabeledDataset(object):
    """Dataset of (sample, label) pairs with optional shuffling and per-batch transform."""

    def __init__(self, X, label, shuffle=True, transform=None):
        self._shuffle = shuffle
        self._transform = transform
        self._N = len(X)
        self._p = np.random.permutation(self._N) if shuffle else np.arange(self._N)
        self._X = np.asarray(X).astype(np.float32)[self._p]
        self._label = np.asarray(label)[self._p]

    def get_all_data(self):
        """All the data (in the stored, possibly shuffled, order)."""
        return self._X, self._label

    def get_labels(self):
        return self._label

    def get_samples(self, N=100, transform=True):
        """The N first samples."""
        if self._transform and transform:
            return self._transform(self._X[:N]), self._label[:N]
        return self._X[:N], self._label[:N]

    def _batches(self, batch_size):
        """Endless stream of index arrays (or scalar indexes for batch_size == 1)."""
        order = np.random.permutation(self._N) if self._shuffle else np.arange(self._N)
        self._p = order
        pos = 0
        while True:
            if batch_size > 1:
                yield order[(pos + np.arange(batch_size)) % self._N]
            else:
                yield order[pos % self._N]
            pos = (pos + batch_size) % self._N

    def iter(self, batch_size=1):
        """Endless iterator over batches `(data[batch_size, ...], label[batch_size])`."""
        return self.__iter__(batch_size)

    def __iter__(self, batch_size=1):
        for sel in self._batches(batch_size):
            data, label = self._X[sel], self._label[sel]
            if self._transform:
                data = self._transform(data)
            yield data, label

    @property
    def shuffled(self):
        return self._shuffle

    @property
    def N(self):
        return self._N


class LabeledDatasetWithNoise(LabeledDataset):
    """Dataset with on-the-fly additive noise whose level ramps linearly from `start_level`
    to `end_level` over the first `nit` batches (data.py:95-147)."""

    def __init__(self, X, label, shuffle=True, start_level=0, end_level=1, nit=0, noise_func=None,
                 all_level=False):
        self._nit = nit
        self._sl = start_level
        self._el = end_level
        if noise_func is None:
            noise_func = GaussianNoise(loc=0, scale=1, seed=1, all_level=all_level)
        self._noise_func = noise_func
        super().__init__(X=X, label=label, shuffle=shuffle, transform=None)

    def _add_noise(self, X, level):
        return X + level * self._noise_func(size=X.shape)

    def noise_level(self, it):
        if it < self._nit:
            return self._sl + it / self._nit * (self._el - self._sl)
        return self._el

    def __iter__(self, batch_size=1):
        for it, sel in enumerate(self._batches(batch_size)):
            yield self._add_noise(self._X[sel], self.noise_level(it)), self._label[sel]


class GaussianNoise(object):
    """Seeded Gaussian noise generator (data.py:149-166)."""

    def __init__(self, loc=0.0, scale=1., seed=None, all_level=False):
        self.loc = loc
        self.scale = scale
        self.seed = seed
        self.all_level = all_level
        self.rs = np.random.RandomState(self.seed)

    def __call__(self, size):
        noise = self.scale * self.rs.randn(*size)
        if self.all_level:
            s = self.rs.rand(size[0])
            s[:int(size[0] / 4 * 3)] = 1
            s = s.reshape([size[0]] + [1] * (len(size) - 1))
            return s * noise + self.loc
        return noise + self.loc


class DeviceDataset(LabeledDataset):
    """The same dataset protocol with the samples RESIDENT IN HBM and batches assembled by one kernel (csrc/input.cu):
    `.iter(batch_size)` yields `(data, labels)` as CUDA tensors, which `fit` / `train_step` consume without a host -> device copy
    (the reference's host-side batch assembly + feed_dict is ~40 % of its training step, experiments/cosmo/demo.ipynb:1003).
    Sample order, shuffling (host numpy permutation, as upstream), noise-level ramp and labels are those of
    LabeledDataset / LabeledDatasetWithNoise; the noise itself comes from a counter-based device generator (Philox4x32-10), so
    its values differ from numpy's Mersenne Twister stream while its distribution and the ramp are the same.

    Arguments as LabeledDatasetWithNoise (`end_level=0` / `nit=0` with `start_level=end_level=0` gives the noise-free
    LabeledDataset), plus `device` and `seed`."""

    def __init__(self, X, label, shuffle=True, start_level=0, end_level=0, nit=0, all_level=False, scale=1., loc=0.0,
                 device=None, seed=1):
        import torch
        super().__init__(X=X, label=label, shuffle=shuffle, transform=None)
        self._nit, self._sl, self._el = nit, start_level, end_level
        self._all_level, self._scale, self._loc = all_level, float(scale), float(loc)
        if self._loc != 0.0:
            raise NotImplementedError('a noise offset (loc) is not part of the device pipeline')
        self._device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self._Xd = torch.from_numpy(np.ascontiguousarray(self._X)).to(self._device)
        lab = np.ascontiguousarray(self._label)
        self._int_labels = lab.dtype.kind in 'iub'
        self._ld = torch.from_numpy(lab.astype(np.int32 if self._int_labels else np.float32)).to(self._device)
        self._seed = int(seed)
        self._calls = 0
        self._rs = np.random.RandomState(self._seed)

    noise_level = LabeledDatasetWithNoise.noise_level

    def __iter__(self, batch_size=1):
        import torch
        from ._lib import call
        row_len = int(np.prod(self._X.shape[1:]))
        lab_len = int(np.prod(self._label.shape[1:])) if self._label.ndim > 1 else 1
        nb = max(batch_size, 1)
        for it, sel in enumerate(self._batches(batch_size)):
            idx = torch.from_numpy(np.atleast_1d(np.asarray(sel)).astype(np.int32)).to(self._device, non_blocking=True)
            out = torch.empty((nb,) + self._X.shape[1:], dtype=torch.float32, device=self._device)
            level = float(self.noise_level(it)) * self._scale
            sscale = None
            if self._all_level and level != 0.0:           # GaussianNoise(all_level=True): a random level for the last quarter
                s = self._rs.rand(nb).astype(np.float32)
                s[:int(nb / 4 * 3)] = 1
                sscale = torch.from_numpy(s).to(self._device, non_blocking=True)
            # every call draws from its own block of the counter space
            call('gather_noise', self._Xd, idx, nb, row_len, level, sscale, self._seed, self._calls << 40, out)
            self._calls += 1
            if self._int_labels:
                lab = torch.empty((nb,) + self._label.shape[1:], dtype=torch.int32, device=self._device)
                call('gather_labels', self._ld, idx, nb, lab_len, lab)
            else:
                lab = self._ld[idx.long()]
            if batch_size == 1:                            # the reference yields un-batched samples for batch_size 1
                out, lab = out[0], lab[0]
            yield out, lab
