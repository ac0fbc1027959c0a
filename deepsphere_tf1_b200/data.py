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
