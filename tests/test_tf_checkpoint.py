"""TensorFlow V2 checkpoint files without TensorFlow (deepsphere_tf1_b200/tf_checkpoint.py): table format, protobuf entries,
checksums.  The reference ships no checkpoint, so the reader is checked against the writer and the format's invariants."""
import struct

import numpy as np
import pytest

from deepsphere_tf1_b200 import tf_checkpoint as tfc


def _vars(n_extra=0):
    rs = np.random.RandomState(0)
    v = {'conv1/weights': rs.randn(5, 16).astype(np.float32), 'conv1/bias': np.zeros((1, 1, 16), np.float32),
         'conv1/batch_normalization/moving_mean': rs.randn(16).astype(np.float32),
         'conv1/batch_normalization/moving_variance': rs.rand(16).astype(np.float32) + 0.5,
         'training/global_step': np.asarray(1440, np.int64), 'conv1/weights/Adam': rs.randn(5, 16).astype(np.float32)}
    for i in range(n_extra):
        v['extra/var_%04d' % i] = rs.randn(3, 2).astype(np.float32)
    return v


@pytest.mark.parametrize('n_extra', [0, 300])
def test_round_trip_and_table_structure(tmp_path, n_extra):
    v = _vars(n_extra)
    prefix = str(tmp_path / 'model-1440')
    tfc.write_checkpoint(prefix, v)
    raw = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == 0xdb4775248b80fb57          # the table magic number
    idx = tfc.read_index(prefix + '.index')
    assert b'' in idx                                                       # bundle header under the empty key
    assert sorted(k.decode() for k in idx if k) == sorted(v)
    got = tfc.read_checkpoint(prefix)
    assert sorted(got) == sorted(v)
    for k in v:
        assert got[k].dtype == v[k].dtype and got[k].shape == v[k].shape and np.array_equal(got[k], v[k]), k
    assert tfc.latest_checkpoint(str(tmp_path)) == prefix
    # several data blocks once there are many variables; keys prefix-compressed inside a block
    if n_extra:
        assert raw.count(b'extra/var_') < n_extra


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'model-1')
    tfc.write_checkpoint(prefix, _vars())
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[10] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    with pytest.raises(ValueError, match='checksum'):
        tfc.read_checkpoint(prefix)
    tfc.write_checkpoint(prefix, _vars())
    raw = bytearray(open(prefix + '.index', 'rb').read())
    raw[5] ^= 1
    open(prefix + '.index', 'wb').write(bytes(raw))
    with pytest.raises(ValueError, match='checksum'):
        tfc.read_checkpoint(prefix)
    open(prefix + '.index', 'wb').write(b'\x00' * 100)
    with pytest.raises(ValueError, match='magic'):
        tfc.read_checkpoint(prefix)


@pytest.mark.gpu
def test_model_restores_from_tensorflow_checkpoint(dev, tmp_path):
    """save_tf_checkpoint / load_tf_checkpoint under the reference's variable names, optimizer slots included; a run directory
    that holds only TensorFlow files is restored like the reference's `op_saver.restore(latest_checkpoint)` (models.py:851-860)."""
    import torch
    from deepsphere_tf1_b200 import models, optimizers
    kw = dict(nsides=[4, 2, 2], new=False, F=[8, 2], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean', num_epochs=1,
              batch_size=4, verbose=False, scheduler=lambda s: 1e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr))
    a = models.deepsphere(seed=1, dir_name='_tfck_a', **kw)
    rs = np.random.RandomState(0)
    x, lab = rs.randn(4, 192, 1).astype(np.float32), rs.randint(0, 2, 4).astype(np.int32)
    for _ in range(3):
        a.train_step(x, lab)
    prefix = str(tmp_path / 'model-3')
    a.save_tf_checkpoint(prefix)
    names = tfc.read_checkpoint(prefix)
    assert {'conv1/weights', 'conv1/bias', 'conv1/batch_normalization/moving_mean', 'conv2/weights', 'training/global_step',
            'conv1/weights/Adam', 'conv1/weights/Adam_1'} <= set(names)
    assert names['conv1/weights'].shape == (3, 8) and int(names['training/global_step']) == 3
    b = models.deepsphere(seed=2, dir_name='_tfck_b', **kw)
    unused = b.load_tf_checkpoint(prefix)
    assert unused == [] and b.global_step == 3
    assert torch.equal(a._P.w, b._P.w) and all(torch.equal(a._state[k], b._state[k]) for k in a._state)
    la, lb = float(a.train_step(x, lab)[1]), float(b.train_step(x, lab)[1])
    # same weights AND same Adam moments; not bit for bit: the batch-norm sums are accumulated with atomics, whose order (hence
    # the last bit of a mean) differs from launch to launch
    assert la == pytest.approx(lb, rel=1e-5), (la, lb)
    # the weight gradients are summed with red.global.add: the order differs from launch to launch, not the moments they meet.  One
    # Adam step moves a weight by ~lr = 1e-2: a lost or misplaced moment shows at 1e-3, the summation order at 1e-7
    assert torch.allclose(a._P.w, b._P.w, rtol=0, atol=1e-5), float((a._P.w - b._P.w).abs().max())
    # a run directory written by the reference holds TensorFlow files only: predict() without a session restores it
    # (models.py:851-860), fit(restore=True) continues from its step (models.py:452-457) and leaves the files alone
    import os
    import shutil
    from deepsphere_tf1_b200.data import LabeledDataset
    c = models.deepsphere(seed=3, dir_name='_tfck_c', **kw)
    d = c._get_path('checkpoints')
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    b.save_tf_checkpoint(os.path.join(d, 'model-4'))
    np.testing.assert_array_equal(c.predict(x), b.predict(x, sess=b))
    assert c.global_step == 4
    xs, ls = rs.randn(24, 192).astype(np.float32), rs.randint(0, 2, 24)
    c.fit(LabeledDataset(xs, ls, shuffle=False), LabeledDataset(xs[:4], ls[:4], shuffle=False), restore=True, verbose=False)
    assert c.global_step == 6 and os.path.exists(os.path.join(d, 'model-4.index')) and os.path.exists(os.path.join(d, 'model-6.pt'))
    shutil.rmtree(d, ignore_errors=True)


def test_snappy_blocks_of_the_index_table_are_decoded(tmp_path, monkeypatch):
    """The table format allows snappy-compressed blocks (type byte 1).  The decoder against hand-made streams of every element
    kind, against a real snappy encoder (pyarrow's), and end to end on an index file whose blocks are compressed."""
    # literal 'ab', then a 1-byte-offset copy of 8 bytes at distance 2: the copy overlaps its own output
    assert tfc._snappy_decompress(bytes([10, 0x04]) + b'ab' + bytes([0x11, 0x02])) == b'ababababab'
    # 2-byte-offset copy (kind 2: length - 1 in the upper six bits) and a 61-byte literal with its length in one extra byte
    lit = bytes(range(61))
    assert tfc._snappy_decompress(bytes([71, 60 << 2, 60]) + lit + bytes([(10 - 1) << 2 | 2, 61, 0])) == lit + lit[:10]
    with pytest.raises(ValueError, match='header says'):
        tfc._snappy_decompress(bytes([5, 0x04]) + b'ab')
    with pytest.raises(ValueError, match='offset'):
        tfc._snappy_decompress(bytes([6, 0x04]) + b'ab' + bytes([0x01, 0x09]))
    pa = pytest.importorskip('pyarrow')
    if not pa.Codec.is_available('snappy'):
        pytest.skip('pyarrow without snappy')
    rs = np.random.RandomState(0)
    for n in (0, 1, 59, 60, 61, 300, 70000, 200000):
        for kind in ('random', 'text', 'runs'):
            if kind == 'random':
                raw = rs.bytes(n)
            elif kind == 'text':
                raw = (b'conv1/weights conv1/bias conv2/batch_normalization/moving_mean ' * (n // 60 + 1))[:n]
            else:
                raw = bytes(np.repeat(rs.randint(0, 4, size=n // 50 + 1), 50).astype(np.uint8)[:n])
            assert tfc._snappy_decompress(pa.compress(raw, codec='snappy', asbytes=True)) == raw, (n, kind)

    def emit_snappy(out, block):
        comp = pa.compress(block, codec='snappy', asbytes=True)
        off = len(out)
        out += comp + b'\x01' + struct.pack('<I', tfc._masked_crc(comp + b'\x01'))
        return tfc._varint(off) + tfc._varint(len(comp))

    import struct
    monkeypatch.setattr(tfc, '_emit', emit_snappy)
    v = _vars()
    prefix = str(tmp_path / 'model-7')
    tfc.write_checkpoint(prefix, v, block_entries=3)
    got = tfc.read_checkpoint(prefix)
    assert set(got) == set(v)
    for k in v:
        assert np.array_equal(got[k], v[k]), k
