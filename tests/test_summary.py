"""TensorBoard event files written without TensorFlow (deepsphere_tf1_b200/summary.py): record framing with masked CRC-32C,
hand-encoded Event / Summary / HistogramProto, read back by the module's own parser."""
import os
import struct

import numpy as np
import pytest

from deepsphere_tf1_b200 import summary


def test_crc32c_known_answers():
    assert summary._crc32c(b'123456789') == 0xE3069283                      # the standard check value of CRC-32C (Castagnoli)
    assert summary._crc32c(b'') == 0
    assert summary._crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4: 32 bytes of zeros


def test_event_file_round_trip(tmp_path):
    w = summary.FileWriter(str(tmp_path))
    rs = np.random.RandomState(0)
    vals = rs.randn(1000) * 0.1
    w.add_summary(200, {'loss/total': 0.6931, 'learning_rate': 2e-4, 'validation/accuracy': 87.5}, {'conv1/weights': vals})
    w.add_summary(400, {'loss/total': 0.25})
    w.close()
    raw = open(w.path, 'rb').read()
    (n,) = struct.unpack('<Q', raw[:8])
    assert raw[12:12 + n].endswith(b'brain.Event:2')                         # the file-version record TensorBoard expects first
    ev = summary.read_events(w.path)
    assert [e[0] for e in ev] == [200, 400]
    assert ev[0][1]['loss/total'] == pytest.approx(0.6931, rel=1e-6) and ev[0][1]['validation/accuracy'] == pytest.approx(87.5)
    h = ev[0][1]['conv1/weights']
    assert h['num'] == 1000 and h['min'] == pytest.approx(vals.min()) and h['max'] == pytest.approx(vals.max())
    assert h['sum'] == pytest.approx(vals.sum()) and h['sum_squares'] == pytest.approx((vals ** 2).sum())
    assert sum(h['bucket']) == 1000 and len(h['bucket']) == len(h['bucket_limit'])
    lim = np.array(h['bucket_limit'])
    assert np.all(np.diff(lim) > 0) and lim[-1] >= vals.max()
    # a flipped byte is detected
    bad = bytearray(raw)
    bad[40] ^= 1
    p2 = tmp_path / 'bad'
    p2.write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        summary.read_events(str(p2))


def test_tensorboard_reads_the_event_file(tmp_path):
    """The files are for TensorBoard: its own loader (record checksums, Event / Summary protos, the scalar and histogram
    plugins' data-compat layer) must accept them and return the values."""
    efl = pytest.importorskip('tensorboard.backend.event_processing.event_file_loader')
    from tensorboard.util import tensor_util
    w = summary.FileWriter(str(tmp_path))
    vals = np.linspace(-1, 1, 101)
    w.add_summary(7, {'loss/total': 0.125, 'validation/accuracy': 93.75}, {'conv1/weights': vals})
    w.close()
    events = list(efl.EventFileLoader(w.path).Load())
    assert events[0].file_version == 'brain.Event:2'
    ev = events[1]
    assert ev.step == 7
    got = {v.tag: v for v in ev.summary.value}
    assert got['loss/total'].metadata.plugin_data.plugin_name == 'scalars'
    assert float(tensor_util.make_ndarray(got['loss/total'].tensor)) == 0.125
    assert float(tensor_util.make_ndarray(got['validation/accuracy'].tensor)) == 93.75
    assert got['conv1/weights'].metadata.plugin_data.plugin_name == 'histograms'
    h = tensor_util.make_ndarray(got['conv1/weights'].tensor)          # [buckets, (left, right, count)]
    assert h.shape[1] == 3 and h[:, 2].sum() == 101


@pytest.mark.gpu
@pytest.mark.parametrize('graph', [False, True])
def test_fit_with_profile_writes_the_device_trace(graph):
    """`profile=True` (models.py:514-519, 604-605): the training step of every evaluation is traced; the summary directory holds
    its timeline and per-kernel table, and the event file the kernel time of that step."""
    import glob
    import json
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from deepsphere_tf1_b200 import models, optimizers
    from deepsphere_tf1_b200.data import LabeledDataset
    rs = np.random.RandomState(0)
    x, labels = rs.randn(16, 192).astype(np.float32), rs.randint(0, 2, size=16)
    train, val = LabeledDataset(x, labels), LabeledDataset(x[:8], labels[:8], shuffle=False)
    model = models.deepsphere(nsides=[4, 2, 2], new=False, F=[8, 2], K=[3, 3], batch_norm=[True, True], M=[], statistics='mean',
                              num_epochs=4, batch_size=8, eval_frequency=4, seed=1, verbose=False, profile=True, cuda_graph=graph,
                              scheduler=lambda step: 1e-2, optimizer=lambda lr: optimizers.AdamOptimizer(lr),
                              dir_name='_test_profile')
    model.fit(train, val, restore=False, verbose=False)
    logdir = model._get_path('summaries')
    assert sorted(os.path.basename(f) for f in glob.glob(logdir + '/trace-step*.json')) == ['trace-step4.json', 'trace-step8.json']
    table = open(logdir + '/kernels-step8.txt').read()
    assert 'optimizer_kernel' in table and 'total kernel time of the step' in table      # this library's kernels, by name
    trace = json.load(open(logdir + '/trace-step8.json'))
    assert any('optimizer_kernel' in str(e.get('name', '')) for e in trace['traceEvents'])
    ev = summary.read_events(glob.glob(logdir + '/events.out.tfevents.*')[0])
    assert [e[0] for e in ev] == [4, 8] and ev[-1][1]['profile/step_kernel_us'] > 0
