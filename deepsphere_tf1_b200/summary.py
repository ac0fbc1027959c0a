"""TensorBoard event files without TensorFlow (the reference logs through `tf.summary.*` / `tf.summary.FileWriter`,
deepsphere/models.py:463, 575-603, 815-819, 838, 868).

A tfevents file is a sequence of records  [uint64 length][masked crc32c(length)][payload][masked crc32c(payload)]  whose payload
is a serialised `tensorflow.Event` protobuf; the few fields needed are encoded by hand:
  Event   : wall_time = 1 (double), step = 2 (int64), file_version = 3 (string), summary = 5 (message)
  Summary : value = 1 (repeated message);  Value: tag = 1 (string), simple_value = 2 (float), histo = 5 (message)
  HistogramProto: min = 1, max = 2, num = 3, sum = 4, sum_squares = 5 (double), bucket_limit = 6, bucket = 7 (packed double)
`read_events` parses the files back (used by the tests; TensorBoard itself reads them as written by TensorFlow)."""
import os
import socket
import struct
import time

import numpy as np

_CRC_TABLE = None


def _crc32c(data):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            tbl.append(c)
        _CRC_TABLE = tbl
    crc = 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _masked_crc(data):
    crc = _crc32c(data)
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wire):
    return _varint((field << 3) | wire)


def _len_delim(field, payload):
    return _key(field, 2) + _varint(len(payload)) + payload


def _double(field, v):
    return _key(field, 1) + struct.pack('<d', float(v))


def _float(field, v):
    return _key(field, 5) + struct.pack('<f', float(v))


def _histogram(values):
    """tf.summary.histogram's default buckets: exponential limits of ratio 1.1 around zero."""
    v = np.asarray(values, dtype=np.float64).ravel()
    pos = [1e-12]
    while pos[-1] < 1e20:
        pos.append(pos[-1] * 1.1)
    limits = np.array([-x for x in reversed(pos)] + [0.0] + pos + [np.finfo(np.float64).max])
    counts = np.histogram(v, bins=np.concatenate([[-np.inf], limits]))[0].astype(np.float64)
    keep = np.nonzero(counts)[0]
    lo, hi = (keep[0], keep[-1] + 1) if len(keep) else (0, 1)
    msg = (_double(1, v.min() if v.size else 0) + _double(2, v.max() if v.size else 0) + _double(3, v.size) + _double(4, v.sum())
           + _double(5, (v * v).sum()))
    msg += _len_delim(6, struct.pack('<%dd' % (hi - lo), *limits[lo:hi]))
    msg += _len_delim(7, struct.pack('<%dd' % (hi - lo), *counts[lo:hi]))
    return msg


class FileWriter(object):
    """`tf.summary.FileWriter(logdir)` for scalars and histograms."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        name = 'events.out.tfevents.{:010d}.{}'.format(int(time.time()), socket.gethostname())
        self.path = os.path.join(logdir, name)
        self._f = open(self.path, 'ab')
        self._write(_double(1, time.time()) + _len_delim(3, b'brain.Event:2'))

    def _write(self, event):
        header = struct.pack('<Q', len(event))
        self._f.write(header + struct.pack('<I', _masked_crc(header)) + event + struct.pack('<I', _masked_crc(event)))

    def add_summary(self, step, scalars=None, histograms=None):
        summ = b''
        for tag, v in (scalars or {}).items():
            summ += _len_delim(1, _len_delim(1, tag.encode()) + _float(2, v))
        for tag, arr in (histograms or {}).items():
            summ += _len_delim(1, _len_delim(1, tag.encode()) + _len_delim(5, _histogram(arr)))
        self._write(_double(1, time.time()) + _key(2, 0) + _varint(int(step)) + _len_delim(5, summ))

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()


class StepTrace(object):
    """`profile=True` (deepsphere/models.py:514-519, 604-605): the reference runs the training step of every evaluation with
    `tf.RunOptions(trace_level=FULL_TRACE)` and hands the `RunMetadata` to the summary writer.  Here the same step runs under
    torch.profiler (CUPTI activity records; the kernels of a CUDA-graph replay are traced like eager launches) and what is kept
    is the device timeline: `trace-step<N>.json` (Chrome trace format: chrome://tracing, Perfetto, TensorBoard's profile
    plugin) and `kernels-step<N>.txt` (time and launch count per kernel, largest first) in the summary directory.  Numbers taken
    under the profiler are attribution, not measurements: the step time that `fit` returns excludes nothing and is slower on
    traced steps, exactly as upstream."""

    def __init__(self):
        from torch.profiler import profile, ProfilerActivity
        self._prof = profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])
        self._on = False

    def start(self):
        self._prof.__enter__()
        self._on = True

    def stop(self):
        import torch
        if self._on:
            torch.cuda.synchronize()
            self._prof.__exit__(None, None, None)
            self._on = False

    def kernels(self):
        """[(kernel name, microseconds, launches)] of the traced region, largest first."""
        tot = {}
        for ev in self._prof.events():
            if ev.device_type.name != 'CUDA' or not ev.name:
                continue
            name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '').replace('dsb200::', '').split('(')[0]
            d = tot.setdefault(name, [0.0, 0])
            d[0] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
            d[1] += 1
        return sorted(((n, v[0], v[1]) for n, v in tot.items()), key=lambda r: -r[1])

    def write(self, logdir, step):
        """Write both files; returns the kernel time of the traced step in microseconds."""
        self.stop()
        os.makedirs(logdir, exist_ok=True)
        self._prof.export_chrome_trace(os.path.join(logdir, 'trace-step{}.json'.format(step)))
        rows = self.kernels()
        total = sum(r[1] for r in rows)
        with open(os.path.join(logdir, 'kernels-step{}.txt'.format(step)), 'w') as f:
            f.write('{:>10s} {:>5s} {:>6s}  kernel\n'.format('us', 'n', 'share'))
            for name, us, n in rows:
                f.write('{:10.1f} {:5d} {:5.1f}%  {}\n'.format(us, n, 100 * us / max(total, 1e-30), name[:120]))
            f.write('{:10.1f} total kernel time of the step\n'.format(total))
        return total


# ---------------------------------------------------------------------------------------------- reader (tests, tooling)
def _read_varint(buf, i):
    n = shift = 0
    while True:
        b = buf[i]
        i += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return n, i


def _fields(buf):
    i = 0
    while i < len(buf):
        key, i = _read_varint(buf, i)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, i = _read_varint(buf, i)
        elif wire == 1:
            v, i = struct.unpack('<d', buf[i:i + 8])[0], i + 8
        elif wire == 5:
            v, i = struct.unpack('<f', buf[i:i + 4])[0], i + 4
        elif wire == 2:
            n, i = _read_varint(buf, i)
            v, i = buf[i:i + n], i + n
        else:
            raise ValueError('unsupported wire type')
        yield field, wire, v


def read_events(path):
    """[(step, {tag: value | histogram dict})] of one event file; checks both checksums of every record."""
    out = []
    data = open(path, 'rb').read()
    i = 0
    while i < len(data):
        header = data[i:i + 8]
        (n,) = struct.unpack('<Q', header)
        if struct.unpack('<I', data[i + 8:i + 12])[0] != _masked_crc(header):
            raise ValueError('corrupt length checksum')
        payload = data[i + 12:i + 12 + n]
        if struct.unpack('<I', data[i + 12 + n:i + 16 + n])[0] != _masked_crc(payload):
            raise ValueError('corrupt payload checksum')
        i += 16 + n
        step, values = 0, {}
        for field, wire, v in _fields(payload):
            if field == 2:
                step = v
            elif field == 5:
                for f2, _, val in _fields(v):
                    if f2 != 1:
                        continue
                    tag, item = None, None
                    for f3, _, x in _fields(val):
                        if f3 == 1:
                            tag = x.decode()
                        elif f3 == 2:
                            item = x
                        elif f3 == 5:
                            h = {}
                            for f4, w4, y in _fields(x):
                                name = {1: 'min', 2: 'max', 3: 'num', 4: 'sum', 5: 'sum_squares', 6: 'bucket_limit', 7: 'bucket'}[f4]
                                h[name] = list(struct.unpack('<%dd' % (len(y) // 8), y)) if w4 == 2 else y
                            item = h
                    values[tag] = item
        if values:
            out.append((step, values))
    return out
