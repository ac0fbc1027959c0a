"""TensorFlow checkpoint (V2 "tensor bundle") files without TensorFlow: the reference saves / restores its variables with
`tf.train.Saver` (deepsphere/models.py:723, 847-860, 607 -> `checkpoints/<dir_name>/model-<step>.{index,data-00000-of-00001}`);
the variable names there are the ones this package uses (`conv1/weights`, `conv1/bias`, `conv1/batch_normalization/moving_mean`,
`logits/weights`, ...), so a reference checkpoint loads into `models.cgcnn` by name and vice versa.

Format (tensorflow/core/util/tensor_bundle + the LevelDB-style table of tensorflow/core/lib/io/table*):
  <prefix>.index   an SSTable: data blocks of prefix-compressed (key, value) entries [varint shared, varint non_shared, varint
                   value_len, key suffix, value] followed by the restart offsets and their count, each block with a 5-byte trailer
                   (compression type, masked CRC-32C); an index block mapping separator keys to block handles; a 48-byte footer
                   (meta-index handle, index handle, padding, magic 0xdb4775248b80fb57).  Key "" -> BundleHeaderProto, key <variable
                   name> -> BundleEntryProto (dtype, shape, shard_id, offset, size, crc32c).
  <prefix>.data-00000-of-00001   the raw little-endian tensor bytes at those offsets.
Written from the published format; the reference repository ships no checkpoint file, so reading is verified against this module's
own writer, the CRCs and the structural invariants only (tests/test_tf_checkpoint.py).  Snappy-compressed index blocks (block type 1) are
decoded (`_snappy_decompress`, checked against a real snappy encoder); the writer emits uncompressed blocks as TensorFlow's BundleWriter does.
"""
import os
import struct

import numpy as np

from .summary import _masked_crc, _varint, _read_varint, _len_delim, _key

MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------- reading
def _fields(buf):
    """(field, wire type, value) of a serialised protobuf message; fixed32 / fixed64 values as unsigned integers."""
    i = 0
    while i < len(buf):
        key, i = _read_varint(buf, i)
        field, wire = key >> 3, key & 7
        if wire == 0:
            v, i = _read_varint(buf, i)
        elif wire == 1:
            v, i = struct.unpack('<Q', buf[i:i + 8])[0], i + 8
        elif wire == 5:
            v, i = struct.unpack('<I', buf[i:i + 4])[0], i + 4
        elif wire == 2:
            n, i = _read_varint(buf, i)
            v, i = buf[i:i + n], i + n
        else:
            raise ValueError('unsupported protobuf wire type {}'.format(wire))
        yield field, wire, v


def _snappy_decompress(buf):
    """Raw snappy format (format_description.txt of google/snappy): varint uncompressed length, then elements tagged by their
    low two bits -- 00 literal (length - 1 in the upper six bits, or in the 1..4 following bytes for 60..63), 01 copy with a
    3-bit length - 4 and an 11-bit offset, 10 / 11 copies with a 6-bit length - 1 and a 2- / 4-byte little-endian offset.
    Copies may overlap their own output (offset < length repeats a pattern)."""
    n, i = _read_varint(buf, 0)
    out = bytearray()
    while i < len(buf):
        tag = buf[i]
        i += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[i:i + nb], 'little')
                i += nb
            ln += 1
            if i + ln > len(buf):
                raise ValueError('snappy: literal runs past the end of the block')
            out += buf[i:i + ln]
            i += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[i]
            i += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[i:i + 2], 'little')
            i += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[i:i + 4], 'little')
            i += 4
        if off == 0 or off > len(out):
            raise ValueError('snappy: copy offset outside the output')
        for _ in range(ln):                             # byte by byte: the source may overlap what is being written
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy: {} bytes decoded, header says {}'.format(len(out), n))
    return bytes(out)


def _block(data, offset, size):
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    (crc,) = struct.unpack('<I', data[offset + size + 1:offset + size + 5])
    if crc != _masked_crc(raw + bytes([ctype])):
        raise ValueError('tensor bundle index: block checksum mismatch')
    if ctype == 1:                                      # kSnappyCompression of the table format
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise NotImplementedError('tensor bundle index: block compression type {} is not supported'.format(ctype))
    (nrestart,) = struct.unpack('<I', raw[-4:])
    end = len(raw) - 4 - 4 * nrestart
    entries, key, i = [], b'', 0
    while i < end:
        shared, i = _read_varint(raw, i)
        non_shared, i = _read_varint(raw, i)
        vlen, i = _read_varint(raw, i)
        key = key[:shared] + raw[i:i + non_shared]
        i += non_shared
        entries.append((key, raw[i:i + vlen]))
        i += vlen
    return entries


def _handle(buf, i=0):
    off, i = _read_varint(buf, i)
    size, i = _read_varint(buf, i)
    return off, size, i


def read_index(path):
    """{key bytes: value bytes} of an SSTable file."""
    data = open(path, 'rb').read()
    if len(data) < 48 or struct.unpack('<Q', data[-8:])[0] != MAGIC:
        raise ValueError('{}: not a tensor bundle index (bad magic)'.format(path))
    footer = data[-48:]
    _, _, i = _handle(footer)                                     # meta-index block (empty)
    ioff, isize, _ = _handle(footer, i)
    out = {}
    for _, h in _block(data, ioff, isize):
        boff, bsize, _ = _handle(h)
        for k, v in _block(data, boff, bsize):
            out[k] = v
    return out


def _parse_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'slices': 0}
    for field, wire, v in _fields(buf):
        if field == 1:
            e['dtype'] = v
        elif field == 2:
            for f2, _, dim in _fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, x in _fields(dim):
                        if f3 == 1:
                            size = x
                    e['shape'].append(size)
        elif field == 3:
            e['shard_id'] = v
        elif field == 4:
            e['offset'] = v
        elif field == 5:
            e['size'] = v
        elif field == 6:
            e['crc32c'] = v
        elif field == 7:
            e['slices'] += 1
    return e


def read_checkpoint(prefix, verify=True):
    """{variable name: numpy array} of the checkpoint `prefix` (e.g. checkpoints/run/model-1440)."""
    index = read_index(prefix + '.index')
    header = index.pop(b'', None)
    num_shards = 1
    if header is not None:
        for field, _, v in _fields(header):
            if field == 1:
                num_shards = v
            elif field == 2 and v != 0:
                raise NotImplementedError('big-endian tensor bundles are not supported')
    shards = {}
    out = {}
    for key, value in index.items():
        e = _parse_entry(value)
        if e['slices']:
            raise NotImplementedError('partitioned variable {} is not supported'.format(key.decode()))
        if e['dtype'] not in _DTYPES:
            continue                                              # strings etc.: nothing the model needs
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = open('{}.data-{:05d}-of-{:05d}'.format(prefix, sid, num_shards), 'rb').read()
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if verify and e['crc32c'] is not None and _masked_crc(raw) != e['crc32c']:
            raise ValueError('tensor {}: data checksum mismatch'.format(key.decode()))
        out[key.decode()] = np.frombuffer(raw, dtype=_DTYPES[e['dtype']]).reshape(e['shape']).copy()
    return out


# ---------------------------------------------------------------------------------------------- writing
def _fixed32(field, v):
    return _key(field, 5) + struct.pack('<I', v)


def _build_block(entries, restart_interval=16):
    buf, restarts, prev = bytearray(), [], b''
    for n, (k, v) in enumerate(entries):
        shared = 0
        if n % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        buf += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    buf += b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))
    return bytes(buf)


def _emit(out, block):
    off = len(out)
    out += block + b'\x00' + struct.pack('<I', _masked_crc(block + b'\x00'))
    return _varint(off) + _varint(len(block))


def write_checkpoint(prefix, tensors, block_entries=64):
    """Write {name: array} as `prefix.index` + `prefix.data-00000-of-00001` (what tf.train.Saver().save produces for the same
    variables, one shard, no compression)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data = bytearray()
    header = _key(1, 0) + _varint(1) + _key(2, 0) + _varint(0) + _len_delim(3, _key(1, 0) + _varint(1))
    entries = [(b'', header)]
    for name in sorted(tensors):
        a = np.asarray(tensors[name])                              # (ascontiguousarray would turn a scalar into shape (1,))
        if a.dtype not in _DTYPE_IDS:
            a = a.astype(np.float32)
        raw = a.tobytes(order='C')
        shape = b''.join(_len_delim(2, _key(1, 0) + _varint(int(d))) for d in a.shape)
        entry = _key(1, 0) + _varint(_DTYPE_IDS[a.dtype]) + _len_delim(2, shape)
        if len(data):
            entry += _key(4, 0) + _varint(len(data))
        entry += _key(5, 0) + _varint(len(raw)) + _fixed32(6, _masked_crc(raw))
        entries.append((name.encode(), entry))
        data += raw
    out = bytearray()
    index_entries = []
    for i in range(0, len(entries), block_entries):
        chunk = entries[i:i + block_entries]
        index_entries.append((chunk[-1][0], _emit(out, _build_block(chunk))))
    meta = _emit(out, _build_block([]))
    idx = _emit(out, _build_block(index_entries, restart_interval=1))
    footer = meta + idx
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC)
    open(prefix + '.index', 'wb').write(bytes(out))
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    d = os.path.dirname(os.path.abspath(prefix))
    open(os.path.join(d, 'checkpoint'), 'w').write('model_checkpoint_path: "{}"\n'.format(os.path.basename(prefix)))


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file, else the highest `model-<step>.index`."""
    state = os.path.join(directory, 'checkpoint')
    if os.path.exists(state):
        for line in open(state):
            if line.startswith('model_checkpoint_path:'):
                name = line.split(':', 1)[1].strip().strip('"')
                p = name if os.path.isabs(name) else os.path.join(directory, name)
                if os.path.exists(p + '.index'):
                    return p
    best, step = None, -1
    for f in os.listdir(directory) if os.path.isdir(directory) else []:
        if f.startswith('model-') and f.endswith('.index'):
            try:
                s = int(f[len('model-'):-len('.index')])
            except ValueError:
                continue
            if s > step:
                best, step = os.path.join(directory, f[:-len('.index')]), s
    return best
