"""TensorFlow checkpoint format without TensorFlow (dnncancerannotator_b200/utils/tf_checkpoint.py; engine.py:55-78,105):
known answers of the primitives, table / bundle round trips, and the object-graph name mapping of the reference's
classes.  CPU only."""
import os
import struct

import numpy as np
import pytest

from dnncancerannotator_b200.utils import tf_checkpoint as T

OPTS = dict(n_filters_first=4, n_downsample=2, rate=2, kernel_size=3, conv_stride=1, padding='same')


def test_crc32c_known_answers():
    assert T.crc32c(b'123456789') == 0xE3069283                      # the CRC-32C check value
    assert T.crc32c(b'\x00' * 32) == 0x8A9136AA                      # RFC 3720 B.4
    assert T.crc32c(b'\xff' * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(b'6789', T.crc32c(b'12345')) == 0xE3069283       # incremental
    for v in (0, 1, 0xdeadbeef, 0xffffffff):
        assert T.unmask_crc(T.mask_crc(v)) == v
    assert T.mask_crc(T.crc32c(b'foo')) != T.crc32c(b'foo')


def test_varint_and_proto_roundtrip():
    for v in (0, 1, 127, 128, 300, 2 ** 32, 2 ** 63 - 1):
        b = T.write_varint(v)
        assert T.read_varint(b, 0) == (v, len(b))
    assert T.write_varint(300) == b'\xac\x02'                        # the protobuf documentation's example
    e = T.parse_bundle_entry(T._entry_proto(1, (3, 3, 4, 8), 4096, 1152, 0x12345678))
    assert e['dtype'] == 1 and e['shape'] == (3, 3, 4, 8) and e['offset'] == 4096 and e['size'] == 1152 and e['crc32c'] == 0x12345678


def test_snappy_known_answers():
    # literal "abcd" + copy(offset 4, length 8): overlapping copy repeats the pattern
    src = T.write_varint(12) + bytes([3 << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4])
    assert T.snappy_uncompress(src) == b'abcdabcdabcd'
    # 2-byte-offset copy and a long literal (length byte follows the tag)
    lit = bytes(range(70))
    src = T.write_varint(70 + 5) + bytes([60 << 2, 69]) + lit + bytes([((5 - 1) << 2) | 2, 70, 0])
    assert T.snappy_uncompress(src) == lit + lit[:5]
    with pytest.raises(ValueError):
        T.snappy_uncompress(T.write_varint(4) + bytes([((4 - 4) << 2) | 1, 9]))       # copy before any output


def test_table_roundtrip_with_prefix_compression_and_snappy_block(tmp_path):
    items = [(f'layer/{i:04d}/kernel'.encode(), os.urandom(5 + i % 7)) for i in range(300)]
    p = str(tmp_path / 't.index')
    T.write_table(p, items, block_entries=37)
    assert T.read_table(p) == items
    # the footer magic and a corrupted block are detected
    raw = bytearray(open(p, 'rb').read())
    assert struct.unpack('<Q', raw[-8:])[0] == T.TABLE_MAGIC
    raw[10] ^= 0xff
    open(p, 'wb').write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_table(p)
    # a snappy-compressed data block (type byte 1), as TensorFlow's table builder writes when it pays
    block = T._build_block(items[:3])
    comp = T.write_varint(len(block)) + b''.join(bytes([(min(60, len(block) - i) - 1) << 2]) + block[i:i + 60]
                                                 for i in range(0, len(block), 60))
    with open(p, 'wb') as f:
        f.write(comp + b'\x01' + struct.pack('<I', T.mask_crc(T.crc32c(comp + b'\x01'))))
        moff, msize = T._emit_block(f, T._build_block([]))
        ioff, isize = T._emit_block(f, T._build_block([(items[2][0], T.write_varint(0) + T.write_varint(len(comp)))], 1))
        footer = T.write_varint(moff) + T.write_varint(msize) + T.write_varint(ioff) + T.write_varint(isize)
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', T.TABLE_MAGIC))
    assert T.read_table(p) == items[:3]


@pytest.mark.parametrize('cls,bn', [('UNetAnnotator', False), ('UNetAnnotator', True), ('MulmoUNetAnnotator', True)])
def test_export_and_load_through_the_reference_object_graph(tmp_path, cls, bn):
    from dnncancerannotator_b200.models import tf_models
    a = getattr(tf_models, cls)(**OPTS, bn=bn, seed=1)
    a.build((None, 32, 32, 3))
    prefix = str(tmp_path / 'checkpoints' / 'ckpt-300')
    a.save_weights(prefix, save_format='tf')
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    rd = T.CheckpointReader(prefix)
    nodes = rd.object_graph()
    # the reference's attribute names (components.py / unet.py) are the edges of the graph
    assert set(nodes[0]['children']) >= {'unet', 'last_conv'}
    enc = 'encoders' if cls.startswith('Mulmo') else 'encoder'
    assert enc in nodes[nodes[0]['children']['unet']]['children']
    path0 = ['unet', enc] + (['0'] if cls.startswith('Mulmo') else []) + ['downsamples', '0', 'convchain', 'layer_with_weights-0', 'kernel']
    key, _ = rd.resolve(path0, nodes)
    assert key == '/'.join(path0) + T.VAR_SUFFIX and key in rd.keys()
    wa = a.get_weights()
    first = 'enc/0/d0/conv0/kernel' if cls.startswith('Mulmo') else 'enc/d0/conv0/kernel'
    np.testing.assert_array_equal(rd.get_tensor(key), wa[first])
    if bn:      # BatchNorm of conv0 is the second layer of the Sequential; the pool's BatchNorm hangs under `pool`
        assert rd.resolve(path0[:-3] + ['convchain', 'layer_with_weights-1', 'moving_variance'], nodes)
        assert rd.resolve(path0[:-3] + ['pool', 'layer_with_weights-0', 'gamma'], nodes)
    # a fresh model with other initial values takes every variable from the file
    b = getattr(tf_models, cls)(**OPTS, bn=bn, seed=2)
    b.build((None, 32, 32, 3))
    assert any(not np.array_equal(wa[k], v) for k, v in b.get_weights().items())
    st = b.load_weights(prefix)
    st.assert_existing_objects_matched()
    wb = b.get_weights()
    for k in wa:
        np.testing.assert_array_equal(wa[k], wb[k])
    assert b.list_checkpoints(str(tmp_path)) == {300: prefix}
    # a deeper model: the decoder shapes differ (TF raises on the shape, so does set_weights) and with the shapes out
    # of the way the third level's variables are reported missing like engine.py:75's assert does
    del rd
    other = getattr(tf_models, cls)(**dict(OPTS, n_downsample=3), bn=bn, seed=0)
    other.build((None, 32, 32, 3))
    with pytest.raises(ValueError):
        other.load_weights(prefix)
    loaded, missing, unused = [], [], []
    rd2 = T.CheckpointReader(prefix)
    nodes2 = rd2.object_graph()
    for name, cands in T.reference_paths(other).items():
        (loaded if any(rd2.resolve(p, nodes2) for p in cands) else missing).append(name)
    assert missing and all('d2' in n or 'u2' in n for n in missing), missing


def test_multiresunet_tf_format_is_refused_clearly(tmp_path):
    """Keras' functional-model key order is not restated: the TF format says so instead of mis-mapping same-shaped layers."""
    from dnncancerannotator_b200.models import tf_models
    m = tf_models.MultiResUnet(None, None, 5)
    m.build((None, 32, 32, 5))
    with pytest.raises(NotImplementedError, match='npz'):
        m.save_weights(str(tmp_path / 'ckpt-1'), save_format='tf')
    m.save_weights(str(tmp_path / 'ckpt-1'))                     # own format works without a device
    m2 = tf_models.MultiResUnet(None, None, 5, seed=5)
    m2.build((None, 32, 32, 5))
    m2.load_weights(str(tmp_path / 'ckpt-1')).assert_existing_objects_matched()
    w1, w2 = m.get_weights(), m2.get_weights()
    assert all(np.array_equal(w1[k], w2[k]) for k in w1)
