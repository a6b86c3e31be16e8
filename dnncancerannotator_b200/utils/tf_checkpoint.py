"""Reader (and a minimal writer) for TensorFlow's checkpoint format, WITHOUT TensorFlow.

The reference saves with ``tf.keras.callbacks.ModelCheckpoint(ckpt_path, save_weights_only=True)`` (engine.py:105) and
resumes with ``model.load_weights(latest_ckpt)`` (engine.py:55-78): TF2 object-based checkpoints
``checkpoints/ckpt-<step>.index`` + ``ckpt-<step>.data-00000-of-00001``.  This module restates the published on-disk
format so that such a file can be loaded into this package's models (TensorFlow is not installable in this image):

* ``.index`` is a LevelDB-format sorted string table (tensorflow/core/lib/io/table*.cc, format.cc): data blocks of
  prefix-compressed (key, value) entries with a restart array, an index block, a 48-byte footer ending in the magic
  0xdb4775248b80fb57; every block carries a 1-byte compression tag (0 none, 1 snappy) and a masked CRC32C.
* key ``""`` holds a ``BundleHeaderProto``; every other key a ``BundleEntryProto`` (tensor_bundle.proto): dtype, shape,
  shard_id, offset, size, crc32c of the bytes in ``.data-<shard>-of-<n>``.
* key ``_CHECKPOINTABLE_OBJECT_GRAPH`` is a scalar string tensor holding a ``TrackableObjectGraph`` proto
  (trackable_object_graph.proto): nodes with named children, per-variable ``checkpoint_key`` and optimizer slot variables.
  Variables are resolved by WALKING that graph along the reference's attribute names (``unet/encoder/downsamples/0/
  convchain/layer_with_weights-0/kernel`` ...), so the loader does not depend on which of several equivalent paths
  TensorFlow chose for the key string.

Parity note: no TF-written file is available here; the reader is checked against files produced by the writer below
(same published format, uncompressed blocks), against hand-built snappy / CRC32C / varint known answers, and for the
name mapping against the reference's class attributes (components.py:46-61,118-134,203-218,294; unet.py:49-61,152,241-247).
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
OBJECT_GRAPH_KEY = '_CHECKPOINTABLE_OBJECT_GRAPH'
VAR_SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_STRING, _DT_BFLOAT16 = 7, 14
_NP2DT = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


# ---------------------------------------------------------------------------------------------
# primitives: varints, CRC32C (Castagnoli, masked as in lib/hash/crc32c.h), snappy (raw format)
# ---------------------------------------------------------------------------------------------
def read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError('malformed varint')


def write_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _crc_table():
    poly, tab = 0x82f63b78, []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_CRC = _crc_table()


def crc32c(data, crc=0):
    c = (~crc) & 0xffffffff
    tab = _CRC
    for b in bytes(data):
        c = int(tab[(c ^ b) & 0xff]) ^ (c >> 8)
    return (~c) & 0xffffffff


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff


def unmask_crc(m):
    rot = (m - 0xa282ead8) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


def snappy_uncompress(src):
    """Raw snappy block format: varint uncompressed length, then literal / copy elements."""
    src = bytes(src)
    n, pos = read_varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:                                   # copy, 1-byte offset
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:                                 # copy, 2-byte offset
            ln = (tag >> 2) + 1
            off = src[pos] | (src[pos + 1] << 8)
            pos += 2
        else:                                           # copy, 4-byte offset
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('snappy: bad copy offset')
        for _ in range(ln):                             # byte-wise: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError(f'snappy: length {len(out)} != header {n}')
    return bytes(out)


# ---------------------------------------------------------------------------------------------
# protobuf wire format (just what the three messages need)
# ---------------------------------------------------------------------------------------------
def parse_proto(buf):
    """-> list of (field number, wire type, value); value = int (varint / fixed) or bytes (length-delimited)."""
    buf = bytes(buf)
    pos, out = 0, []
    while pos < len(buf):
        key, pos = read_varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = read_varint(buf, pos)
        elif wt == 1:
            v = int.from_bytes(buf[pos:pos + 8], 'little')
            pos += 8
        elif wt == 2:
            ln, pos = read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = int.from_bytes(buf[pos:pos + 4], 'little')
            pos += 4
        else:
            raise ValueError(f'unsupported wire type {wt}')
        out.append((fn, wt, v))
    return out


def _pb_field(fn, wt, payload):
    return write_varint((fn << 3) | wt) + payload


def _pb_varint(fn, v):
    return _pb_field(fn, 0, write_varint(v & 0xffffffffffffffff))


def _pb_bytes(fn, b):
    return _pb_field(fn, 2, write_varint(len(b)) + bytes(b))


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def parse_bundle_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, slices=0)
    for fn, wt, v in parse_proto(buf):
        if fn == 1:
            e['dtype'] = v
        elif fn == 2:
            dims = []
            for f2, _, v2 in parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in parse_proto(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    dims.append(size)
            e['shape'] = tuple(dims)
        elif fn == 3:
            e['shard_id'] = v
        elif fn == 4:
            e['offset'] = v
        elif fn == 5:
            e['size'] = v
        elif fn == 6:
            e['crc32c'] = v
        elif fn == 7:
            e['slices'] += 1
    return e


def parse_object_graph(buf):
    """TrackableObjectGraph -> list of nodes {children: {local_name: node_id}, attributes: {name: (full_name, key)},
    slots: [(original_variable_node_id, slot_name, slot_variable_node_id)]}."""
    nodes = []
    for fn, _, v in parse_proto(buf):
        if fn != 1:
            continue
        node = dict(children=OrderedDict(), attributes=OrderedDict(), slots=[])
        for f2, _, v2 in parse_proto(v):
            if f2 == 1:
                nid, name = 0, ''
                for f3, _, v3 in parse_proto(v2):
                    if f3 == 1:
                        nid = v3
                    elif f3 == 2:
                        name = v3.decode()
                node['children'][name] = nid
            elif f2 == 2:
                name = full = key = ''
                for f3, _, v3 in parse_proto(v2):
                    if f3 == 1:
                        name = v3.decode()
                    elif f3 == 2:
                        full = v3.decode()
                    elif f3 == 3:
                        key = v3.decode()
                node['attributes'][name] = (full, key)
            elif f2 == 3:
                orig = slot = 0
                sname = ''
                for f3, _, v3 in parse_proto(v2):
                    if f3 == 1:
                        orig = v3
                    elif f3 == 2:
                        sname = v3.decode()
                    elif f3 == 3:
                        slot = v3
                node['slots'].append((orig, sname, slot))
        nodes.append(node)
    return nodes


# ---------------------------------------------------------------------------------------------
# the sorted string table (.index)
# ---------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify=True):
    raw = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack('<I', data[offset + size + 1:offset + size + 5])[0]
        if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
            raise ValueError(f'block at {offset}: CRC mismatch')
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_uncompress(raw)
    raise ValueError(f'block at {offset}: unknown compression type {ctype}')


def _block_entries(block):
    nrestarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key, out = 0, b'', []
    while pos < end:
        shared, pos = read_varint(block, pos)
        non_shared, pos = read_varint(block, pos)
        vlen, pos = read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    data = open(path, 'rb').read()
    if len(data) < 48:
        raise ValueError(f'{path}: too short for a table footer')
    footer = data[-48:]
    if struct.unpack('<Q', footer[40:])[0] != TABLE_MAGIC:
        raise ValueError(f'{path}: not a TensorFlow checkpoint index (bad table magic)')
    pos = 0
    _, pos = read_varint(footer, pos)          # metaindex handle
    _, pos = read_varint(footer, pos)
    ioff, pos = read_varint(footer, pos)       # index handle
    isize, pos = read_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, p = read_varint(handle, 0)
        bsize, p = read_varint(handle, p)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


class CheckpointReader:
    """``tf.train.load_checkpoint(prefix)`` without TensorFlow: ``keys()``, ``get_tensor(key)``, ``object_graph()``."""

    def __init__(self, prefix, verify=True):
        if prefix.endswith('.index'):
            prefix = prefix[:-len('.index')]
        self.prefix, self.verify = prefix, verify
        self.entries = OrderedDict()
        self.num_shards = 1
        for k, v in read_table(prefix + '.index', verify):
            if k == b'':
                for fn, _, val in parse_proto(v):
                    if fn == 1:
                        self.num_shards = val
                    elif fn == 2 and val != 0:
                        raise ValueError('big-endian checkpoints are not supported')
                continue
            self.entries[k.decode()] = parse_bundle_entry(v)
        self._shards = {}

    def keys(self):
        return list(self.entries)

    def _shard(self, i):
        if i not in self._shards:
            self._shards[i] = np.memmap(f'{self.prefix}.data-{i:05d}-of-{self.num_shards:05d}', dtype=np.uint8, mode='r')
        return self._shards[i]

    def _bytes(self, e):
        raw = bytes(self._shard(e['shard_id'])[e['offset']:e['offset'] + e['size']])
        if self.verify and e['crc32c'] is not None and e['dtype'] != _DT_STRING and unmask_crc(e['crc32c']) != crc32c(raw):
            raise ValueError('tensor bytes: CRC mismatch')
        return raw

    def get_tensor(self, key):
        e = self.entries[key]
        if e['slices']:
            raise NotImplementedError(f'{key}: partitioned (sliced) variables are not supported')
        raw = self._bytes(e)
        if e['dtype'] == _DT_STRING:
            # [varint64 length]*N, 4-byte masked CRC32C of the lengths, then the bytes
            n = int(np.prod(e['shape'])) if e['shape'] else 1
            pos, lens = 0, []
            for _ in range(n):
                ln, pos = read_varint(raw, pos)
                lens.append(ln)
            pos += 4
            vals = []
            for ln in lens:
                vals.append(raw[pos:pos + ln])
                pos += ln
            return vals[0] if not e['shape'] else np.array(vals, dtype=object).reshape(e['shape'])
        if e['dtype'] == _DT_BFLOAT16:
            a = np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16
            return a.view(np.float32).reshape(e['shape'])
        if e['dtype'] not in _DTYPES:
            raise NotImplementedError(f'{key}: dtype enum {e["dtype"]}')
        return np.frombuffer(raw, dtype=_DTYPES[e['dtype']]).reshape(e['shape']).copy()

    def object_graph(self):
        if OBJECT_GRAPH_KEY not in self.entries:
            return None
        return parse_object_graph(self.get_tensor(OBJECT_GRAPH_KEY))

    # ---- object-graph walking -----------------------------------------------------------------
    def resolve(self, path, nodes=None):
        """Checkpoint key of the variable reached from the root along ``path`` (a list of child names), or None."""
        nodes = nodes if nodes is not None else self.object_graph()
        if nodes is None:
            return None
        nid = 0
        for name in path:
            ch = nodes[nid]['children']
            if name not in ch:
                return None
            nid = ch[name]
        attr = nodes[nid]['attributes'].get('VARIABLE_VALUE')
        return (attr[1], nid) if attr else None


# ---------------------------------------------------------------------------------------------
# name mapping: this package's variable names -> object-graph paths of the reference's classes
# ---------------------------------------------------------------------------------------------
_LEAF = {'kernel': 'kernel', 'bias': 'bias', 'gamma': 'gamma', 'beta': 'beta', 'moving_mean': 'moving_mean',
         'moving_var': 'moving_variance'}


def reference_paths(model):
    """{our variable name: [candidate paths]} for UNetAnnotator / MulmoUNetAnnotator built from the reference's attribute
    names: ``unet`` / ``last_conv`` (unet.py:245-247), ``encoder`` / ``encoders`` / ``decoder`` (unet.py:49-61,152),
    ``downsamples`` / ``upsamples`` lists (components.py:203,276), ``convchain`` Sequential of conv[, bn] pairs, ``pool``
    Sequential [MaxPool, BN], ``conv_transpose`` layer or Sequential [ConvT, BN] (components.py:46-61,118-134).
    A ``keras.Sequential`` names its children ``layer_with_weights-<i>`` (layers that own variables) and ``layer-<i>``."""
    if type(model).__name__ not in ('UNetAnnotator', 'MulmoUNetAnnotator'):
        # MultiResUnet is a Keras FUNCTIONAL model: its checkpoint keys are `layer_with_weights-<i>` in the order Keras'
        # graph traversal assigns, which is not restated here (nothing to pin it against) -- use the .npz format for it
        raise NotImplementedError(f'TensorFlow-format checkpoints are mapped for UNetAnnotator / MulmoUNetAnnotator only, not '
                                  f'{type(model).__name__}: save / load its weights in the .npz format')
    bn = bool(model.configs.get('bn'))
    out = {}
    for name in model.params.specs:
        parts = name.split('/')
        leaf = _LEAF[parts[-1]]
        if parts[0] == 'head':
            out[name] = [['last_conv', leaf]]
            continue
        if parts[0] == 'enc':
            if parts[1].startswith('d'):                       # UNet: enc/d<i>/...
                base, rest = ['unet', 'encoder', 'downsamples', parts[1][1:]], parts[2:]
            else:                                              # MulmoUNet: enc/<m>/d<i>/...
                base, rest = ['unet', 'encoders', parts[1], 'downsamples', parts[2][1:]], parts[3:]
        elif parts[0] == 'dec':
            base, rest = ['unet', 'decoder', 'upsamples', parts[1][1:]], parts[2:]
        else:
            raise KeyError(f'no reference path known for variable {name}')
        layer = rest[0]
        cands = []
        if layer.startswith('conv'):
            k = int(layer[4:])
            j = 2 * k if bn else k
            cands = [base + ['convchain', f'layer_with_weights-{j}', leaf], base + ['convchain', f'layer-{j}', leaf]]
            if parts[0] == 'dec':
                cands.append(base + ['conv_layers', str(j), leaf])
        elif layer.startswith('bn'):
            k = int(layer[2:])
            cands = [base + ['convchain', f'layer_with_weights-{2 * k + 1}', leaf], base + ['convchain', f'layer-{2 * k + 1}', leaf]]
            if parts[0] == 'enc':
                cands.append(base + ['batchnorms', str(k), leaf])
        elif layer == 'pool_bn':
            cands = [base + ['pool', 'layer_with_weights-0', leaf], base + ['pool', 'layer-1', leaf]]
        elif layer == 'tconv':
            cands = ([base + ['conv_transpose', 'layer_with_weights-0', leaf], base + ['conv_transpose', 'layer-0', leaf]]
                     if bn else [base + ['conv_transpose', leaf]])
        elif layer == 'tconv_bn':
            cands = [base + ['conv_transpose', 'layer_with_weights-1', leaf], base + ['conv_transpose', 'layer-1', leaf]]
        else:
            raise KeyError(f'no reference path known for variable {name}')
        out[name] = cands
    return out


def load_into(model, prefix, verify=True):
    """Loads a TF2 object-based checkpoint written by the reference into ``model`` (variables, and the Adam slots /
    iteration count when the file holds them).  Returns (loaded names, missing names, unused checkpoint keys)."""
    rd = CheckpointReader(prefix, verify)
    nodes = rd.object_graph()
    if nodes is None:
        raise ValueError(f'{prefix}: no {OBJECT_GRAPH_KEY} entry (a TF1 name-based checkpoint?)')
    weights, node_of, missing, used = {}, {}, [], set()
    for name, cands in reference_paths(model).items():
        hit = None
        for path in cands:
            hit = rd.resolve(path, nodes)
            if hit:
                break
        if not hit:
            missing.append(name)
            continue
        key, nid = hit
        weights[name] = rd.get_tensor(key)
        node_of[nid] = name
        used.add(key)
    model.set_weights(weights)
    # optimizer: root/optimizer -> iter, and slot variables (m, v) keyed by the node of the variable they belong to
    opt = nodes[0]['children'].get('optimizer')
    slots = {}
    if opt is not None:
        for orig, sname, slot_node in nodes[opt]['slots']:
            attr = nodes[slot_node]['attributes'].get('VARIABLE_VALUE')
            if attr and orig in node_of and sname in ('m', 'v'):
                slots[(node_of[orig], sname)] = rd.get_tensor(attr[1])
                used.add(attr[1])
        it = nodes[opt]['children'].get('iter')
        if it is not None and 'VARIABLE_VALUE' in nodes[it]['attributes']:
            k = nodes[it]['attributes']['VARIABLE_VALUE'][1]
            slots['iter'] = int(np.asarray(rd.get_tensor(k)).reshape(-1)[0])
            used.add(k)
    if slots and model.params.device is not None:
        import torch
        ps = model.params
        for key, arr in slots.items():
            if key == 'iter':
                ps.step.fill_(int(arr))
                continue
            name, sname = key
            s = ps.specs[name]
            if not s['trainable']:
                continue
            flat = ps.m if sname == 'm' else ps.v
            flat[s['offset']:s['offset'] + s['numel']].copy_(torch.from_numpy(np.ascontiguousarray(arr, np.float32).ravel()))
    unused = [k for k in rd.keys() if k not in used and k != OBJECT_GRAPH_KEY]
    return sorted(weights), missing, unused


# ---------------------------------------------------------------------------------------------
# writer (uncompressed blocks): export this package's variables as a checkpoint the reference's load_weights accepts
# ---------------------------------------------------------------------------------------------
def _build_block(entries, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += write_varint(shared) + write_varint(len(k) - shared) + write_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b'\x00')
    f.write(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
    return off, len(block)


def write_table(path, items, block_entries=64):
    """``items``: (key bytes, value bytes) sorted by key."""
    items = list(items)
    assert all(items[i][0] < items[i + 1][0] for i in range(len(items) - 1)), 'keys must be strictly increasing'
    with open(path, 'wb') as f:
        index = []
        for i in range(0, max(len(items), 1), block_entries):
            chunk = items[i:i + block_entries]
            off, size = _emit_block(f, _build_block(chunk))
            index.append(((chunk[-1][0] if chunk else b''), write_varint(off) + write_varint(size)))
        moff, msize = _emit_block(f, _build_block([]))
        ioff, isize = _emit_block(f, _build_block(index, restart_interval=1))
        footer = write_varint(moff) + write_varint(msize) + write_varint(ioff) + write_varint(isize)
        f.write(footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC))


def _entry_proto(dtype, shape, offset, size, crc):
    shp = b''.join(_pb_bytes(2, _pb_varint(1, d)) for d in shape)
    return _pb_varint(1, dtype) + _pb_bytes(2, shp) + _pb_varint(4, offset) + _pb_varint(5, size) + _pb_field(6, 5, struct.pack('<I', crc))


def write_checkpoint(prefix, tensors, graph_nodes=None):
    """``tensors``: {checkpoint key: ndarray}; ``graph_nodes``: the object graph as ``parse_object_graph`` returns it."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or '.', exist_ok=True)
    items, data = [], bytearray()
    if graph_nodes is not None:
        g = b''
        for node in graph_nodes:
            body = b''
            for name, nid in node['children'].items():
                body += _pb_bytes(1, _pb_varint(1, nid) + _pb_bytes(2, name.encode()))
            for name, (full, key) in node['attributes'].items():
                body += _pb_bytes(2, _pb_bytes(1, name.encode()) + _pb_bytes(2, full.encode()) + _pb_bytes(3, key.encode()))
            for orig, sname, slot in node['slots']:
                body += _pb_bytes(3, _pb_varint(1, orig) + _pb_bytes(2, sname.encode()) + _pb_varint(3, slot))
            g += _pb_bytes(1, body)
        lens = write_varint(len(g))
        blob = lens + struct.pack('<I', mask_crc(crc32c(lens))) + g
        crc = crc32c(g, crc32c(lens))
        items.append((OBJECT_GRAPH_KEY.encode(), _entry_proto(_DT_STRING, (), len(data), len(blob), mask_crc(crc))))
        data += blob
    for key in sorted(tensors):
        a = np.ascontiguousarray(tensors[key])
        if a.dtype not in _NP2DT:
            a = a.astype(np.float32)
        raw = a.tobytes()
        items.append((key.encode(), _entry_proto(_NP2DT[a.dtype], a.shape, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    items.sort(key=lambda kv: kv[0])
    header = _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1))          # num_shards = 1, little endian, version.producer = 1
    write_table(prefix + '.index', [(b'', header)] + items)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        f.write(bytes(data))
    return prefix


def export_from(model, prefix):
    """Writes ``model``'s variables (and Adam slots / iteration when materialised) as a TF2 object-based checkpoint whose
    object graph follows the reference's attribute names, i.e. the file ``engine.py:75`` / ``:230`` loads."""
    paths = reference_paths(model)
    w = model.get_weights()
    nodes = [dict(children=OrderedDict(), attributes=OrderedDict(), slots=[])]

    def node_for(path):
        nid = 0
        for name in path:
            ch = nodes[nid]['children']
            if name not in ch:
                nodes.append(dict(children=OrderedDict(), attributes=OrderedDict(), slots=[]))
                ch[name] = len(nodes) - 1
            nid = ch[name]
        return nid
    tensors, var_node = {}, {}
    for name, cands in paths.items():
        path = cands[0]
        nid = node_for(path)
        key = '/'.join(path) + VAR_SUFFIX
        nodes[nid]['attributes']['VARIABLE_VALUE'] = (path[-1], key)
        tensors[key] = w[name]
        var_node[name] = nid
    ps = model.params
    if ps.device is not None:
        opt = node_for(['optimizer'])
        it = node_for(['optimizer', 'iter'])
        nodes[it]['attributes']['VARIABLE_VALUE'] = ('Adam/iter', 'optimizer/iter' + VAR_SUFFIX)
        tensors['optimizer/iter' + VAR_SUFFIX] = np.asarray(int(ps.step.item()), dtype=np.int64)
        m, v = ps.m.cpu().numpy(), ps.v.cpu().numpy()
        for name, s in ps.specs.items():
            if not s['trainable']:
                continue
            for sname, flat in (('m', m), ('v', v)):
                nodes.append(dict(children=OrderedDict(), attributes=OrderedDict(), slots=[]))
                sid = len(nodes) - 1
                key = '/'.join(paths[name][0]) + f'/.OPTIMIZER_SLOT/optimizer/{sname}' + VAR_SUFFIX
                nodes[sid]['attributes']['VARIABLE_VALUE'] = (f'Adam/{name}/{sname}', key)
                nodes[opt]['slots'].append((var_node[name], sname, sid))
                tensors[key] = flat[s['offset']:s['offset'] + s['numel']].reshape(s['shape'])
    return write_checkpoint(prefix, tensors, nodes)
