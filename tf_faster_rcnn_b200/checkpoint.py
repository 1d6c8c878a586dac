"""Reads (and, for tests / conversion, writes) TensorFlow "V2" checkpoints -- the `<prefix>.index` +
`<prefix>.data-SSSSS-of-NNNNN` pair that the reference restores with `tf.train.Saver().restore(sess, model)`
(tools/demo.py:139-140, tools/test_net.py:111-113) -- without TensorFlow.

Format (restated from the published LevelDB table format and TensorFlow's tensor_bundle.proto; no TensorFlow exists in
this container, so the reader is checked against this module's own writer, the CRC-32C check value and the table magic,
NOT against a TensorFlow-written file -- see tests/test_checkpoint.py):

  .index   an immutable sorted string table:
             [data block]* [metaindex block] [index block] [footer: 2 block handles padded to 40 B + 8 B magic]
           block  = entries (varint shared, varint non_shared, varint value_len, key suffix, value), then the restart
                    offsets (u32 each) and their count (u32); followed on disk by a 1-byte compression type
                    (0 none, 1 snappy) and the masked CRC-32C (u32) of contents + type.
           key "" -> BundleHeaderProto {num_shards=1, endianness=2, version=3};
           key <variable name> -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32,
           masked CRC-32C of the tensor bytes), slices=7 (partitioned variables; not supported here)}.
  .data-*  raw little-endian tensor bytes at [offset, offset + size) of shard `shard_id`.

Only numeric dtypes are decoded (Faster R-CNN checkpoints hold fp32 weights plus int counters); anything else is skipped
with its name reported in `skipped`."""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_FOOTER_LEN = 48
_BLOCK_TRAILER = 5

# tensorflow/core/framework/types.proto enum values -> numpy
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class CheckpointError(IOError):
    pass


# ---- CRC-32C (Castagnoli), the checksum of both the table blocks and the tensor payloads -------------------------
def _make_table():
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        t[i] = c
    return t


_CRC_TABLE = _make_table()
_CRC_LIST = [int(v) for v in _CRC_TABLE]


def _advance(state, data_cols):
    """state: uint32 [n]; data_cols: uint8 [steps, n].  One table-driven byte step per row, all lanes at once."""
    for row in data_cols:
        state = _CRC_TABLE[(state ^ row) & 0xFF] ^ (state >> 8)
    return state


def _crc_update_scalar(state, data):
    tab = _CRC_LIST
    for b in data:
        state = tab[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


_CHUNK = 4096
_shift_tables = {}


def _shift_table(nbytes):
    """Lookup tables of the linear map "run the register over `nbytes` zero bytes", one 256-entry table per state byte."""
    tabs = _shift_tables.get(nbytes)
    if tabs is None:
        seeds = np.concatenate([np.arange(256, dtype=np.uint32) << np.uint32(8 * b) for b in range(4)])
        out = _advance(seeds, np.zeros((nbytes, seeds.size), dtype=np.uint8))
        tabs = [[int(v) for v in out[256 * b:256 * (b + 1)]] for b in range(4)]
        _shift_tables[nbytes] = tabs
    return tabs


def crc32c(data):
    """CRC-32C of a bytes-like object.  Large inputs are cut into 4 KiB chunks whose registers advance in lock-step as
    numpy lanes; chunk results are stitched with the zero-advance operator (CRC is affine in its start state)."""
    buf = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8)
    n = buf.size
    state = 0xFFFFFFFF
    nchunks = n // _CHUNK
    if nchunks >= 8:
        body = buf[:nchunks * _CHUNK].reshape(nchunks, _CHUNK)
        partial = _advance(np.zeros(nchunks, dtype=np.uint32), np.ascontiguousarray(body.T))
        t0, t1, t2, t3 = _shift_table(_CHUNK)
        for g in partial.tolist():
            state = t0[state & 0xFF] ^ t1[(state >> 8) & 0xFF] ^ t2[(state >> 16) & 0xFF] ^ t3[state >> 24] ^ g
        rest = buf[nchunks * _CHUNK:]
    else:
        rest = buf
    state = _crc_update_scalar(state, rest.tolist())
    return state ^ 0xFFFFFFFF


def mask_crc(crc):
    """LevelDB/TensorFlow store CRCs rotated and offset so that a CRC of data containing CRCs stays well distributed."""
    return (((crc >> 15) | (crc << 17)) + 0xa282ead8) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - 0xa282ead8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format -------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    result = 0
    shift = 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _proto_fields(buf):
    """Yields (field_number, wire_type, value) of one protobuf message; value is an int or a bytes slice."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _get_varint(buf, pos)
        elif wire == 1:
            val, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wire == 2:
            ln, pos = _get_varint(buf, pos)
            if pos + ln > len(buf):
                raise CheckpointError("truncated protobuf field")
            val, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wire == 5:
            val, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wire)
        yield field, wire, val


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_shape(buf):
    dims = []
    for field, _, val in _proto_fields(buf):
        if field == 2:                                   # TensorShapeProto.dim
            size = 0
            for f2, _, v2 in _proto_fields(val):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
        elif field == 3 and val:
            raise CheckpointError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": 0}
    for field, _, val in _proto_fields(buf):
        if field == 1:
            e["dtype"] = val
        elif field == 2:
            e["shape"] = _parse_shape(val)
        elif field == 3:
            e["shard_id"] = val
        elif field == 4:
            e["offset"] = val
        elif field == 5:
            e["size"] = val
        elif field == 6:
            e["crc32c"] = val
        elif field == 7:
            e["slices"] += 1
    return e


def _parse_header(buf):
    h = {"num_shards": 0, "endianness": 0, "producer": 0}
    for field, _, val in _proto_fields(buf):
        if field == 1:
            h["num_shards"] = val
        elif field == 2:
            h["endianness"] = val
        elif field == 3:
            for f2, _, v2 in _proto_fields(val):
                if f2 == 1:
                    h["producer"] = v2
    return h


# ---- snappy (block compression type 1; TensorFlow writes index tables uncompressed, other writers may not) ------
def _snappy_uncompress(src):
    total, pos = _get_varint(src, 0)
    out = bytearray()
    n = len(src)
    while pos < n:
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                extra = ln - 59
                ln = int.from_bytes(src[pos:pos + extra], "little")
                pos += extra
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = src[pos] | (src[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise CheckpointError("corrupt snappy block")
        for _ in range(ln):                              # copies may overlap their own output
            out.append(out[-off])
    if len(out) != total:
        raise CheckpointError("corrupt snappy block (length)")
    return bytes(out)


# ---- table reading ---------------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    end = offset + size + _BLOCK_TRAILER
    if end > len(data):
        raise CheckpointError("block handle points outside the index file")
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(data[offset:offset + size + 1]):
            raise CheckpointError("index block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        contents = _snappy_uncompress(contents)
    elif ctype != 0:
        raise CheckpointError("unknown block compression type %d" % ctype)
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("block too small")
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    if limit < 0:
        raise CheckpointError("bad restart count")
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt block entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_table(data, verify):
    if len(data) < _FOOTER_LEN:
        raise CheckpointError("index file shorter than a table footer")
    footer = data[-_FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise CheckpointError("not a TensorFlow V2 checkpoint index (bad table magic)")
    _, pos = _get_varint(footer, 0)                      # metaindex handle (unused: bundles carry no filter)
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        for kv in _block_entries(_read_block(data, off, size, verify)):
            yield kv


def _shard_path(prefix, shard, num_shards):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def is_bundle(prefix):
    return os.path.isfile(prefix + ".index")


def list_variables(prefix, verify=True):
    """[(name, dtype code, shape)] in key order, like tf.train.list_variables."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    return [(k.decode("utf-8"), e["dtype"], e["shape"]) for k, e in
            ((k, _parse_entry(v)) for k, v in _read_table(data, verify) if k != b"")]


def read_bundle(prefix, verify=True, names=None, skipped=None):
    """{variable name: ndarray} for every numeric tensor of the checkpoint `prefix` (what `Saver.restore` would assign).
    `names`: optional predicate/collection to restrict loading; `skipped`: optional list receiving (name, reason)."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    entries = []
    header = None
    for k, v in _read_table(data, verify):
        if k == b"":
            header = _parse_header(v)
        else:
            entries.append((k.decode("utf-8"), _parse_entry(v)))
    if header is None:
        raise CheckpointError("checkpoint index has no bundle header")
    if header["endianness"] != 0:
        raise CheckpointError("big-endian checkpoints are not supported")
    want = (lambda n: True) if names is None else names if callable(names) else (lambda n, s=set(names): n in s)
    shards = {}
    out = {}
    for name, e in entries:
        if not want(name):
            continue
        dt = _DTYPES.get(e["dtype"])
        if dt is None or e["slices"]:
            if skipped is not None:
                skipped.append((name, "partitioned variable" if e["slices"] else "dtype %d" % e["dtype"]))
            continue
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if count * np.dtype(dt).itemsize != e["size"]:
            raise CheckpointError("%s: %d bytes stored for shape %s" % (name, e["size"], e["shape"]))
        mm = shards.get(e["shard_id"])
        if mm is None:
            path = _shard_path(prefix, e["shard_id"], max(header["num_shards"], 1))
            if not os.path.isfile(path):
                raise CheckpointError("missing data shard %s" % path)
            mm = shards[e["shard_id"]] = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else \
                np.zeros(0, np.uint8)
        if e["offset"] + e["size"] > mm.size:
            raise CheckpointError("%s: data shard too short" % name)
        raw = np.array(mm[e["offset"]:e["offset"] + e["size"]])
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise CheckpointError("%s: tensor checksum mismatch" % name)
        out[name] = raw.view(dt).reshape(e["shape"])
    return out


def load_variables(path, verify=True):
    """The lookup `Saver.restore` performs here: a TF V2 bundle at `path` if `<path>.index` exists, else `<path>.npz`
    (or `path` itself when it already names an .npz)."""
    if not path.endswith(".npz") and is_bundle(path):
        return read_bundle(path, verify=verify)
    npz = path if path.endswith(".npz") else path + ".npz"
    if not os.path.isfile(npz):
        raise IOError("no weights at %s: expected a TensorFlow V2 checkpoint (%s.index + .data-*) or %s"
                      % (path, path, npz))
    with np.load(npz) as z:
        return {k: z[k] for k in z.files}


# ---- writing (tests, tools/make_synthetic_ckpt.py --format bundle) --------------------------------------------------
def _field(num, wire, payload):
    return _put_varint((num << 3) | wire) + payload


def _message(num, payload):
    return _field(num, 2, _put_varint(len(payload)) + payload)


def _entry_proto(dtype_code, shape, offset, size, crc_masked):
    shape_msg = b"".join(_message(2, _field(1, 0, _put_varint(int(d)))) for d in shape)
    msg = _field(1, 0, _put_varint(dtype_code)) + _message(2, shape_msg)
    if offset:
        msg += _field(4, 0, _put_varint(offset))
    if size:
        msg += _field(5, 0, _put_varint(size))
    return msg + _field(6, 5, struct.pack("<I", crc_masked))


class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count and self.count % self.interval == 0:
            self.restarts.append(len(self.buf))
        elif self.count:
            lim = min(len(key), len(self.last))
            while shared < lim and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _emit_block(out, contents):
    offset = len(out)
    out += contents + b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))
    return _put_varint(offset) + _put_varint(len(contents))


def write_bundle(prefix, tensors, block_size=4096):
    """Writes {name: array} as a single-shard V2 checkpoint (`prefix.index`, `prefix.data-00000-of-00001`)."""
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    index_pairs = [(b"", _field(1, 0, _put_varint(1)) + _message(3, _field(1, 0, _put_varint(1))))]
    offset = 0
    with open(_shard_path(prefix, 0, 1), "wb") as f:
        for name in names:
            arr = np.asarray(tensors[name])
            code = _DTYPE_CODES.get(arr.dtype)
            if code is None:
                raise CheckpointError("%s: dtype %s cannot be stored" % (name, arr.dtype))
            raw = (np.ascontiguousarray(arr) if arr.ndim else arr.reshape(1)).tobytes()
            f.write(raw)
            index_pairs.append((name.encode("utf-8"), _entry_proto(code, arr.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)
    block = _BlockBuilder()
    for key, value in index_pairs:
        block.add(key, value)
        if len(block.buf) >= block_size:
            index.add(block.last, _emit_block(out, block.finish()))
            block = _BlockBuilder()
    if block.count:
        index.add(block.last, _emit_block(out, block.finish()))
    meta_handle = _emit_block(out, _BlockBuilder().finish())
    index_handle = _emit_block(out, index.finish())
    footer = meta_handle + index_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def _main(argv=None):
    """python -m tf_faster_rcnn_b200.checkpoint list <prefix> | to-npz <prefix> <out.npz> | to-bundle <in.npz> <prefix>"""
    import argparse
    ap = argparse.ArgumentParser(description=_main.__doc__)
    ap.add_argument("command", choices=("list", "to-npz", "to-bundle"))
    ap.add_argument("src")
    ap.add_argument("dst", nargs="?")
    ap.add_argument("--no-verify", action="store_true", help="skip CRC-32C verification")
    a = ap.parse_args(argv)
    if a.command == "list":
        for name, code, shape in list_variables(a.src, verify=not a.no_verify):
            print("%-72s %-8s %s" % (name, np.dtype(_DTYPES[code]).name if code in _DTYPES else "dtype%d" % code, list(shape)))
    elif a.command == "to-npz":
        skipped = []
        np.savez(a.dst, **read_bundle(a.src, verify=not a.no_verify, skipped=skipped))
        for name, why in skipped:
            print("skipped %s (%s)" % (name, why))
    else:
        with np.load(a.src) as z:
            write_bundle(a.dst, {k: z[k] for k in z.files})


if __name__ == "__main__":
    _main()
