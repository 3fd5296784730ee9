"""TensorFlow V2 checkpoint (tensor bundle) reader / writer (cmpc_refseg_b200/tf_bundle.py; trainval_model.py:46-56,185-190).

PARITY UNPINNED for the container format itself: no TensorFlow and no real checkpoint exists in this image (the reference's are
behind a Baidu link, README.md:22).  The reader is therefore tested against (1) the CRC32C known-answer vector of RFC 3720, (2) a
table ENCODER written here independently of the product writer, following leveldb's published block layout the way the real
bundle writer uses it -- prefix-compressed keys, restart interval 16, several data blocks, shortest-separator index keys -- and
(3) the product writer's own output."""
import struct

import numpy as np
import pytest
import torch

from cmpc_refseg_b200 import tf_bundle as tb


def test_crc32c_known_answers():
    assert tb.crc32c(b"123456789") == 0xE3069283                       # RFC 3720 B.4 check value
    assert tb.crc32c(b"\x00" * 32) == 0x8A9136AA and tb.crc32c(b"\xff" * 32) == 0x62A8AB43
    rng = np.random.default_rng(0)
    blob = rng.integers(0, 256, 70001, dtype=np.uint8).tobytes()       # bulk (NumPy) path == bytewise path
    slow = 0xFFFFFFFF
    for b in blob:
        slow = (slow >> 8) ^ tb._T0[(slow ^ b) & 0xFF]
    assert tb.crc32c(blob) == slow ^ 0xFFFFFFFF
    assert tb._unmask(tb._mask(0x12345678)) == 0x12345678


def _leveldb_table(kvs, block_size=300, restart_interval=16):
    """an independent SSTable encoder (leveldb TableBuilder semantics): prefix compression inside restart intervals"""
    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append(v & 0x7F | 0x80); v >>= 7
        out.append(v)
        return bytes(out)

    def build_block(items):
        out, restarts, last = bytearray(), [], b""
        for i, (k, v) in enumerate(items):
            if i % restart_interval == 0:
                restarts.append(len(out)); shared = 0
            else:
                shared = 0
                while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                    shared += 1
            out += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
            last = k
        for r in restarts or [0]:
            out += struct.pack("<I", r)
        out += struct.pack("<I", len(restarts or [0]))
        return bytes(out)

    def emit(f, blk):
        off = len(f)
        f += blk + b"\x00" + struct.pack("<I", tb._mask(tb.crc32c(blk + b"\x00")))
        return varint(off) + varint(len(blk))
    f, index, cur, size = bytearray(), [], [], 0
    for k, v in kvs:
        cur.append((k, v)); size += len(k) + len(v)
        if size >= block_size:
            index.append((cur[-1][0], emit(f, build_block(cur)))); cur, size = [], 0
    if cur:
        index.append((cur[-1][0] + b"~", emit(f, build_block(cur))))
    mh = emit(f, build_block([]))
    ih = emit(f, build_block(index))
    foot = mh + ih
    f += foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", tb.MAGIC)
    return bytes(f)


def test_reader_on_an_independently_encoded_bundle(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {f"text_objseg/vis_trans_c{3 + i % 3}_head{1 + i % 5}/{'DW' if i % 2 else 'biases'}/part{i}":
               rng.standard_normal((1 + i % 3, 2 + i % 4)).astype(np.float32) for i in range(57)}
    tensors["text_objseg/Variable_1"] = np.array(123456, np.int32)
    tensors["text_objseg/c5_lateral/DW/Adam"] = rng.standard_normal((1, 1, 6, 4)).astype(np.float64)
    data, kvs = bytearray(), []
    dt = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3}

    def pbv(field, v):
        return tb._put_varint(field << 3) + tb._put_varint(v)

    def pbb(field, b):
        return tb._put_varint(field << 3 | 2) + tb._put_varint(len(b)) + b
    kvs.append((b"", pbv(1, 1) + pbb(3, pbv(1, 1))))
    for name in sorted(tensors, key=lambda s: s.encode()):
        a = tensors[name]
        raw = a.tobytes()
        shape = b"".join(pbb(2, pbv(1, d)) for d in a.shape)
        e = pbv(1, dt[a.dtype]) + pbb(2, shape) + pbv(4, len(data)) + pbv(5, len(raw)) + tb._put_varint(6 << 3 | 5) + \
            struct.pack("<I", tb._mask(tb.crc32c(raw)))
        kvs.append((name.encode(), e)); data += raw
    prefix = str(tmp_path / "model.ckpt-123456")
    (tmp_path / "model.ckpt-123456.index").write_bytes(_leveldb_table(kvs))
    (tmp_path / "model.ckpt-123456.data-00000-of-00001").write_bytes(bytes(data))
    entries, header = tb.read_index(prefix)
    assert header["num_shards"] == 1 and set(entries) == set(tensors)
    back = tb.read_bundle(prefix, verify=True)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    # corruption is detected: flip one byte of a tensor / of the index
    raw = bytearray((tmp_path / "model.ckpt-123456.data-00000-of-00001").read_bytes()); raw[5] ^= 1
    (tmp_path / "model.ckpt-123456.data-00000-of-00001").write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        tb.read_bundle(prefix, verify=True)
    idx = bytearray((tmp_path / "model.ckpt-123456.index").read_bytes()); idx[10] ^= 1
    (tmp_path / "model.ckpt-123456.index").write_bytes(bytes(idx))
    with pytest.raises(ValueError):
        tb.read_index(prefix)


def test_checkpoint_prefix_round_trip_through_load_variables(tmp_path):
    """save_variables / load_variables on a tf.train.Saver-style prefix: scope text_objseg/, optimizer slots and backbone ignored"""
    from cmpc_refseg_b200.CMPC_model import head_param_shapes, reference_init
    from cmpc_refseg_b200.checkpoint import TF_SCOPE, load_variables, save_variables
    kw = dict(vf_h=4, vf_w=4, vf_dim=32, v_emb_dim=16, rnn_size=16, mlp_dim=8, c4_dim=16, c3_dim=8, parse_hidden=12)
    shapes = head_param_shapes(**kw)
    params = reference_init(shapes, seed=3)
    prefix = str(tmp_path / "model.ckpt-42")
    save_variables(prefix, params, global_step=42)
    back = load_variables(prefix, shapes, verify=True)
    assert set(back) == set(shapes) and all(torch.equal(back[k], params[k]) for k in shapes)
    assert load_variables.last_ignored == [TF_SCOPE + "Variable_1"]
    # a full reference checkpoint also holds Adam slots, the backbone and the word encoder
    arrs = {TF_SCOPE + k: v.numpy() for k, v in params.items()}
    arrs[TF_SCOPE + "c5_lateral/DW/Adam"] = np.zeros((1, 1, 32, 16), np.float32)
    arrs["res5c_branch2c/weights"] = np.zeros((1, 1, 4, 4), np.float32)
    arrs[TF_SCOPE + "rnn/lstm_cell/kernel"] = np.ones((4, 4), np.float32)
    arrs["beta1_power"] = np.array(0.9, np.float32)
    tb.write_bundle(str(tmp_path / "full"), arrs)
    back = load_variables(str(tmp_path / "full"), shapes)
    assert set(back) == set(shapes) | {"rnn/lstm_cell/kernel"} and len(load_variables.last_ignored) == 3
