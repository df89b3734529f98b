"""TensorFlow checkpoint interop: the tensor-bundle reader / writer of tf_bundle.py and the
import / export of engine state under the reference's variable names (model.py:689-699,758-764,1138-1139).

No TensorFlow exists here, so what is pinned is what can be: the CRC against its published check values and
against TensorBoard's TF-compatible record CRC, the nested protos against the TF .proto classes TensorBoard
ships, and the table / bundle layer through its own invariants (sorted keys, prefix sharing, restart arrays,
block and tensor checksums, the footer)."""
import os
import struct

import numpy as np
import pytest

from vnet_tensorflow_b200 import _ffi, checkpoint, tf_bundle as tb
from tests.helpers import engine_for, perturbed_params


# ---- CRC-32C ---------------------------------------------------------------------------------------------------
def test_crc32c_known_answers():
    assert tb.crc32c(b"") == 0
    assert tb.crc32c(b"123456789") == 0xE3069283                      # the CRC catalogue's check value
    assert tb.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4
    assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    assert tb.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    # tensorflow/core/lib/hash/crc32c_test.cc: Mask differs from the CRC, is not an involution, and Unmask inverts it
    c = tb.crc32c(b"foo")
    assert tb.crc_mask(c) != c and tb.crc_mask(tb.crc_mask(c)) != c
    assert tb.crc_unmask(tb.crc_mask(c)) == c and tb.crc_unmask(tb.crc_unmask(tb.crc_mask(tb.crc_mask(c)))) == c


def test_crc32c_lockstep_equals_the_byte_loop_and_tensorboards_crc():
    stub = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    rng = np.random.default_rng(0)
    for n in (1, 255, 65535, 65536, 65537, 200_003):                   # both sides of the lock-step threshold
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        mine = tb.crc32c(d)
        assert mine == stub.crc32c(d) and tb.crc_mask(mine) == stub.masked_crc32c(d)
        cut = n // 3
        assert tb.crc32c(d[cut:], tb.crc32c(d[:cut])) == mine           # crc32c::Extend
    big = rng.integers(0, 256, 3_000_001, dtype=np.uint8)
    parts = [tb.crc32c(big)]
    r = 0
    for lo in range(0, big.size, 50_000):                               # byte-loop path, piecewise
        r = tb.crc32c(big[lo:lo + 50_000], r)
    parts.append(r)
    assert parts[0] == parts[1]


# ---- protos ----------------------------------------------------------------------------------------------------
def test_shape_and_version_protos_match_tensorflows_definitions():
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    versions_pb2 = pytest.importorskip("tensorboard.compat.proto.versions_pb2")
    for dims in [(), (1,), (5, 5, 5, 16, 32), (0, 3), (2 ** 40, 7)]:
        ref = shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=d) for d in dims])
        assert tb.encode_shape(dims) == ref.SerializeToString()
        assert tb.decode_shape(ref.SerializeToString()) == dims
    header = tb.encode_header(1)
    fields = list(tb._fields(header))
    assert [(n, w) for n, w, _ in fields] == [(1, 0), (3, 2)] and fields[0][2] == 1
    assert fields[1][2] == versions_pb2.VersionDef(producer=1).SerializeToString()
    assert tb.decode_header(header) == {"num_shards": 1, "endianness": 0, "producer": 1}


def _bundle_entry_class():
    """BundleEntryProto as tensor_bundle.proto declares it, built at run time on TF's own TensorShapeProto / DataType."""
    pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="vnb_test/tensor_bundle.proto", package="vnb_test", syntax="proto3",
                                            dependency=[tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name])
    m = fd.message_type.add(name="BundleEntryProto")
    m.field.add(name="dtype", number=1, type=F.TYPE_ENUM, type_name=".tensorboard.DataType", label=F.LABEL_OPTIONAL)
    m.field.add(name="shape", number=2, type=F.TYPE_MESSAGE, type_name=".tensorboard.TensorShapeProto", label=F.LABEL_OPTIONAL)
    m.field.add(name="shard_id", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    m.field.add(name="offset", number=4, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    m.field.add(name="size", number=5, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    m.field.add(name="crc32c", number=6, type=F.TYPE_FIXED32, label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.Default()
    try:
        desc = pool.FindMessageTypeByName("vnb_test.BundleEntryProto")
    except KeyError:
        pool.AddSerializedFile(fd.SerializeToString())
        desc = pool.FindMessageTypeByName("vnb_test.BundleEntryProto")
    return message_factory.GetMessageClass(desc), types_pb2


def test_bundle_entry_wire_format_against_protobuf():
    cls, types_pb2 = _bundle_entry_class()
    assert (types_pb2.DT_FLOAT, types_pb2.DT_INT32, types_pb2.DT_INT64, types_pb2.DT_DOUBLE, types_pb2.DT_HALF) == (1, 3, 9, 2, 19)
    for dtype, shape, shard, off, size, crc in [(1, (5, 5, 5, 16, 32), 0, 0, 256000, 0xDEADBEEF), (9, (), 0, 123456789012, 8, 7),
                                                (3, (1,), 2, 300, 4, 0x80000000)]:
        e = tb.BundleEntry(dtype, shape, shard, off, size, crc)
        msg = cls()
        msg.ParseFromString(e.encode())
        assert (msg.dtype, tuple(d.size for d in msg.shape.dim), msg.shard_id, msg.offset, msg.size, msg.crc32c) == \
               (dtype, shape, shard, off, size, crc)
        msg.shape.SetInParent()  # BundleWriter::Add always touches mutable_shape()
        assert msg.SerializeToString() == e.encode()
        back = tb.BundleEntry.decode(msg.SerializeToString())
        assert (back.dtype, back.shape, back.shard_id, back.offset, back.size, back.crc32c) == (dtype, shape, shard, off, size, crc)


# ---- table -----------------------------------------------------------------------------------------------------
def test_table_layout_single_block(tmp_path):
    items = [(b"", b"H"), (b"vnet/a/weights", b"V1"), (b"vnet/a/weights/Adam", b"V2"), (b"vnet/b", b"V3")]
    path = str(tmp_path / "t.index")
    tb.write_table(path, items)
    buf = open(path, "rb").read()
    # footer: 40 bytes of handles + padding, then the magic number as two little-endian fixed32 halves
    assert buf[-8:] == struct.pack("<II", 0x8B80FB57, 0xDB477524)
    assert list(tb.read_table(path).items()) == items
    # data block: second key shares no prefix with "", third shares all 14 bytes of the second, fourth shares "vnet/"
    assert buf[:4] == bytes([0, 0, 1]) + b"H"
    assert buf[4:7] == bytes([0, 14, 2]) and buf[23:26] == bytes([14, 5, 2]) and buf[26:31] == b"/Adam"
    assert buf[33:36] == bytes([5, 1, 2]) and buf[36:37] == b"b"
    # one restart point (offset 0), then type byte 0 and the masked CRC of block + type
    assert buf[39:47] == struct.pack("<II", 0, 1) and buf[47] == 0
    assert struct.unpack_from("<I", buf, 48)[0] == tb.crc_mask(tb.crc32c(buf[:48]))
    # metaindex block: empty block = one restart, 8 bytes + 5 trailer
    assert buf[52:60] == struct.pack("<II", 0, 1)
    # index block: one entry whose key is the short successor of the last key ("w") and whose value is handle (0, 47)
    assert buf[65:70] == bytes([0, 1, 2]) + b"w" + bytes([0]) and buf[70] == 47


def test_table_many_blocks_restarts_and_corruption(tmp_path):
    rng = np.random.default_rng(1)
    keys = sorted({("scope_%03d/var_%d/%s" % (rng.integers(0, 40), rng.integers(0, 50), s)).encode()
                   for _ in range(600) for s in ("w", "w/Adam", "w/Adam_1")})
    items = [(b"", b"hdr")] + [(k, bytes(rng.integers(0, 256, int(rng.integers(0, 60)), dtype=np.uint8))) for k in keys]
    path = str(tmp_path / "many.index")
    tb.write_table(path, items, block_size=1024)                        # dozens of data blocks, many restart points
    assert list(tb.read_table(path).items()) == items
    buf = bytearray(open(path, "rb").read())
    assert len(buf) > 20 * 1024
    with pytest.raises(ValueError, match="strictly increasing"):
        tb.write_table(str(tmp_path / "bad.index"), [(b"b", b""), (b"a", b"")])
    buf[100] ^= 0x40
    open(path, "wb").write(buf)
    with pytest.raises(ValueError, match="checksum"):
        tb.read_table(path)
    buf[-1] ^= 0xFF
    open(path, "wb").write(buf)
    with pytest.raises(ValueError, match="magic"):
        tb.read_table(path)


def test_separator_rules():
    # leveldb's BytewiseComparator, which the table builder uses for index keys
    assert tb._shortest_separator(b"abc1", b"abd") == b"abc1"          # 'c'+1 == 'd': no room
    assert tb._shortest_separator(b"abc1", b"abz") == b"abd"
    assert tb._shortest_separator(b"ab", b"abc") == b"ab"              # prefix: unchanged
    assert tb._short_successor(b"vnet/x") == b"w" and tb._short_successor(b"\xff\xffa") == b"\xff\xffb"
    assert tb._short_successor(b"\xff") == b"\xff"


# ---- bundle ----------------------------------------------------------------------------------------------------
def test_bundle_roundtrip_dtypes_scalars_and_empty(tmp_path):
    rng = np.random.default_rng(2)
    tensors = {
        "vnet/conv/weights": rng.normal(size=(5, 5, 5, 3, 4)).astype(np.float32),
        "vnet/conv/weights/Adam": rng.normal(size=(5, 5, 5, 3, 4)).astype(np.float32),
        "global_step": np.asarray(1234567890123, np.int64),
        "start_epoch": np.asarray([7], np.int32),
        "training/beta1_power": np.asarray(0.9 ** 5, np.float32),
        "big": rng.normal(size=(70_000,)).astype(np.float64),           # lock-step CRC path
        "empty": np.zeros((0, 3), np.float32),
        "flags": np.asarray([True, False]),
        "half": rng.normal(size=(3,)).astype(np.float16),
    }
    prefix = str(tmp_path / "ckpt" / "checkpoint-5")
    tb.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path / "ckpt")) == ["checkpoint-5.data-00000-of-00001", "checkpoint-5.index"]
    assert os.path.getsize(tb.data_path(prefix)) == sum(a.nbytes for a in tensors.values())
    r = tb.BundleReader(prefix)
    assert r.header == {"num_shards": 1, "endianness": 0, "producer": 1}
    assert r.keys() == sorted(tensors)                                  # byte order of the names
    for k, a in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == a.dtype and got.shape == a.shape and np.array_equal(got, a), k
    offs = [r.entries[k].offset for k in r.keys()]
    assert offs == sorted(offs) and offs[0] == 0                        # tensors laid out back to back in key order
    with pytest.raises(KeyError, match="not found in checkpoint"):
        r.get_tensor("vnet/missing")
    raw = bytearray(open(tb.data_path(prefix), "rb").read())            # flipped payload bit -> tensor CRC
    raw[r.entries["vnet/conv/weights"].offset + 5] ^= 1
    open(tb.data_path(prefix), "wb").write(raw)
    with pytest.raises(ValueError, match="tensor checksum"):
        tb.BundleReader(prefix).get_tensor("vnet/conv/weights")
    assert np.array_equal(tb.BundleReader(prefix).get_tensor("global_step"), tensors["global_step"])


def test_command_line_listing_and_conversion(tmp_path, capsys):
    prefix = str(tmp_path / "c")
    tb.write_bundle(prefix, {"a/w": np.ones((2, 3), np.float32), "global_step": np.asarray(3, np.int64)})
    assert tb._main(["list", prefix]) == 0
    out = capsys.readouterr().out
    assert "a/w" in out and "[2, 3]" in out and "int64" in out
    assert tb._main(["to-npz", prefix, str(tmp_path / "c.npz")]) == 0
    assert tb._main(["from-npz", str(tmp_path / "c.npz"), str(tmp_path / "d")]) == 0
    assert open(prefix + ".index", "rb").read() == open(str(tmp_path / "d.index"), "rb").read()
    assert tb._main([]) == 2


# ---- engine state <-> TF checkpoint ---------------------------------------------------------------------------
def _trained_engine(emul_lib, optimizer="Adam"):
    from oracle import ref_vnet as R
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=2, num_convolutions=(1, 2), bottom_convolutions=1)
    eng = engine_for(spec, 8, 1, "weighted_sorensen", (0.1, 1.0), emul_lib, optimizer=optimizer)
    eng.set_params(perturbed_params(spec, 3))
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 255, (1, 8, 8, 8, spec.in_channels)).astype(np.float32)
    y = (rng.uniform(size=(1, 8, 8, 8)) > 0.6).astype(np.int32)
    for _ in range(2):
        eng.train_step(x, y, dropout_rate=0.0)
    return spec, eng, (x, y)


def test_engine_state_roundtrips_through_a_tf_checkpoint(emul_lib, tmp_path):
    spec, eng, (x, y) = _trained_engine(emul_lib)
    prefix = checkpoint.save(eng, str(tmp_path), eng.global_step, start_epoch=3, format="tf")
    assert not os.path.exists(prefix + ".npz") and tb.is_bundle(prefix)
    assert checkpoint.latest(str(tmp_path)) == prefix
    r = tb.BundleReader(prefix)
    names = set(r.keys())
    # what tf.train.Saver() holds for the reference graph: variables, Adam slots, accumulators, step and epoch
    assert {"global_step", "start_epoch", "training/beta1_power", "training/beta2_power"} <= names
    for name, (shape, trainable) in eng.variables().items():
        assert r.entries[name].shape == shape and r.entries[name].dtype == 1
        assert ((name + "/Adam" in names) and (name + "/Adam_1" in names)) == trainable
    assert len(names) == len(eng.variables()) + 2 * sum(t for _, t in eng.variables().values()) + 4
    assert r.entries["global_step"].shape == () and r.entries["global_step"].dtype == 9
    assert r.entries["start_epoch"].shape == (1,) and r.entries["start_epoch"].dtype == 3
    assert r.get_tensor("training/beta1_power") == np.float32(0.9 ** 3)  # beta * beta^2 after two steps
    # a fresh engine restored from the bundle continues exactly like the original
    fresh = engine_for(spec, 8, 1, "weighted_sorensen", (0.1, 1.0), emul_lib)
    assert checkpoint.restore(fresh, prefix) == (2, 3)
    assert fresh.global_step == 2
    for name, (_, trainable) in eng.variables().items():
        assert np.array_equal(fresh.get_param(name), eng.get_param(name)), name
        if trainable:
            assert np.array_equal(fresh.get_param(name, _ffi.SLOT_ADAM_V), eng.get_param(name, _ffi.SLOT_ADAM_V)), name
    assert fresh.train_step(x, y, dropout_rate=0.0) == eng.train_step(x, y, dropout_rate=0.0)
    eng.close()
    fresh.close()


def test_reference_written_checkpoints_restore_by_name(emul_lib, tmp_path):
    """A checkpoint as the reference leaves it: Saver file set, `checkpoint-latest` state file with a second
    `all_model_checkpoint_paths` line, momentum slots, variables of another graph on the side."""
    spec, eng, _ = _trained_engine(emul_lib, optimizer="Momentum")
    state = checkpoint.tf_variables(eng, 2, 1)
    assert any(k.endswith("/Momentum") for k in state) and not any(k.endswith("/Adam") for k in state)
    assert "training/beta1_power" not in state
    state["unrelated/moving_mean"] = np.zeros(3, np.float32)
    tb.write_bundle(str(tmp_path / "checkpoint-2"), state)
    open(tmp_path / "checkpoint-2.meta", "wb").write(b"\x0a\x00")      # Saver's MetaGraphDef: never read here
    open(tmp_path / "checkpoint-latest", "w").write(
        'model_checkpoint_path: "checkpoint-2"\nall_model_checkpoint_paths: "checkpoint-2"\n')
    fresh = engine_for(spec, 8, 1, "weighted_sorensen", (0.1, 1.0), emul_lib, optimizer="Momentum")
    assert checkpoint.restore(fresh, checkpoint.latest(str(tmp_path))) == (2, 1)
    name = next(n for n, (_, t) in eng.variables().items() if t)
    assert np.array_equal(fresh.get_param(name, _ffi.SLOT_ADAM_M), eng.get_param(name, _ffi.SLOT_ADAM_M))
    # Adam engine, momentum checkpoint: variables load, slots stay zero (evaluation needs no more)
    adam = engine_for(spec, 8, 1, "weighted_sorensen", (0.1, 1.0), emul_lib)
    checkpoint.restore(adam, str(tmp_path / "checkpoint-2"))
    assert np.array_equal(adam.get_param(name), eng.get_param(name)) and not adam.get_param(name, _ffi.SLOT_ADAM_M).any()
    # a checkpoint of another architecture fails like Saver.restore: names the first missing key
    del state[name]
    tb.write_bundle(str(tmp_path / "other"), state)
    with pytest.raises(KeyError, match="not found in checkpoint"):
        checkpoint.restore(adam, str(tmp_path / "other"))
    state[name] = np.zeros((1, 2), np.float32)
    tb.write_bundle(str(tmp_path / "other"), state)
    with pytest.raises(ValueError, match="checkpoint shape"):
        checkpoint.restore(adam, str(tmp_path / "other"))
    for e in (eng, fresh, adam):
        e.close()


def test_checkpoint_format_setting_writes_both_kinds(emul_lib, tmp_path):
    from tests.test_host_mirror import _config
    from vnet_tensorflow_b200.model import image2label
    cfg = _config(tmp_path, Epoches=1, CheckpointFormat="both", Testing=False)
    m = image2label(None, cfg, library=emul_lib)
    m.train()
    step = m.engine.global_step
    prefix = str(tmp_path / "ckpt" / ("checkpoint-%d" % step))
    assert os.path.exists(prefix + ".npz") and tb.is_bundle(prefix)
    with np.load(prefix + ".npz") as z:
        r = tb.BundleReader(prefix)
        for k in z.files:
            assert np.array_equal(z[k], r.get_tensor(k)), k
    with pytest.raises(ValueError, match="CheckpointFormat"):
        checkpoint.save(m.engine, str(tmp_path / "ckpt"), step, 0, format="h5")
