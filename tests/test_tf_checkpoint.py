"""TensorFlow V2 checkpoint files under the reference's variable names (gaccum_b200/tf_checkpoint.py).
No TensorFlow in this image: the container format is checked against its published invariants and known-answer CRCs, and
by round trips; interchange with a TF-written file is NOT verified (said in the module header and DESIGN.md)."""
import os
import struct

import numpy as np
import pytest

from gaccum_b200 import tf_checkpoint as ck


def test_crc32c_known_answers_rfc3720_and_vectorised_path():
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                      # RFC 3720 B.4
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    rng = np.random.default_rng(5)
    for n in (65535, 65536, 65537, 300001):                         # around the switch to the lane-parallel path, ragged tails
        big = rng.integers(0, 256, n, dtype=np.uint8)
        assert ck.crc32c(big) == ck._raw_serial(0xFFFFFFFF, big.tobytes()) ^ 0xFFFFFFFF
    x = rng.standard_normal(50_000).astype(np.float32)
    assert ck.crc32c(x) == ck.crc32c(x.tobytes())
    # crc32c::Mask: rotate right by 15, add the constant
    c = ck.crc32c(b"123456789")
    assert ck.masked_crc32c(b"123456789") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _blocks(index_bytes):
    """Walk the table by its footer: [(offset, size)] of the data blocks, plus the index / metaindex handles."""
    footer = index_bytes[-48:-8]
    moff, p = ck._read_varint(footer, 0); msize, p = ck._read_varint(footer, p)
    ioff, p = ck._read_varint(footer, p); isize, p = ck._read_varint(footer, p)
    handles = []
    for _, h in ck._read_block(index_bytes, ioff, isize, True):
        o, q = ck._read_varint(h, 0); s, _ = ck._read_varint(h, q)
        handles.append((o, s))
    return handles, (moff, msize), (ioff, isize)


def test_bundle_container_invariants(tmp_path):
    prefix = str(tmp_path / "model.ckpt-12")
    tensors = {"bert/embeddings/word_embeddings": np.arange(24, dtype=np.float32).reshape(6, 4),
               "bert/embeddings/word_embeddings/adam_m": np.ones((6, 4), np.float32),
               "global_step": np.asarray(12, np.int64), "Variable": np.zeros((6, 4), np.float32)}
    ck.write_bundle(prefix, tensors)
    assert sorted(os.listdir(tmp_path)) == ["model.ckpt-12.data-00000-of-00001", "model.ckpt-12.index"]
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xDB4775248B80FB57          # table magic
    handles, meta, index = _blocks(idx)
    assert len(handles) == 1 and handles[0][0] == 0
    # blocks are laid out back to back, each followed by its 5-byte trailer; the footer closes the file
    assert meta[0] == handles[0][1] + 5 and index[0] == meta[0] + meta[1] + 5 and len(idx) == index[0] + index[1] + 5 + 48
    entries = ck._read_block(idx, *handles[0], True)
    keys = [k for k, _ in entries]
    assert keys == sorted(keys) and keys[0] == b""                          # header first, then bytewise key order ("Variable" < "bert/…" < "global_step")
    assert entries[0][1] == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])     # num_shards 1, version.producer 1
    # the data file holds the tensors in key order at the recorded offsets
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    off = 0
    for k, v in entries[1:]:
        e = ck._parse_entry(v)
        arr = tensors[k.decode()]
        assert e["offset"] == off and e["size"] == arr.nbytes and e["shape"] == list(arr.shape) and e["shard_id"] == 0
        assert e["dtype"] == (9 if arr.dtype == np.int64 else 1)           # DT_INT64 / DT_FLOAT
        assert data[off:off + arr.nbytes] == arr.tobytes() and e["crc32c"] == ck.masked_crc32c(arr)
        off += arr.nbytes
    assert off == len(data)
    # a flipped bit is detected, in the index and in the data
    bad = bytearray(idx); bad[3] ^= 1
    open(prefix + ".index", "wb").write(bad)
    with pytest.raises(ValueError, match="checksum"):
        ck.read_bundle(prefix)
    open(prefix + ".index", "wb").write(idx)
    bad = bytearray(data); bad[5] ^= 0x10
    open(prefix + ".data-00000-of-00001", "wb").write(bad)
    with pytest.raises(ValueError, match="checksum"):
        ck.read_bundle(prefix)
    assert ck.read_bundle(prefix, verify=False)                             # … unless told not to look
    with pytest.raises(ValueError, match="magic"):
        open(prefix + ".index", "wb").write(idx[:-1] + b"\x00")
        ck.read_bundle(prefix)


@pytest.mark.parametrize("block_size", [262144, 4096])
def test_bundle_round_trip_many_keys_spans_several_blocks(tmp_path, monkeypatch, block_size):
    """14 000 long keys: the index table needs more than one 256 KB data block (TensorFlow's block size), restart points
    every 16 keys, shared prefixes; with 4 KB blocks the index block itself gets long."""
    monkeypatch.setattr(ck, "_BLOCK_SIZE", block_size)
    rng = np.random.default_rng(11)
    tensors = {}
    for i in range(7000):
        name = f"bert/encoder/layer_{i % 24}/attention/self/some_rather_long_scope_name_to_fill_blocks/kernel_{i:05d}"
        tensors[name] = rng.standard_normal((i % 7, 3)).astype(np.float32)    # includes empty tensors
        tensors[name + "/adam_m"] = np.asarray(rng.standard_normal(), np.float32)   # scalars
    tensors["global_step"] = np.asarray(2 ** 40 + 3, np.int64)
    tensors["flags"] = np.array([True, False, True])
    tensors["ids"] = np.arange(5, dtype=np.int32)
    prefix = str(tmp_path / "m")
    ck.write_bundle(prefix, tensors)
    handles, _, _ = _blocks(open(prefix + ".index", "rb").read())
    assert len(handles) >= (2 if block_size == 262144 else 100)
    back = ck.read_bundle(prefix)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k


def test_reference_variable_names_both_variants():
    names = ["dense/kernel", "dense/bias"]
    rng = np.random.default_rng(1)
    state = {"global_step": np.asarray(9)}
    for n, shp in zip(names, [(3, 2), (2,)]):
        for suffix in ("", "/adam_m", "/adam_v", "/accum_grad"):
            state[n + suffix] = rng.standard_normal(shp).astype(np.float32)
    # variant A: AdamWeightDecayOptimizer slots (optimization.py:137-148); accumulators are unnamed tf.Variables (:78)
    ref = ck.to_reference_names(state, names, ck.ADAM_WEIGHT_DECAY)
    assert set(ref) == {"global_step", "dense/kernel", "dense/bias", "dense/kernel/adam_m", "dense/kernel/adam_v",
                        "dense/bias/adam_m", "dense/bias/adam_v", "Variable", "Variable_1"}
    assert ref["global_step"].dtype == np.int64 and np.array_equal(ref["Variable_1"], state["dense/bias/accum_grad"])
    back = ck.from_reference_names(ref, names, ck.ADAM_WEIGHT_DECAY)
    assert set(back) == set(state) and all(np.array_equal(back[k], state[k]) for k in state)
    # variant B: tf.train.AdamOptimizer slots + beta powers (02:41, another-example.py:135)
    state.update(beta1_power=np.float32(0.81), beta2_power=np.float32(0.998))
    ref = ck.to_reference_names(state, names, ck.ADAM)
    assert {"dense/kernel/Adam", "dense/kernel/Adam_1", "dense/bias/Adam", "dense/bias/Adam_1", "beta1_power", "beta2_power"} <= set(ref)
    assert not any(k.endswith("adam_m") for k in ref)
    back = ck.from_reference_names(ref, names, ck.ADAM)
    assert all(np.array_equal(back[k], state[k]) for k in state)
    # a parameters-only checkpoint (a pre-trained model) loads non-strictly and is refused strictly
    only_params = {n: state[n] for n in names}
    assert set(ck.from_reference_names(only_params, names, ck.ADAM_WEIGHT_DECAY, strict=False)) == {"global_step", *names}
    with pytest.raises(KeyError, match="adam_m"):
        ck.from_reference_names(only_params, names, ck.ADAM_WEIGHT_DECAY)


def test_checkpoint_state_file(tmp_path):
    assert ck.latest_checkpoint(str(tmp_path)) is None
    ck.write_checkpoint_state(str(tmp_path), "model.ckpt-300", ["model.ckpt-200", "model.ckpt-300"])
    text = open(tmp_path / "checkpoint").read()
    assert text.splitlines()[0] == 'model_checkpoint_path: "model.ckpt-300"' and text.count("all_model_checkpoint_paths") == 2
    assert ck.latest_checkpoint(str(tmp_path)) == str(tmp_path / "model.ckpt-300")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["adam_weight_decay", "adam"])
def test_interrupted_run_resumes_from_tf_checkpoint_bit_identically(tmp_path, variant):
    """Train into the 4th window, save in the TF format mid-window, restore into a FRESH train_op (fresh tensors),
    continue: parameters, moments and accumulators equal the uninterrupted run's bit for bit."""
    import torch
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    dev = torch.device("cuda", 0)
    names = ["bert/embeddings/word_embeddings", "bert/encoder/layer_0/output/dense/kernel", "bert/encoder/layer_0/output/dense/bias",
             "bert/encoder/layer_0/output/LayerNorm/gamma"]
    shapes = [(1000, 64), (64, 300), (300,), (64,)]
    hp = g.HParams.bert() if variant == "adam_weight_decay" else g.HParams.tf_adam()
    N, total, cut = 3, 12, 8            # windows {0},{1..3},{4..6},{7..9}: after 8 micro-steps g=7 has only accumulated
    rng = np.random.default_rng(3)
    init = [rng.standard_normal(s).astype(np.float32) * 0.02 for s in shapes]
    grads = [[torch.from_numpy(rng.standard_normal(s).astype(np.float32) * 1e-2).to(dev) for s in shapes] for _ in range(total)]
    lr_fn = lambda s: 1e-3 * (1 + s) / 16

    def fresh():
        return GaccumTrainOp([torch.from_numpy(x.copy()).to(dev) for x in init], names, hp, N, lr_fn)

    a = fresh()
    for s in range(total):
        a.run(grads[s])
    b = fresh()
    for s in range(cut):
        b.run(grads[s])
    prefix = ck.save(str(tmp_path), b)
    assert os.path.basename(prefix) == f"model.ckpt-{cut}" and ck.latest_checkpoint(str(tmp_path)) == prefix
    raw = ck.read_bundle(prefix)
    assert int(raw["global_step"]) == cut and raw["Variable_1"].shape == (64, 300)
    assert any(np.any(raw[k] != 0) for k in ("Variable", "Variable_1"))          # saved mid-window: accumulators are live
    slot = "/adam_m" if variant == "adam_weight_decay" else "/Adam"
    assert names[1] + slot in raw
    c = GaccumTrainOp([torch.zeros(s, device=dev) for s in shapes], names, hp, N, lr_fn)
    assert ck.restore(str(tmp_path), c) == prefix
    for s in range(cut, total):
        c.run(grads[s])
    torch.cuda.synchronize()
    assert c.global_step == a.global_step == total
    sa, sc = a.state_dict(), c.state_dict()
    assert set(sa) == set(sc)
    for k in sa:
        assert torch.equal(sa[k].cpu(), sc[k].cpu()), k


def test_cli_lists_and_converts_between_reference_and_shim_layouts(tmp_path):
    """python -m gaccum_b200.tf_checkpoint: reference-keyed file -> the shim's packed variables -> back, bit-identical."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = ["a/kernel", "a/LayerNorm/gamma", "a/bias"]
    rng = np.random.default_rng(2)
    ref = {"global_step": np.asarray(5, np.int64), "beta1_power": np.float32(0.729), "beta2_power": np.float32(0.997)}
    for i, (n, shp) in enumerate(zip(names, [(3, 5), (33,), (5,)])):
        for k in ("", "/Adam", "/Adam_1"):
            ref[n + k] = rng.standard_normal(shp).astype(np.float32)
        ref["Variable" if i == 0 else f"Variable_{i}"] = rng.standard_normal(shp).astype(np.float32)
    ck.write_bundle(str(tmp_path / "ref"), ref)
    (tmp_path / "names").write_text("\n".join(names) + "\n")
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    run = lambda *a: subprocess.run([sys.executable, "-m", "gaccum_b200.tf_checkpoint", *a], env=env, capture_output=True, text=True, timeout=120)
    r = run("convert", "to-shim", str(tmp_path / "ref"), str(tmp_path / "shim"), "--names", str(tmp_path / "names"), "--variant", "1")
    assert r.returncode == 0, r.stderr
    shim = ck.read_bundle(str(tmp_path / "shim"))
    assert shim["gaccum/adam_m"].shape == (32 + 64 + 32,)                    # 15 -> 32, 33 -> 64, 5 -> 32: gaccum_offsets' layout
    assert np.array_equal(shim["gaccum/adam_m"][32:65], ref["a/LayerNorm/gamma/Adam"]) and not shim["gaccum/adam_m"][65:96].any()
    assert np.array_equal(shim["gaccum/beta_powers"], np.float32([0.729, 0.997]))
    r = run("list", str(tmp_path / "shim"))
    assert r.returncode == 0 and "gaccum/accum_grads (float32) [128]" in r.stdout and "global_step (int64) []" in r.stdout
    r = run("convert", "to-reference", str(tmp_path / "shim"), str(tmp_path / "ref2"), "--names", str(tmp_path / "names"), "--variant", "1")
    assert r.returncode == 0, r.stderr
    back = ck.read_bundle(str(tmp_path / "ref2"))
    assert set(back) == set(ref) and all(np.array_equal(back[k], ref[k]) for k in ref)


def test_slab_offsets_agree_with_the_c_abi_layout():
    import gaccum_b200 as g
    from gaccum_b200 import _lib
    sizes = [100, 7, 2048, 1, 33, 0, 4097]
    plan = _lib.Plan(sizes, None, g.HParams.bert(), device=-1)
    offs, total = ck._slab_offsets(sizes)
    assert offs == list(plan.offsets) and total == plan.padded_size
