"""CPU tests of the TensorFlow-bundle reader/writer (rlcontrol_b200/tf_bundle.py).
Pinned on the reference's own checkpoints: every block and tensor crc32c of the five
``Bimodal1DEnv_trueQ_ckpt`` bundles verifies, the decoded critics equal the golden weights, and re-writing a
decoded checkpoint reproduces BOTH files byte for byte (those tests need /root/reference and skip without it)."""
import glob
import os

import numpy as np
import pytest

from conftest import golden
from rlcontrol_b200 import tf_bundle as tb

REF_DIR = "/root/reference/Bimodal1DEnv_trueQ_ckpt"
REF = sorted(glob.glob(os.path.join(REF_DIR, "*.index")))
need_ref = pytest.mark.skipif(not REF, reason="reference checkpoints not present (GPU box)")


def test_crc32c_known_answers():
    assert tb.crc32c(b"") == 0 and tb.crc32c(b"123456789") == 0xE3069283        # RFC 3720 check value
    assert tb.crc32c(bytes(32)) == 0x8A9136AA and tb.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tb.mask_crc(0) == 0xA282EAD8


def test_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    t = {"main/qf/fully_connected/weights": rng.randn(3, 7).astype(np.float32), "main/qf/fully_connected/biases": rng.randn(7).astype(np.float32),
         "step": np.array(12345, np.int64), "scalar": np.float32(0.25), "d": rng.randn(2, 2, 2), "i": np.arange(5, dtype=np.int32)}
    for i in range(40):                                           # more than one restart interval of keys
        t["z/var_%02d" % i] = rng.randn(i % 4 + 1).astype(np.float32)
    pre = str(tmp_path / "ck" / "model")
    tb.write_bundle(pre, t)
    back = tb.read_bundle(pre)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == np.asarray(t[k]).dtype and back[k].shape == np.asarray(t[k]).shape
        np.testing.assert_array_equal(back[k], t[k])
    raw = bytearray(open(pre + ".data-00000-of-00001", "rb").read())
    raw[5] ^= 1
    open(pre + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tb.read_bundle(pre)                                       # tensor crc32c catches the flipped bit
    assert tb.read_bundle(pre, verify=False)
    with pytest.raises(ValueError):
        open(pre + ".index", "wb").write(b"not a table" * 10)
        tb.read_index(pre)


@need_ref
def test_reads_reference_checkpoints_and_matches_golden():
    g = golden("trueq.npz")
    for f in REF:
        pre = f[:-len(".index")]
        name = os.path.basename(pre).replace("Bimodal1DEnv_", "").replace("_trueQ_learned", "")
        ent = tb.read_index(pre)                                  # verifies the table block checksums
        assert len(ent) == 20 and ent["main/qf/fully_connected_1/weights"]["shape"] == [201, 200]
        W1, b1, W2, b2, W3, b3 = tb.read_critic(pre)               # verifies the per-tensor crc32c
        for k, v in zip(("W1", "b1", "W2", "b2", "W3", "b3"), (W1, b1, W2, b2, W3, b3)):
            np.testing.assert_array_equal(v, g["%s_%s" % (name, k)])


@need_ref
def test_rewriting_a_reference_checkpoint_is_byte_identical(tmp_path):
    for f in REF:
        pre = f[:-len(".index")]
        out = str(tmp_path / os.path.basename(pre))
        tb.write_bundle(out, tb.read_bundle(pre))
        for ext in (".index", ".data-00000-of-00001"):
            assert open(pre + ext, "rb").read() == open(out + ext, "rb").read(), (pre, ext)
