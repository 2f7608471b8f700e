"""CPU models of the integer / float tricks the streaming kernels rely on (csrc/pix4.cuh, csrc/metric.cu), checked
exhaustively or on adversarial data in numpy.  They pin the arithmetic itself; the GPU tier checks the kernels."""
import numpy as np
import pytest

U32 = np.uint32


def _bytes(w):
    return np.stack([(w >> U32(8 * i)) & U32(255) for i in range(4)], axis=-1)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5])
def test_invalid_byte_detector_is_exact(K):
    """tc_invalid / confusion_bytes16: bit 7 of byte i of (((w & 0x7f7f7f7f) + ADD) | w) is set iff byte i >= K."""
    rng = np.random.default_rng(K)
    vals = np.arange(256, dtype=np.uint32)
    add = U32((0x80 - K) * 0x01010101)
    for pos in range(4):
        others = rng.integers(0, 256, size=(64, 256, 4), dtype=np.uint32)
        others[..., pos] = vals[None, :]
        w = (others[..., 0] | (others[..., 1] << U32(8)) | (others[..., 2] << U32(16)) | (others[..., 3] << U32(24))).astype(np.uint32)
        flag = (((w & U32(0x7F7F7F7F)) + add) | w) & U32(0x80808080)
        got = _bytes(flag) != 0
        assert np.array_equal(got, _bytes(w) >= K)


def test_zero_byte_detector_is_exact():
    """ignore mask of confusion_bytes16: bit 7 of byte i of (((x & 0x7f..) + 0x7f..) | x) is set iff byte i != 0."""
    rng = np.random.default_rng(0)
    x = rng.integers(0, 2 ** 32, size=200000, dtype=np.uint64).astype(np.uint32)
    x[:4096] &= rng.choice(np.array([0xFFFFFF00, 0xFFFF00FF, 0xFF00FFFF, 0x00FFFFFF, 0, 0xFF00FF00], dtype=np.uint32), 4096)
    nz = ((x & U32(0x7F7F7F7F)) + U32(0x7F7F7F7F)) | x
    assert np.array_equal(_bytes(nz & U32(0x80808080)) != 0, _bytes(x) != 0)


@pytest.mark.parametrize("K,FW", [(2, 8), (3, 8), (4, 8), (5, 6)])
def test_field_shift_from_one_multiply(K, FW):
    """tc_fields: for label bytes < K, (word * FW) carries FW * label in every byte and its low five bits survive the
    wrap-mode shift, so 1 << ((m >> 8 i) & 31) is the counter field of label i; fields of 16 labels cannot overflow."""
    labels = np.stack(np.meshgrid(*[np.arange(K, dtype=np.uint32)] * 4, indexing="ij"), -1).reshape(-1, 4)
    w = labels[:, 0] | (labels[:, 1] << U32(8)) | (labels[:, 2] << U32(16)) | (labels[:, 3] << U32(24))
    m = (w * U32(FW)).astype(np.uint32)
    for i in range(4):
        shift = (m >> U32(8 * i)) & U32(31)
        assert np.array_equal(shift, labels[:, i] * FW)
        assert np.all(shift + FW <= 32)
    assert 16 < (1 << FW)                       # one frame of 16 labels fits a field before the spill


def _argmax_float_domain(x):
    """pix4.cuh argmax2f: m = max, n_c = (x_c != m), idx = n_0 (1 + n_1 (1 + ...)) with t <- fma(n_c, t, n_c)."""
    x = x.astype(np.float32)
    m = x.max(axis=0)
    n = (x != m[None]).astype(np.float32)
    t = n[x.shape[0] - 2]
    for c in range(x.shape[0] - 3, -1, -1):
        t = n[c] * t + n[c]
    return t


@pytest.mark.parametrize("C", [2, 3, 4, 5, 8])
def test_float_domain_argmax_matches_first_maximum(C):
    rng = np.random.default_rng(C)
    x = rng.standard_normal((C, 50000)).astype(np.float32).round(1)          # many ties
    x[:, :64] = 0.0
    x[::2, :64] = -0.0                                                        # -0 == +0: index 0
    x[:, 64:128] = np.float32(1e-45) * rng.integers(-1, 2, size=(C, 64))      # denormal ties
    x[C - 1, 128:256] = 1e30                                                  # last class
    x[:, 256:300] = np.float32(np.inf) * rng.choice([-1.0, 1.0], size=(C, 44)).astype(np.float32)
    idx = _argmax_float_domain(x)
    assert np.array_equal(idx, np.argmax(x, axis=0).astype(np.float32))       # np.argmax: first maximum, like torch.max


def test_label_bytes_from_biased_float_indices():
    """PixIO<2>::label_word: {i0 + 65536 i2, i1 + 65536 i3} + 2^23 are exact in fp32 for class indices < 64, and the
    byte permute 0x6240 of the two mantissas is the little-endian label word."""
    rng = np.random.default_rng(1)
    i = rng.integers(0, 64, size=(4, 10000)).astype(np.float32)
    lo = (i[2] * np.float32(65536.0) + (i[0] * np.float32(1.0) + np.float32(8388608.0))).astype(np.float32)
    hi = (i[3] * np.float32(65536.0) + (i[1] * np.float32(1.0) + np.float32(8388608.0))).astype(np.float32)
    a, b = lo.view(np.uint32), hi.view(np.uint32)
    ab = [(a >> U32(8 * k)) & U32(255) for k in range(4)] + [(b >> U32(8 * k)) & U32(255) for k in range(4)]
    word = ab[0] | (ab[4] << U32(8)) | (ab[2] << U32(16)) | (ab[6] << U32(24))  # selector nibbles 0, 4, 2, 6
    exp = i.astype(np.uint32)
    assert np.array_equal(word, exp[0] | (exp[1] << U32(8)) | (exp[2] << U32(16)) | (exp[3] << U32(24)))
    fw = np.float32(6.0)
    m6 = (i[0] * fw + np.float32(8388608.0)).astype(np.float32).view(np.uint32)
    ok = i[0] * 6 < 32
    assert np.array_equal((m6 & U32(31))[ok], (exp[0] * 6)[ok])


def test_interleaved_confusion_mapping_covers_every_label_once():
    """fuvs_confusion with an int64 operand: lane l of a warp owns the label pairs 32 k + l (k = 0..7) of a 512-label
    block; together the lanes cover the block exactly once and each load instruction k is one contiguous run."""
    owned = np.zeros(512, dtype=np.int32)
    for lane in range(32):
        for k in range(8):
            pair = 32 * k + lane
            owned[2 * pair:2 * pair + 2] += 1
    assert np.all(owned == 1)
    for k in range(8):
        pairs = np.array([32 * k + lane for lane in range(32)])
        assert np.array_equal(pairs, np.arange(32 * k, 32 * k + 32))
