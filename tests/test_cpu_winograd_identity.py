"""The algebra behind csrc/nf_wino.cu, checked on the CPU: a 3-tap vertical correlation producing two output rows from four
input rows equals the F(2,3) form with U = G g (what `to_winograd` folds on the host), T = B^T d (what the owner lane
publishes) and y = A^T (U * T) (the kernel's output transform, bias riding on m1)."""
import numpy as np


def test_vertical_f23_equals_direct_correlation():
    rng = np.random.RandomState(0)
    for _ in range(100):
        g = rng.randn(3)            # taps dy = 0, 1, 2 of one (dx, out, in)
        d = rng.randn(4)            # input rows r-1 .. r+2 of one column / channel
        b = rng.randn()
        direct = np.array([g @ d[0:3] + b, g @ d[1:4] + b])          # out[r] = sum_dy g[dy] * in[r + dy - 1]
        U = np.array([g[0], 0.5 * (g[0] + g[1] + g[2]), 0.5 * (g[0] - g[1] + g[2]), g[2]])
        T = np.array([d[0] - d[2], d[1] + d[2], d[2] - d[1], d[1] - d[3]])
        m = U * T
        sb = m[1] + b
        y = np.array([(m[0] + sb) + m[2], (sb - m[2]) - m[3]])
        assert np.allclose(y, direct, rtol=0, atol=1e-12)


def test_same_padding_rows_enter_as_zero_inputs():
    """Rows -1 and 32 of the convolution input are zero: the first tile of a pass has d0 = 0 (and d1 = 0 for conv-1's
    row pair (-1, 0)), the last one d3 = 0; the kernel forces h2 rows -1 and 32 to zero before transforming."""
    rng = np.random.RandomState(1)
    g = rng.randn(3)
    x = rng.randn(32)
    xp = np.concatenate([[0.0], x, [0.0, 0.0]])                      # rows -1 .. 33
    direct = np.array([g @ xp[r:r + 3] for r in range(32)])          # out[r] uses rows r-1, r, r+1
    U = np.array([g[0], 0.5 * g.sum(), 0.5 * (g[0] - g[1] + g[2]), g[2]])
    out = np.zeros(32)
    for k in range(16):                                              # tile k: output rows 2k, 2k+1 from rows 2k-1 .. 2k+2
        d = xp[2 * k:2 * k + 4]
        m = U * np.array([d[0] - d[2], d[1] + d[2], d[2] - d[1], d[1] - d[3]])
        out[2 * k], out[2 * k + 1] = m[0] + m[1] + m[2], m[1] - m[2] - m[3]
    assert np.allclose(out, direct, atol=1e-12)


def test_affine_tail_from_one_reciprocal():
    """The kernel's affine tail (csrc/nf_wino.cu `affine`): conv-3's tanh outputs arrive times 2 log2(e) (folded into the filters
    and bias by `to_winograd`), r = 1 / (exp2(h') + 1), tanh(h) = 1 - 2 r, exp(+-scale * tanh(h)) = exp2(+-(s2 - 2 s2 r)) with
    s2 = scale * log2(e), and the log-det of a pass is scale * (count - 2 sum r).  In fp32 this has to stay within the
    kernel's parity bounds (|dNLL| 1e-6 nats/dim, z 2e-5 relative) of the reference form, layers.py:342-372."""
    rng = np.random.RandomState(2)
    scale = np.float32(0.7)
    h = (rng.randn(64, 1024) * 3.0).astype(np.float32)               # pre-tanh outputs of one lane-column set, wide range
    z = rng.randn(64, 1024).astype(np.float32)
    shift = rng.randn(64, 1024).astype(np.float32)
    k23 = np.float32(2.8853900817779268)
    r = np.float32(1.0) / (np.exp2(h * k23, dtype=np.float32) + np.float32(1.0))
    s2 = np.float32(scale * np.float32(1.4426950408889634))
    x_kernel = z * np.exp2(np.float32(-2.0) * s2 * r + s2, dtype=np.float32) + shift          # inverse direction
    ls = scale.astype(np.float64) * np.tanh(h.astype(np.float64))
    x_ref = z.astype(np.float64) * np.exp(ls) + shift
    assert np.abs(x_kernel - x_ref).max() <= 2e-6 * np.abs(x_ref).max()
    z_kernel = (x_kernel - shift) * np.exp2(np.float32(2.0) * s2 * r - s2, dtype=np.float32)  # forward direction undoes it
    assert np.abs(z_kernel - z).max() <= 1e-5 * np.abs(z).max()      # (x - shift cancels: absolute, not relative, accuracy)
    ldj_kernel = scale * (np.float32(h.size) - np.float32(2.0) * r.sum(dtype=np.float32))
    assert abs(float(ldj_kernel) - ls.sum()) / h.size < 1e-6          # nats per dimension
    # the extremes saturate instead of overflowing: exp2(+inf) -> r = 0 -> tanh = 1; exp2(-inf) -> r = 1 -> tanh = -1
    with np.errstate(over="ignore"):
        big = np.float32(1.0) / (np.exp2(np.float32([400.0, -400.0]), dtype=np.float32) + np.float32(1.0))
    assert big.tolist() == [0.0, 1.0]
