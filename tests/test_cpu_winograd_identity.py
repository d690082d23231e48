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
