"""The TF-1.12 / TFP-0.5 API stand-in (`oracle/tf1_shim.py`) against TensorFlow's own documented behaviour.

The reference-generated goldens (tests/golden/ref_*.npz) are the reference's Python executed over this stand-in, so the
stand-in is the one restated component left on the reference side.  These tests pin its primitives to the examples and
rules in TensorFlow's / TensorFlow-Probability's documentation (quoted in the docstrings below) and to first principles
(direct loops), independently of both the oracle and the CUDA path."""
import numpy as np
import pytest

from oracle import tf1_shim as tf


@pytest.fixture(autouse=True)
def fresh_graph():
    tf.reset_default_graph()
    yield
    tf.reset_default_graph()


def _np(t):
    return t.numpy()


def test_fill_triangular_docstring_examples():
    """tfp.distributions.fill_triangular docstring:
    fill_triangular([1, 2, 3, 4, 5, 6])            == [[4, 0, 0], [6, 5, 0], [3, 2, 1]]
    fill_triangular([1, 2, 3, 4, 5, 6], upper=True) == [[1, 2, 3], [0, 5, 6], [0, 0, 4]]"""
    v = tf.constant([1., 2, 3, 4, 5, 6])
    assert np.array_equal(_np(tf.fill_triangular(v)), [[4, 0, 0], [6, 5, 0], [3, 2, 1]])
    assert np.array_equal(_np(tf.fill_triangular(v, upper=True)), [[1, 2, 3], [0, 5, 6], [0, 0, 4]])
    for upper in (False, True):       # fill_triangular_inverse docstring: the inverse packing
        m = tf.fill_triangular(v, upper=upper)
        assert np.array_equal(_np(tf.fill_triangular_inverse(m, upper=upper)), [1, 2, 3, 4, 5, 6])
    b = tf.constant(np.arange(20.).reshape(2, 10))      # batched, n = 4
    for upper in (False, True):
        assert np.array_equal(_np(tf.fill_triangular_inverse(tf.fill_triangular(b, upper=upper), upper=upper)), _np(b))


def test_matrix_band_part_docstring_examples():
    """tf.matrix_band_part docstring:
    input = [[0, 1, 2, 3], [-1, 0, 1, 2], [-2, -1, 0, 1], [-3, -2, -1, 0]]
    band_part(input, 1, -1) = [[0, 1, 2, 3], [-1, 0, 1, 2], [0, -1, 0, 1], [0, 0, -1, 0]]
    band_part(input, 2, 1)  = [[0, 1, 0, 0], [-1, 0, 1, 0], [-2, -1, 0, 1], [0, -2, -1, 0]]"""
    x = tf.constant([[0., 1, 2, 3], [-1, 0, 1, 2], [-2, -1, 0, 1], [-3, -2, -1, 0]])
    assert np.array_equal(_np(tf.matrix_band_part(x, 1, -1)), [[0, 1, 2, 3], [-1, 0, 1, 2], [0, -1, 0, 1], [0, 0, -1, 0]])
    assert np.array_equal(_np(tf.matrix_band_part(x, 2, 1)), [[0, 1, 0, 0], [-1, 0, 1, 0], [-2, -1, 0, 1], [0, -2, -1, 0]])
    assert np.array_equal(_np(tf.matrix_set_diag(x, tf.constant([9., 8, 7, 6]))).diagonal(), [9, 8, 7, 6])


def test_conv2d_is_nhwc_cross_correlation_with_hwio_filter():
    """tf.nn.conv2d: output[b, i, j, k] = sum_{di, dj, q} input[b, i + di, j + dj, q] * filter[di, dj, q, k] (no kernel
    flip); 'SAME' with stride 1 pads (k - 1) / 2 zeros on each side, 'VALID' pads nothing."""
    rng = np.random.RandomState(0)
    x, w = rng.randn(2, 5, 6, 3), rng.randn(3, 3, 3, 4)
    for padding in ("SAME", "VALID"):
        got = _np(tf.conv2d(tf.constant(x), tf.constant(w), [1, 1, 1, 1], padding))
        xp = np.pad(x, [(0, 0), (1, 1), (1, 1), (0, 0)]) if padding == "SAME" else x
        H, W = xp.shape[1] - 2, xp.shape[2] - 2
        want = np.zeros((2, H, W, 4))
        for i in range(H):
            for j in range(W):
                want[:, i, j, :] = np.einsum("bdeq,deqk->bk", xp[:, i:i + 3, j:j + 3, :], w)
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-12
    one = _np(tf.conv2d(tf.constant(x), tf.constant(w[:1, :1]), [1, 1, 1, 1], "SAME"))
    assert np.abs(one - x @ w[0, 0]).max() < 1e-12


def test_conv2d_against_scipy_correlate2d():
    """A third, independent implementation: scipy.signal.correlate2d per (batch, in, out) plane -- mode 'same' with zero fill is
    TF's SAME at stride 1 and an odd kernel, mode 'valid' is VALID (the reference's convs are 3x3 / 1x1 SAME, layers.py:452-498)."""
    from scipy.signal import correlate2d
    rng = np.random.RandomState(3)
    x, w = rng.randn(2, 7, 5, 2), rng.randn(3, 3, 2, 3)
    for padding, mode in (("SAME", "same"), ("VALID", "valid")):
        got = _np(tf.conv2d(tf.constant(x), tf.constant(w), [1, 1, 1, 1], padding))
        for b in range(2):
            for k in range(3):
                want = sum(correlate2d(x[b, :, :, q], w[:, :, q, k], mode=mode, boundary="fill", fillvalue=0.0) for q in range(2))
                assert np.abs(got[b, :, :, k] - want).max() < 1e-12


def test_moments_where_one_hot_pad_tile_split_gather():
    """tf.nn.moments: mean and POPULATION variance over the axes; tf.where(cond): int64 coordinates [k, rank] of the true
    elements; tf.one_hot: indices outside [0, depth) give all-zero rows; tf.pad / tf.tile / tf.split / tf.gather as
    documented."""
    rng = np.random.RandomState(1)
    x = rng.randn(4, 3, 3, 2)
    m, v = tf.moments(tf.constant(x), [0, 1, 2])
    assert np.allclose(_np(m), x.mean(axis=(0, 1, 2))) and np.allclose(_np(v), x.var(axis=(0, 1, 2)))       # ddof = 0
    iso_vals = tf.constant([100., 400, 800, 1600, 3200])
    idx = tf.where(tf.equal(iso_vals, tf.constant([800.])))
    assert _np(idx).tolist() == [[2]] and _np(idx).dtype == np.int64
    assert _np(tf.where(tf.equal(iso_vals, tf.constant([250.])))).shape == (0, 1)                          # unknown ISO: empty
    assert _np(tf.one_hot(idx, 5)).tolist() == [[[0, 0, 1, 0, 0]]]
    assert _np(tf.one_hot(tf.constant([1, -1, 7]), 3)).tolist() == [[0, 1, 0], [0, 0, 0], [0, 0, 0]]
    assert float(_np(tf.reduce_sum(tf.one_hot(tf.where(tf.equal(iso_vals, tf.constant([250.]))), 5) * iso_vals))) == 0.0
    t = tf.constant([[1., 2, 3], [4, 5, 6]])
    assert _np(tf.pad(t, [[1, 1], [2, 2]])).tolist() == [[0] * 7, [0, 0, 1, 2, 3, 0, 0], [0, 0, 4, 5, 6, 0, 0], [0] * 7]   # tf.pad docstring
    assert _np(tf.tile(t, [1, 2])).tolist() == [[1, 2, 3, 1, 2, 3], [4, 5, 6, 4, 5, 6]]
    a, b = tf.split(tf.constant(np.arange(8.).reshape(1, 1, 2, 4)), 2, axis=-1)
    assert _np(a).ravel().tolist() == [0, 1, 4, 5] and _np(b).ravel().tolist() == [2, 3, 6, 7]
    assert _np(tf.gather(t, [2, 0], axis=-1)).tolist() == [[3, 1], [6, 4]]
    p = tf.Permute(permutation=[3, 2, 1, 0])
    z = tf.constant(np.arange(8.).reshape(1, 1, 2, 4))
    assert np.array_equal(_np(p._inverse(p._forward(z))), _np(z)) and _np(p._forward(z))[0, 0, 0].tolist() == [3, 2, 1, 0]
    assert not hasattr(p, "_inverse_and_log_det_jacobian")        # TFP 0.5 Permute has no fused method (reference falls back)


def test_variable_scope_get_variable_and_template_naming_rules():
    """variable_scope.py / template.py (TF 1.12): names are `<scope>/<name>`; creating an existing variable without reuse
    raises, reusing a missing one raises, AUTO_REUSE does either; reuse is inherited by nested scopes and AUTO_REUSE
    overrides an inherited True; `variable_scope(None, default_name=d)` gives d, d_1, d_2 ... within the current scope;
    make_template opens its scope at the FIRST CALL (in the scope current then) and re-enters it with reuse afterwards."""
    init = tf.constant_initializer(1.5)
    with tf.variable_scope("model"):
        v = tf.get_variable("w", [2], initializer=init)
        assert v.var_name == "model/w" and _np(v).tolist() == [1.5, 1.5]
        with pytest.raises(ValueError):
            tf.get_variable("w", [2], initializer=init)
        with tf.variable_scope("sdn_gain", reuse=tf.AUTO_REUSE):
            g1 = tf.get_variable("gain_val", [1], initializer=init)
            g2 = tf.get_variable("gain_val", [1], initializer=init)
            assert g1 is g2 and g1.var_name == "model/sdn_gain/gain_val"
    with tf.variable_scope("model", reuse=True):
        assert tf.get_variable("w", [2], initializer=init) is v
        with pytest.raises(ValueError):
            tf.get_variable("missing", [1], initializer=init)
        with tf.variable_scope("inner"):                       # reuse=True is inherited
            with pytest.raises(ValueError):
                tf.get_variable("x", [1], initializer=init)
            with tf.variable_scope("auto", reuse=tf.AUTO_REUSE):   # ... unless AUTO_REUSE overrides it
                assert tf.get_variable("x", [1], initializer=init).var_name == "model/inner/auto/x"

    def fn(x):
        return x * tf.get_variable("k", [1], initializer=init)
    t_a, t_b, t_c = (tf.make_template("real_nvp_conv_template", fn) for _ in range(3))
    assert t_a.variable_scope is None                          # nothing is named at construction
    one = tf.constant([1.0])
    with tf.variable_scope("model", reuse=True):               # like NoiseFlow.sample: first calls under reuse=True ...
        with tf.variable_scope("l", reuse=tf.AUTO_REUSE):      # ... inside the layers' AUTO_REUSE scopes
            pass
        with pytest.raises(ValueError):                        # the template's own get_variable is NOT under AUTO_REUSE here
            t_c(one)
    tf.reset_default_graph()
    t_a, t_b, t_c = (tf.make_template("real_nvp_conv_template", fn) for _ in range(3))
    with tf.variable_scope("model"):
        t_b(one); t_a(one)                                     # call order, not construction order, hands out the names
        assert t_b.variable_scope.name == "model/real_nvp_conv_template"
        assert t_a.variable_scope.name == "model/real_nvp_conv_template_1"
    with tf.variable_scope("other"):
        t_c(one)
        t_b(one)                                               # later calls re-enter the captured scope, wherever they happen
    assert t_c.variable_scope.name == "other/real_nvp_conv_template"
    assert sorted(tf.get_default_graph().vars) == ["model/real_nvp_conv_template/k", "model/real_nvp_conv_template_1/k",
                                                   "other/real_nvp_conv_template/k"]


def test_graph_mode_placeholders_cond_assign_and_gradients():
    """Deferred execution: a tensor built from a variable follows later assignments; tf.cond evaluates only the taken
    branch in a run; assign_sub takes effect once per Session.run that needs it (and not at graph construction);
    tf.gradients is d ys / d xs; AdamOptimizer's first step moves a variable by lr * g / (|g| + eps) ~ lr (adam.py docstring:
    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; var -= lr_t m / (sqrt(v) + eps))."""
    x = tf.placeholder(tf.float32, [None, 2], name="x")
    flag = tf.placeholder(tf.bool_, name="is_training")
    w = tf.get_variable("w", [2], initializer=tf.constant_initializer([1.0, 2.0]))
    m = tf.get_variable("m", [2], initializer=tf.zeros_initializer, trainable=False)
    derived = w * 3.0                                           # built once, like Conv2d1x1's A in its constructor
    batch_mean = tf.reduce_mean(x, axis=[0])
    upd = tf.assign_sub(m, 0.1 * (m - batch_mean))

    def train_branch():
        with tf.control_dependencies([upd]):
            return x - batch_mean

    y = tf.cond(tf.equal(flag, tf.constant(True)), train_branch, lambda: x - m)
    loss = tf.reduce_sum(y * derived)
    assert _np(m).tolist() == [0.0, 0.0]                        # graph construction moved nothing
    sess = tf.Session()
    feed = {x: np.array([[1.0, 2.0], [3.0, 6.0]]), flag: False}
    assert abs(float(sess.run(loss, feed)) - (3 * (1 + 3) + 6 * (2 + 6))) < 1e-12
    assert _np(m).tolist() == [0.0, 0.0]                        # eval branch: no update
    w.load([2.0, 2.0])
    assert abs(float(sess.run(loss, feed)) - (6 * 4 + 6 * 8)) < 1e-12      # `derived` follows the assignment
    feed[flag] = True
    out, _ = sess.run([y, loss], feed)
    assert np.allclose(out, [[-1, -2], [1, 2]]) and np.allclose(_np(m), [0.2, 0.4])     # one update per run, not per fetch
    gw, = tf.gradients(loss, [w])
    feed[flag] = False
    assert np.allclose(sess.run(gw, feed), 3 * (np.array([4.0, 8.0]) - 2 * _np(m)))
    opt = tf.AdamOptimizer(learning_rate=1e-3)
    step = opt.minimize(loss)
    before = _np(w).copy()
    sess.run(step, feed)
    assert np.allclose(before - _np(w), 1e-3, rtol=1e-6)        # first Adam step = lr * sign(g) for |g| >> eps
