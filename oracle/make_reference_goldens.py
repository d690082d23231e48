"""TEST INFRASTRUCTURE -- generates `tests/golden/ref_*.npz` by executing the REFERENCE'S OWN Python
(`/root/reference/borealisflows/*.py`, imported unmodified from where it lies) over the TF-1.12 API stand-in of
`oracle/tf1_shim.py` in double precision.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_reference_goldens.py            # rewrites tests/golden/ref_*.npz

What runs from the reference: `NoiseFlow.__init__/noise_flow_arch/inverse/forward/sample/_loss/loss/prior`
(noise_flow_model.py), `AffineCoupling`, `Conv2d1x1`, `real_nvp_conv_template`, `conv2d`, `conv2d_zeros`,
`add_edge_padding`, `batch_norm` (layers.py), `matrix_param_lu` + `_vec2stricttri/_stricttri2vec` (matrix_param.py),
`squeeze2d/unsqueeze2d` (utils.py), every `AffineCouplingSdn*/Gain*` class + `cond_utils.py`, and
`NoiseFlowWrapper.hps_loader` (NoiseFlowWrapper.py:89-138).  What is restated: the TF primitives (see tf1_shim.py).

Flows (the training scripts themselves need the SIDD data set and cannot run; their graph construction and
`sess.run` calls are repeated here with the same placeholders):
  * training graph, train_noise_flow.py:284-302: placeholders, `NoiseFlow(x_shape[1:], is_training, hps)`,
    `nf.loss(x, y, nlf0, nlf1, iso, cam)` FIRST (this names the template scopes data->latent),
    `AdamOptimizer(lr, 0.9, 0.999, 1e-8).minimize(loss)` (:187-198), `nf.sample(...)` (sidd_utils.py:1165-1170), then
    `Saver.restore` and the `sess.run` calls of test_multithread (:108-112, is_training False), sample_multithread
    (:165-167, True) and train_multithread (:62-71, `[train_op, loss, sd_z]`, True);
  * the reference's `NoiseFlowWrapper` class ITSELF, unmodified: constructor (NoiseFlowWrapper.py:20-79: placeholders,
    model, `nf.sample` ONLY -- template scopes get named latent->data --, Saver.restore) and `sample_noise_nf` (:81-87);
  * small cases for every token `noise_flow_arch` parses, with perturbed variables; `squeeze2d/unsqueeze2d`.
`tf.random_normal` draws are injected (the TF RNG stream is not reproducible anywhere else).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import tf1_shim  # noqa: E402

tf = tf1_shim.install()
sys.path.insert(0, REF)
from borealisflows.noise_flow_model import NoiseFlow  # noqa: E402  (the reference's class)
from borealisflows.NoiseFlowWrapper import NoiseFlowWrapper as RefWrapper  # noqa: E402
from borealisflows.utils import squeeze2d as ref_squeeze2d, unsqueeze2d as ref_unsqueeze2d  # noqa: E402

from noise_flow_b200.tf_checkpoint import load_checkpoint  # noqa: E402  (TF-free reader of the shipped bundle)

sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import CAM_ISO_NLF, synth_batch  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MODEL_DIR = os.path.join(REF, "models", "NoiseFlow")


def ref_hps():
    """hps.txt through the reference's own parser (NoiseFlowWrapper.py:89-138; `self` is unused there)."""
    hps = RefWrapper.hps_loader(None, os.path.join(MODEL_DIR, "hps.txt"))
    hps.x_shape = [None, 32, 32, 4]                     # NoiseFlowWrapper.py:47-49
    return hps


def c(a):
    return tf.constant(np.asarray(a, dtype=np.float64))


def var_snapshot():
    return {n: v.t.detach().numpy().copy() for n, v in tf.get_default_graph().vars.items()}


def inject_eps(eps):
    tf.get_default_graph().random_normal_hook = lambda shape: eps.reshape(shape)


def placeholders():
    """train_noise_flow.py:284-291 / NoiseFlowWrapper.py:50-56"""
    x_shape = [None, 32, 32, 4]
    return dict(is_training=tf.placeholder(tf.bool, name='is_training'),
                x=tf.placeholder(tf.float32, x_shape, name='noise_image'),
                y=tf.placeholder(tf.float32, x_shape, name='clean_image'),
                nlf0=tf.placeholder(tf.float32, [None], name='nlf0'), nlf1=tf.placeholder(tf.float32, [None], name='nlf1'),
                iso=tf.placeholder(tf.float32, [None], name='iso'), cam=tf.placeholder(tf.float32, [None], name='cam'),
                lr=tf.placeholder(tf.float32, None, name='learning_rate'))


def training_graph_goldens():
    tf.reset_default_graph()
    np.random.seed(0)
    hps = ref_hps()
    n = 4
    x, y = synth_batch(n, cam=2, iso=100, seed=11)       # fp32-representable inputs
    eps = np.random.RandomState(12).randn(n, 32, 32, 4).astype(np.float32)
    out = {"x": x, "y": y, "eps": eps, "iso": np.float32(100), "cam": np.float32(2)}

    # ---- graph construction, in the order of train_noise_flow.py:284-302,187-198 and sidd_utils.py:1165-1170
    p = placeholders()
    nf = NoiseFlow(hps.x_shape[1:], p["is_training"], hps)
    loss_val, sd_z = nf.loss(p["x"], p["y"], nlf0=p["nlf0"], nlf1=p["nlf1"], iso=p["iso"], cam=p["cam"])
    opt = tf.train.AdamOptimizer(learning_rate=p["lr"], beta1=0.9, beta2=0.999, epsilon=1e-08)
    train_op = opt.minimize(loss_val)
    nll_vec, _ = nf._loss(p["x"], p["y"], nlf0=p["nlf0"], nlf1=p["nlf1"], iso=p["iso"], cam=p["cam"], reuse=True)
    x_sample_T1 = nf.sample(p["y"], 1.0, p["y"], p["nlf0"], p["nlf1"], p["iso"], p["cam"])
    x_sample_T06 = nf.sample(p["y"], 0.6, p["y"], p["nlf0"], p["nlf1"], p["iso"], p["cam"])
    with tf.variable_scope("model", reuse=True):
        z_t, obj_t = nf.inverse(p["x"], tf.zeros_like(p["x"], dtype='float32')[:, 0, 0, 0], yy=p["y"], nlf0=p["nlf0"],
                                nlf1=p["nlf1"], iso=p["iso"], cam=p["cam"])
        xr_t = nf.forward(z_t, None, yy=p["y"], nlf0=p["nlf0"], nlf1=p["nlf1"], iso=p["iso"], cam=p["cam"])
    tv = tf.trainable_variables()
    grads_t = tf.gradients(loss_val, tv)

    g = tf.get_default_graph()
    out["layer_names"] = np.array(nf.get_layer_names())
    out["var_names"] = np.array(list(g.vars))
    out["var_trainable"] = np.array([v.trainable for v in g.vars.values()])
    out["var_shapes"] = np.array([",".join(map(str, v.t.shape)) for v in g.vars.values()])
    out["num_params"] = np.int64(np.sum([np.prod(v.get_shape().as_list()) for v in tv]))     # train_noise_flow.py:309-310
    out["init_std_l_1_W"] = np.float64(np.std(np.concatenate([v.numpy().ravel() for v in tv if v.var_name.endswith("/l_1/W")])))

    sess = tf.Session()
    sess.run(tf.global_variables_initializer())
    saver = tf.train.Saver()
    saver.restore(sess, os.path.join(MODEL_DIR, "ckpt", "model.ckpt.best"))
    out["ckpt_unused"] = np.array(saver.last_unused)

    def fd(training):
        return {p["x"]: x, p["y"]: y, p["nlf0"]: [0.000479], p["nlf1"]: [0.000002], p["iso"]: [100.0], p["cam"]: [2.0],
                p["is_training"]: training, p["lr"]: 1e-4}

    # ---- moving statistics (is_training: False), as test_multithread does (train_noise_flow.py:108-112)
    inject_eps(eps)
    r = sess.run([nll_vec, sd_z, loss_val, z_t, obj_t, xr_t, x_sample_T1, x_sample_T06] + grads_t, feed_dict=fd(False))
    out["nll"], out["sd_z"], out["loss"], out["z"], out["logdet"] = r[0], r[1], r[2], r[3], r[4]
    out["roundtrip_err"] = np.float64(np.abs(r[5] - x).max())
    out["sample_T1"], out["sample_T0.6"] = r[6], r[7]
    for v, gr in zip(tv, r[8:]):    # tf.gradients gives None for variables the loss does not depend on: stored as zeros
        out["grad_moving/" + v.var_name] = gr if gr is not None else np.zeros(tuple(v.t.shape))
    assert not g.update_log, "moving statistics must not move when is_training is False"

    # ---- sampling on batch statistics, as sample_multithread does (train_noise_flow.py:165-167, is_training: True)
    before = var_snapshot()
    xs = sess.run(x_sample_T1, feed_dict=fd(True))
    out["sample_batch_T1"] = xs
    after = var_snapshot()
    for name in after:
        if not np.array_equal(before[name], after[name]):
            out["after_sample_batch/" + name] = after[name]
    saver.restore(sess, os.path.join(MODEL_DIR, "ckpt", "model.ckpt.best"))

    # ---- one training step, train_multithread (train_noise_flow.py:62-71): sess.run([train_op, loss, sd_z], training)
    before = var_snapshot()
    g.update_log.clear()
    _, loss1, sdz1 = sess.run([train_op, loss_val, sd_z], feed_dict=fd(True))
    out["train_loss"], out["train_sd_z"] = loss1, sdz1
    out["train_bn_updates"] = np.int64(len(g.update_log))
    for v in tv:                    # (minimize skips the None gradients: those variables do not move)
        out["grad_batch/" + v.var_name] = opt.last_grads.get(v.var_name, np.zeros(tuple(v.t.shape)))
    after = var_snapshot()
    for name in after:
        if not np.array_equal(before[name], after[name]):
            out["after_step/" + name] = after[name]
    # second step: Adam slots and the moved statistics in use
    _, loss2, sdz2 = sess.run([train_op, loss_val, sd_z], feed_dict=fd(True))
    out["train_loss_step2"] = loss2
    after2 = var_snapshot()
    for name in after2:
        if not np.array_equal(before[name], after2[name]):
            out["after_step2/" + name] = after2[name]
    return out


def wrapper_goldens():
    """The reference's NoiseFlowWrapper class itself: constructor (graph, Saver.restore) and sample_noise_nf."""
    tf.reset_default_graph()
    np.random.seed(0)
    n = 4
    _, y = synth_batch(n, cam=2, iso=100, seed=21)
    eps = np.random.RandomState(22).randn(n, 32, 32, 4).astype(np.float32)
    b1, b2, iso, cam = 0.000479, 0.000002, 100, 2
    w = RefWrapper(MODEL_DIR, sampling_temperature=0.6)
    out = {"y": y, "eps": eps, "iso": np.float32(iso), "cam": np.float32(cam), "b1": np.float64(b1), "b2": np.float64(b2),
           "temp": np.float64(w.temp), "var_names": np.array(list(tf.get_default_graph().vars))}
    owners = []                 # which coupling owns which template scope in THIS graph (bijector index -> scope name)
    for i, b in enumerate(w.nf_model.model[0]):
        fn = getattr(b, "_shift_and_log_scale_fn", None)
        if fn is not None:
            owners.append("%d:%s" % (i, fn.variable_scope.name))
    out["template_scopes"] = np.array(owners)
    before = var_snapshot()
    inject_eps(eps)
    out["sample"] = w.sample_noise_nf(y, b1, b2, iso, cam)
    after = var_snapshot()
    for name in after:
        if not np.array_equal(before[name], after[name]):
            out["after_call/" + name] = after[name]
    y2 = synth_batch(n, cam=2, iso=100, seed=23)[1]
    eps2 = np.random.RandomState(24).randn(n, 32, 32, 4).astype(np.float32)
    inject_eps(eps2)
    out["y_call2"], out["eps_call2"] = y2, eps2
    out["sample_call2"] = w.sample_noise_nf(y2, b1, b2, 800, 0)     # the moving statistics moved by call 1 do not matter
    return out


# (tag, arch, flow_permutation, cam, iso) -- every token `noise_flow_arch` parses (noise_flow_model.py:79-234).
# The legacy scale layers create their variables straight in scope `model` without reuse (cond_utils.py:43-45,103-111,
# 143-151,321-323,361-365,401-405), so the reference cannot build two layers that share a name: `sdn2|gain2`,
# `sdn3|gain3`, `sdn2|sdn3`, `gain|gain1`, any legacy token twice ... raise "Variable model/... already exists" (the shim
# reproduces that).  The combinations below are buildable in the reference.
ARCH_CASES = [
    ("sdn_gain", "sdn|unc|gain|unc", 1, 0, 400),
    ("sdn1_gain1", "sdn1|unc|gain1", 0, 1, 800),
    ("sdn2_gain", "sdn2|unc|gain", 1, 2, 1600),
    ("sdn_gain2", "sdn|gain2|unc", 1, 3, 400),
    ("sdn3_gain1", "sdn3|unc|gain1", 2, 3, 3200),
    ("sdn1_gain3", "sdn1|gain3|unc", 1, 4, 100),
    ("sdn4_gain4", "sdn4|unc|gain4|unc", 1, 4, 100),
    ("sdn6_gain4", "sdn6|unc|gain4", 0, 2, 800),
    ("sdn5_unknown_iso", "sdn5|unc|gain4", 1, 2, 250),
    ("sdn2_unknown_iso", "sdn2|gain1", 1, 0, 250),
    ("gain3_unknown_iso", "gain3|unc", 1, 1, 250),
    # wider coupling nets (`--width`, sidd/ArgParser.py:43): the CTA-per-patch kernels of csrc/nf_wide.cu
    ("width8", "sdn5|unc|gain4|unc", 1, 2, 400, 8),
    ("width16", "sdn5|unc|unc|gain4", 0, 0, 800, 16),
    ("width32", "sdn5|unc|gain4|unc", 1, 4, 100, 32),
]


# widths of the tensor-core kernels (csrc/nf_wide_tc.cu: 32 / 64 / 128 resident weights, 256 / 512 streamed weights), up to the
# reference's default `--width 512` (sidd/ArgParser.py:43).  Kept in their own file (ref_wide_cases.npz): the l_2/W tensors
# are W x W.  Few couplings at the large widths bound the fixture size.
WIDE_CASES = [
    ("width64", "sdn5|unc|gain4|unc", 1, 2, 100, 64),
    ("width128", "sdn5|unc|unc|gain4", 1, 1, 800, 128),
    ("width256", "sdn5|unc|gain4", 1, 3, 100, 256),
    ("width512", "unc|gain4", 1, 0, 400, 512),
]


def perturb(rng, width=4):
    """Fresh variables make every coupling the identity (W3 = 0): move everything off its initial value (activations
    kept O(1) at any width)."""
    w_std, last_std = 0.5 / np.sqrt(width / 4.0), (0.1 if width == 4 else 0.02)
    for name, v in tf.get_default_graph().vars.items():
        leaf = name.split("/")[-1]
        a = v.t.detach().numpy()
        if leaf.startswith(("P_matpar", "sign_S")):
            continue
        if name.endswith("/l_1/W") or name.endswith("/l_2/W"):
            new = rng.randn(*a.shape) * w_std
        elif name.endswith("/l_last/W"):
            new = rng.randn(*a.shape) * last_std
        elif leaf in ("b", "logs"):
            new = rng.randn(*a.shape) * 0.2
        elif leaf == "mean":
            new = rng.randn(*a.shape) * 0.1
        elif leaf == "var":
            new = 0.5 + rng.rand(*a.shape)
        elif leaf.startswith("rescaling_scale"):
            new = np.full(a.shape, 0.5)
        elif leaf.startswith(("L_vec", "U_vec", "log_S")):
            new = a + rng.randn(*a.shape) * 0.1
        else:                                           # scale-layer scalars / tables: small relative move
            new = a + rng.randn(*a.shape) * 0.05 * np.maximum(1.0, np.abs(a))
        v.load(new.astype(np.float32))                  # fp32-representable, like a checkpoint


def arch_case_goldens(tag, arch, perm, cam, iso, width=4):
    tf.reset_default_graph()
    np.random.seed(5)
    hps = ref_hps()
    hps.arch, hps.flow_permutation, hps.width = arch, perm, width
    # train_noise_flow.py:200-213 (`init_params`) -- the wrapper's hps_loader leaves npcam unset for other archs
    npcam = 1 if "sdn6" in arch else 3
    cam_i = np.ndarray([npcam, 5])
    cam_i[:, :] = 1.0
    gp = np.ndarray([5])
    gp[:] = -5.0
    hps.param_inits = (1.0, -5.0, 0.0, gp, cam_i)
    n = 2
    x, y = synth_batch(n, cam=2, iso=iso if (2, iso) in CAM_ISO_NLF else 800, seed=31)
    eps = np.random.RandomState(32).randn(n, 32, 32, 4).astype(np.float32)
    nlf0, nlf1 = 0.0012, 0.000004
    is_training = tf.placeholder(tf.bool, name='is_training')
    nf = NoiseFlow(hps.x_shape[1:], is_training, hps)
    a = dict(nlf0=c([nlf0]), nlf1=c([nlf1]), iso=c([float(iso)]), cam=c([float(cam)]))
    nll_t, sd_t = nf._loss(c(x), c(y), **a)
    with tf.variable_scope("model", reuse=True):
        z_t, obj_t = nf.inverse(c(x), tf.zeros([n]), yy=c(y), **a)
    xs_t = nf.sample(c(y), 0.6, c(y), a["nlf0"], a["nlf1"], a["iso"], a["cam"])
    perturb(np.random.RandomState(33), width)
    out = {"arch": np.array(arch), "flow_permutation": np.int64(perm), "width": np.int64(width), "x": x, "y": y, "eps": eps, "iso": np.float32(iso),
           "cam": np.float32(cam), "nlf0": np.float64(nlf0), "nlf1": np.float64(nlf1),
           "layer_names": np.array(nf.get_layer_names()),
           "num_params": np.int64(sum(int(np.prod(v.get_shape().as_list())) for v in tf.trainable_variables()))}
    for name, v in var_snapshot().items():
        out["var/" + name] = v.astype(np.float32)
    sess = tf.Session()
    inject_eps(eps)
    out["nll"], out["sd_z"], z, out["logdet"], xs = sess.run([nll_t, sd_t, z_t, obj_t, xs_t], feed_dict={is_training: False})
    out["z"], out["sample_T0.6"] = z.astype(np.float32), xs.astype(np.float32)       # fixture size; nll / logdet stay fp64
    out["nll_batch"], out["sd_z_batch"] = sess.run([nll_t, sd_t], feed_dict={is_training: True})
    return out


# legacy `revnet2d` models (noise_flow_model.py:237-392; hps.arch unset): (tag, sidd_cond, flow_permutation, append_* flags).
# The CUDA engine does not implement them; the goldens pin the oracle's restatement for the round that will.
# (The ISO-polynomial layers create `model/p1 ... q4` without reuse, cond_utils.py:13-35: two of them in one graph -- `fitSDN`
# at depth > 1, or `fitSDN` / append_sdn together with append_sdn2 -- cannot be built in the reference.)
LEGACY_CASES = [
    ("condY", "condY", 1, {}),
    ("condYG", "condYG", 0, {}),
    ("condXY_cY", "condXY", 1, {"append_cY": True}),
    ("condXYG_sdn", "condXYG", 1, {"append_sdn": True}),
    ("condSDN_sdnfirst", "condSDN", 0, {"append_sdn_first": True}),
    ("fitSDN_depth1", "fitSDN", 2, {"depth": 1}),
    ("condXY_sdn2", "condXY", 1, {"append_sdn2": True}),
    ("uncond", "uncond", 1, {}),
]


def legacy_case_goldens(tag, cond, perm, flags):
    tf.reset_default_graph()
    np.random.seed(7)
    hps = ref_hps()
    hps.arch, hps.depth, hps.sidd_cond, hps.flow_permutation = None, int(flags.get("depth", 2)), cond, perm
    for f in ("append_sdn2", "append_sdn_first", "append_cY", "append_sdn"):
        setattr(hps, f, bool(flags.get(f, False)))
    n, iso, cam, nlf0, nlf1 = 2, 100.0, 2.0, 0.0012, 0.000004
    x, y = synth_batch(n, cam=2, iso=100, seed=61)
    eps = np.random.RandomState(62).randn(n, 32, 32, 4).astype(np.float32)
    is_training = tf.placeholder(tf.bool, name='is_training')
    nf = NoiseFlow(hps.x_shape[1:], is_training, hps)
    a = dict(nlf0=c([nlf0]), nlf1=c([nlf1]), iso=c([iso]), cam=c([cam]))
    nll_t, sd_t = nf._loss(c(x), c(y), **a)
    if cond == "uncond":
        xs_t = nf.sample(c(y), 0.6)
    else:
        xs_t = nf.sample(c(y), 0.6, c(y), a["nlf0"], a["nlf1"], a["iso"], a["cam"])
    perturb(np.random.RandomState(63))
    out = {"sidd_cond": np.array(cond), "flow_permutation": np.int64(perm), "x": x, "y": y, "eps": eps,
           "iso": np.float32(iso), "cam": np.float32(cam), "nlf0": np.float64(nlf0), "nlf1": np.float64(nlf1),
           "flags": np.array([f for f in sorted(flags) if f != "depth" and flags[f]]), "depth": np.int64(hps.depth), "layer_names": np.array(nf.get_layer_names()),
           "num_params": np.int64(sum(int(np.prod(v.get_shape().as_list())) for v in tf.trainable_variables()))}
    for name, v in var_snapshot().items():
        out["var/" + name] = v.astype(np.float32)
    sess = tf.Session()
    inject_eps(eps)
    out["nll"], out["sd_z"], xs = sess.run([nll_t, sd_t, xs_t], feed_dict={is_training: False})
    out["sample_T0.6"] = xs.astype(np.float32)
    out["nll_batch"], out["sd_z_batch"] = sess.run([nll_t, sd_t], feed_dict={is_training: True})
    return out


def squeeze_goldens():
    rng = np.random.RandomState(41)
    out = {}
    x = rng.randint(0, 1 << 20, size=(2, 8, 12, 3)).astype(np.float64)     # exact in fp32
    out["x"] = x.astype(np.float32)
    for factor in (1, 2):
        for kind in ("chessboard", "patch", "bogus"):
            s = ref_squeeze2d(c(x), factor, kind)
            out["squeeze_%d_%s" % (factor, kind)] = s.numpy().astype(np.float32)
            if factor == 1 or s.numpy().shape[-1] % 4 == 0:
                out["unsqueeze_%d_%s" % (factor, kind)] = ref_unsqueeze2d(s, factor, kind).numpy().astype(np.float32)
    return out


def metrics_goldens():
    """The evaluation metrics the training driver logs next to the NLL, from the reference's own numpy code:
    `sidd_utils.get_histogram / kl_div_forward / kl_div_inverse / kl_div_sym / kl_div_3_data` (:1202-1274) with the 66-bin
    edges of `kldiv_patch_set` (:1044-1046), `pack_raw / unpack_raw` (:732-764), and
    `PatchStatsCalculator.calc_baselines` (PatchStatsCalculator.py:92-121) fed through a queue as the driver does."""
    import queue
    import tempfile
    from types import SimpleNamespace
    import sidd.sidd_utils as su
    from sidd.PatchStatsCalculator import PatchStatsCalculator
    rng = np.random.RandomState(51)
    out = {}
    x, y = synth_batch(6, cam=2, iso=800, seed=52)
    xs = (x * 1.3 + rng.randn(*x.shape).astype(np.float32) * 0.004).astype(np.float32)       # a "sampled" noise batch
    out["x"], out["y"], out["x_sampled"] = x, y, xs
    bw = 0.2 / 64
    bin_edges = np.concatenate(([-1000.0], np.arange(-0.1, 0.1 + 1e-9, bw), [1000.0]), axis=0)   # sidd_utils.py:1044-1046
    out["bin_edges"] = bin_edges
    hp, centers = su.get_histogram(x, bin_edges=bin_edges)
    hq, _ = su.get_histogram(xs, bin_edges=bin_edges)
    out["hist_p"], out["hist_q"], out["bin_centers"] = hp, hq, centers
    out["kl_forward"], out["kl_inverse"], out["kl_sym"] = su.kl_div_forward(hp, hq), su.kl_div_inverse(hp, hq), su.kl_div_sym(hp, hq)
    out["kl_3_data_edges"] = np.array(su.kl_div_3_data(x, xs, bin_edges=bin_edges))
    out["kl_3_data_default"] = np.array(su.kl_div_3_data(np.abs(x) * 5, np.abs(xs) * 5))         # default 1000 bins on [0, 1]
    h1000, c1000 = su.get_histogram(np.abs(x) * 5)
    out["hist_default_1000"] = h1000
    raw = rng.rand(8, 12).astype(np.float32)
    out["bayer"], out["packed"] = raw, su.pack_raw(raw.copy())
    out["unpacked"] = su.unpack_raw(out["packed"])
    # baselines: two test minibatches through the queue, as initialize_data_stats_queues_baselines_histograms does
    with tempfile.TemporaryDirectory() as tmp:
        psc = PatchStatsCalculator(None, patch_height=32, n_channels=4, save_dir=tmp, file_postfix="", n_threads=1,
                                   hps=SimpleNamespace(test_its=2))
        psc.stats["sc_in_vr"] = np.float64(np.var(x))
        q = queue.Queue()
        nlf0, nlf1 = 0.003696, 0.000002
        x64, y64 = x.astype(np.float64), y.astype(np.float64)      # the minibatch sampler allocates float64 (MiniBatchSampler.py:54-55)
        q.put({"_x": x64[:3], "_y": y64[:3], "nlf0": nlf0, "nlf1": nlf1})
        q.put({"_x": x64[3:], "_y": y64[3:], "nlf0": nlf0, "nlf1": nlf1})
        nll_gauss, _, nll_sdn, _ = psc.calc_baselines(q)
    out["var_gauss"], out["nlf0"], out["nlf1"] = np.float64(np.var(x)), np.float64(nlf0), np.float64(nlf1)
    out["nll_gauss_mean"], out["nll_sdn_mean"] = np.float64(nll_gauss), np.float64(nll_sdn)
    out["bpd_of_nll_sdn"] = np.float64(su.bpd(nll_sdn, 256, 4096))
    return out


def save(name, d):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **d)
    print("wrote %s: %d arrays, %.0f KB" % (path, len(d), os.path.getsize(path) / 1024))


def main():
    tf1_shim.checkpoint_reader = load_checkpoint
    wide = {}
    for case in WIDE_CASES:
        for k, v in arch_case_goldens(*case).items():
            wide[case[0] + "::" + k] = v
    save("ref_wide_cases.npz", wide)
    if "--wide-only" in sys.argv:
        return
    save("ref_training_graph.npz", training_graph_goldens())
    save("ref_wrapper_graph.npz", wrapper_goldens())
    cases = {}
    for case in ARCH_CASES:
        tag = case[0]
        for k, v in arch_case_goldens(*case).items():
            cases[tag + "::" + k] = v
    save("ref_arch_cases.npz", cases)
    legacy = {}
    for tag, cond, perm, flags in LEGACY_CASES:
        for k, v in legacy_case_goldens(tag, cond, perm, flags).items():
            legacy[tag + "::" + k] = v
    save("ref_legacy_cases.npz", legacy)
    save("ref_squeeze.npz", squeeze_goldens())
    save("ref_metrics.npz", metrics_goldens())


if __name__ == "__main__":
    main()
