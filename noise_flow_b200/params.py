"""Host-side model assembly: TF-style variable store, arch parsing, LU construction, per-(camera, ISO)
scale tables.  Pure numpy (float64 arithmetic on the handful of scalars, float32 storage) -- no GPU
needed, so everything here is covered by the CPU test-suite.

Reference being mirrored (paths relative to the reference repo):
  * ``borealisflows/noise_flow_model.py:71-235``  arch string -> bijector list, scopes and names
  * ``borealisflows/noise_flow_model.py:237-392`` legacy ``revnet2d`` models (``hps.arch`` unset): clean-image-conditioned
    couplings, ISO-conditioned templates (``layers.py:501-547,616-648``), ISO-polynomial scale layers (``cond_utils.py:11-38``)
  * ``borealisflows/matrix_param.py:31-140``       LU parameterisation of the 1x1 conv
  * ``borealisflows/noise_flow_layers/cond_utils.py``  scale functions of the sdn*/gain* layers
  * ``borealisflows/layers.py:271-273,598-599,662-673,382-387``  initialisers
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

ISO_VALS = (100.0, 400.0, 800.0, 1600.0, 3200.0)          # cond_utils.py:224
CAM_NAMES = ("IP", "GP", "S6", "N6", "G4")                 # sidd/sidd_utils.py:262, cond_utils.py:212
N_STD_ROWS = 25                                            # row = cam * 5 + iso_index
MAX_ROWS = 32                                              # NF_MAX_ROWS in csrc/nf_params.h
BN_EPS = 1e-4                                              # layers.py:378

SCALE_SDN, SCALE_GAIN = 1, 2                               # include/noiseflow_b200.h
SDN_TOKENS = ("sdn", "sdn1", "sdn2", "sdn3", "sdn4", "sdn5", "sdn6", "camsdn", "sdngain", "fitsdngain2")
COUPLING_X, COUPLING_XY, COUPLING_Y = 0, 1, 2              # include/noiseflow_b200.h NF_COUPLING_*
GAIN_TOKENS = ("gain", "gain1", "gain2", "gain3", "gain4")
# Gain, GainEx1, GainEx3 return log(scale) without the sum over the 4096 dims (AffineCouplingGain.py:86,96,111,125)
NO_FULL_SUM = ("gain", "gain1", "gain3")


# ================================================================================================
# variables
# ================================================================================================
class VariableStore:
    """``{tf_variable_name: float32 ndarray}`` with ``tf.get_variable`` semantics: an existing name is
    returned as is (= ``Saver.restore``), a missing one is created from its reference initialiser."""

    def __init__(self, values: Optional[Dict[str, np.ndarray]] = None, seed: int = 0):
        self.vars: "OrderedDict[str, np.ndarray]" = OrderedDict()
        self.trainable: Dict[str, bool] = {}
        self.created: List[str] = []
        self.rng = np.random.RandomState(seed)
        self._scope_counts: Dict[str, int] = {}
        if values:
            for k, v in values.items():
                self.vars[k] = np.array(v, dtype=np.float32)

    def get(self, name: str, shape, init, trainable: bool = True) -> np.ndarray:
        if name not in self.vars:
            val = init() if callable(init) else init
            self.vars[name] = np.broadcast_to(np.asarray(val, dtype=np.float32), tuple(shape)).copy()
            self.created.append(name)
        v = self.vars[name]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %s has shape %s, expected %s" % (name, v.shape, tuple(shape)))
        self.trainable.setdefault(name, trainable)
        return v

    def unique_scope(self, prefix: str, default_name: str) -> str:
        """``tf.variable_scope(None, default_name=...)``: name, name_1, name_2, ... in first-use order."""
        key = prefix + "/" + default_name
        n = self._scope_counts.get(key, 0)
        self._scope_counts[key] = n + 1
        return key if n == 0 else "%s_%d" % (key, n)

    def num_trainable(self) -> int:
        return int(sum(self.vars[k].size for k, t in self.trainable.items() if t))


# ================================================================================================
# matrix_param.py: fill_triangular ordering and the LU construction
# ================================================================================================
def fill_triangular(v: np.ndarray, upper: bool = False) -> np.ndarray:
    """TFP ``fill_triangular`` as used by matrix_param.py:44 (spiral order, docstring example:
    [1..6] -> [[4,0,0],[6,5,0],[3,2,1]] lower, [[1,2,3],[0,5,6],[0,0,4]] upper)."""
    v = np.asarray(v)
    m = v.shape[-1]
    n = int(round((math.sqrt(8 * m + 1) - 1) / 2))
    if n * (n + 1) // 2 != m:
        raise ValueError("vector length %d is not triangular" % m)
    if upper:
        return np.triu(np.concatenate([v, v[n:][::-1]]).reshape(n, n))
    return np.tril(np.concatenate([v[n:], v[::-1]]).reshape(n, n))


def vec2stricttri(vec: np.ndarray, upper: bool) -> np.ndarray:
    """matrix_param.py:31-57."""
    base = fill_triangular(vec, upper)
    k = base.shape[0]
    out = np.zeros((k + 1, k + 1), dtype=base.dtype)
    if upper:
        out[:k, 1:] = base
    else:
        out[1:, :k] = base
    return out


def stricttri2vec(mat: np.ndarray, upper: bool) -> np.ndarray:
    """matrix_param.py:60-97 (inverse of :func:`vec2stricttri`)."""
    mat = np.asarray(mat)
    n = mat.shape[0] - 1
    m = n * (n + 1) // 2
    pos = vec2stricttri(np.arange(1, m + 1, dtype=np.float64), upper).astype(np.int64)
    out = np.zeros(m, dtype=mat.dtype)
    out[pos[pos > 0] - 1] = mat[pos > 0]
    return out


def lu_to_matrix(p, l_vec, u_vec, log_s, sign_s) -> Tuple[np.ndarray, np.ndarray, float]:
    """matrix_param.py:117-140: ``A = P L U``, ``A_inv = U^-1 L^-1 P^T``, ``log|det| = sum(log_S)``."""
    import scipy.linalg
    p = np.asarray(p, np.float64)
    log_s = np.asarray(log_s, np.float64)
    n = p.shape[0]
    l = vec2stricttri(np.asarray(l_vec, np.float64), upper=False) + np.eye(n)
    u = vec2stricttri(np.asarray(u_vec, np.float64), upper=True) + np.diag(np.asarray(sign_s, np.float64) * np.exp(log_s))
    a = p @ (l @ u)
    y = scipy.linalg.solve_triangular(l, p.T, lower=True, unit_diagonal=True)
    a_inv = scipy.linalg.solve_triangular(u, y, lower=False)
    return a, a_inv, float(log_s.sum())


# ================================================================================================
# layer specs (what gets handed to the C-ABI builder)
# ================================================================================================
@dataclass
class LayerSpec:
    kind: str                 # 'conv1x1' | 'permute' | 'coupling' | 'scale'
    name: str                 # bijector name as in get_layer_names() / hps.txt
    scope: str                # level0/bijector{i}
    token: str = ""           # arch token for scale layers
    data: dict = field(default_factory=dict)


def _sigmoid(v):
    return 1.0 / (1.0 + math.exp(-float(v)))


class ModelSpec:
    """Bijector list of ``NoiseFlow.noise_flow_arch`` plus the variable store behind it."""

    def __init__(self, hps, variables: Optional[Dict[str, np.ndarray]] = None, seed: int = 0):
        self.hps = hps
        self.store = VariableStore(variables, seed)
        if getattr(hps, "n_levels", 1) != 1:
            raise NotImplementedError("n_levels > 1 (split2d) is outside the hot path")
        if getattr(hps, "squeeze_factor", 1) != 1:
            # the reference's prior is built on the un-squeezed shape (noise_flow_model.py:488-493), so
            # squeeze_factor != 1 fails shape-wise in logp(); only the stand-alone squeeze ops are provided.
            raise NotImplementedError("squeeze_factor != 1 does not run end-to-end in the reference either")
        self.x_shape = [32, 32, 4]
        self.width = int(hps.width)
        self.layers: List[LayerSpec] = []
        self._template_scopes_assigned = False
        if getattr(hps, "arch", None):                                       # noise_flow_model.py:63-68
            self._parse_arch(hps.arch, hps.flow_permutation)
        else:
            self._parse_revnet2d(hps.flow_permutation)
        self.has_iso_templates = any(l.kind == "coupling" and l.data.get("iso") for l in self.layers)

    # ---- noise_flow_model.py:71-235 ---------------------------------------------------------------
    def _parse_arch(self, arch: str, flow_permutation: int):
        st = self.store
        ic = self.x_shape[-1]
        for i, lyr in enumerate(arch.split("|")):
            scope = "level0/bijector%d" % i
            if lyr == "unc":
                if flow_permutation == 0:
                    self.layers.append(LayerSpec("permute", "permute", scope,
                                                 data={"perm": list(range(ic))[::-1]}))          # :80-84
                elif flow_permutation == 1:
                    name = "Conv2d_1x1_%d" % i                                                   # :85-90
                    self._create_conv1x1(scope + "/" + name, "conv2d_1x1_%d_0" % i, ic)
                    self.layers.append(LayerSpec("conv1x1", name, scope,
                                                 data={"vscope": scope + "/" + name, "pname": "conv2d_1x1_%d_0" % i}))
                st.get(scope + "/rescaling_scale0", (), 1e-4)                                    # layers.py:271-273
                self.layers.append(LayerSpec("coupling", "unc_%d" % i, scope, data={"template": None}))
            elif lyr in SDN_TOKENS or lyr in GAIN_TOKENS:
                st.get(scope + "/rescaling_scale0", (), 1e-4)     # created by every scale bijector, never used
                pre = "gain" if lyr in GAIN_TOKENS else "sdn"
                self.layers.append(LayerSpec("scale", "%s_%d" % (pre, i), scope, token=lyr))
            # any other token falls through the reference's if/elif chain and adds nothing

    # ---- noise_flow_model.py:237-392 ---------------------------------------------------------------
    def _parse_revnet2d(self, flow_permutation: int):
        """Legacy models: ``depth`` x ([permutation] + a bijector chosen by ``hps.sidd_cond``) plus the ``append_*`` layers."""
        st, h = self.store, self.hps
        ic = self.x_shape[-1]
        depth = int(h.depth)
        name = "level0"

        def scale(token, scope, lname):
            st.get(scope + "/rescaling_scale0", (), 1e-4)
            self.layers.append(LayerSpec("scale", lname, scope, token=token))

        def coupling(mode, iso, scope, lname):
            st.get(scope + "/rescaling_scale0", (), 1e-4)
            self.layers.append(LayerSpec("coupling", lname, scope, data={"template": None, "mode": mode, "iso": iso}))

        if getattr(h, "append_sdn2", False):                                                         # :243-253
            scale("fitsdngain2", name + "/bijector_sdn2", "ac_fitSdnGain2_%d" % depth)
        if getattr(h, "append_sdn_first", False):                                                    # :255-265
            scale("sdngain", name + "/bijector_sdn", "ac_fitSdnGain_%d" % depth)
        if getattr(h, "append_cY", False):                                                           # :267-279
            coupling(COUPLING_Y, False, name + "/bijector_cy", "ac_cY_first")
        for i in range(depth):                                                                       # :280-378
            scope = "%s/bijector%d" % (name, i)
            if flow_permutation == 0:
                self.layers.append(LayerSpec("permute", "permute", scope, data={"perm": list(range(ic))[::-1]}))
            elif flow_permutation == 1:
                lname = "Conv2d_1x1_%d" % i
                self._create_conv1x1(scope + "/" + lname, "conv2d_1x1_%d_0" % i, ic)
                self.layers.append(LayerSpec("conv1x1", lname, scope,
                                             data={"vscope": scope + "/" + lname, "pname": "conv2d_1x1_%d_0" % i}))
            cond = getattr(h, "sidd_cond", "uncond")
            if cond == "condY":
                coupling(COUPLING_Y, False, scope, "ac_cY_%d" % i)
            elif cond == "condYG":
                coupling(COUPLING_Y, True, scope, "ac_cYG_%d" % i)
            elif cond == "condXY":
                coupling(COUPLING_XY, False, scope, "ac_cXY_%d" % i)
            elif cond == "condXYG":
                coupling(COUPLING_XY, True, scope, "ac_cXYG_%d" % i)
            elif cond == "condSDN":
                scale("camsdn", scope, "ac_cSDN_%d" % i)
            elif cond == "fitSDN":
                scale("sdngain", scope, "ac_fitSDN_%d" % i)
            else:                                                                                    # uncond | unc_sdn
                coupling(COUPLING_X, False, scope, "ac_unc_%d" % i)
        if getattr(h, "append_sdn", False):                                                          # :379-390
            scale("sdngain", "%s/bijector%d" % (name, depth), "ac_fitSDN_%d" % depth)

    def _create_conv1x1(self, vscope: str, pname: str, n: int):
        """Conv2d1x1._init_weights (layers.py:92-100) + matrix_param_lu initialisers (matrix_param.py:100-126)."""
        st = self.store
        decomp = getattr(self.hps, "decomp", "LU")
        if decomp != "LU":
            raise NotImplementedError("decomp=%s (only the shipped 'LU' parameterisation is implemented)" % decomp)
        cache = {}

        def lu():
            if not cache:
                import scipy.linalg
                w = scipy.linalg.qr(st.rng.randn(n, n))[0].astype("float32")                    # layers.py:95
                p_, l_, u_ = scipy.linalg.lu(w)                                                  # matrix_param.py:102
                cache.update(p=p_, l=l_, u=u_)
            return cache

        st.get("%s/P_matpar_lu_%s" % (vscope, pname), (n, n), lambda: lu()["p"], trainable=False)
        st.get("%s/sign_S_matpar_lu_%s" % (vscope, pname), (n,), lambda: np.sign(np.diag(lu()["u"])), trainable=False)
        st.get("%s/log_S_matpar_lu_%s" % (vscope, pname), (n,), lambda: np.log(np.abs(np.diag(lu()["u"]))))
        nv = n * (n - 1) // 2
        st.get("%s/L_vec_matpar_lu_%s" % (vscope, pname), (nv,), lambda: stricttri2vec(lu()["l"], upper=False))
        st.get("%s/U_vec_matpar_lu_%s" % (vscope, pname), (nv,),
               lambda: stricttri2vec(np.triu(lu()["u"], k=1), upper=True))

    # ---- tf.make_template scope naming (layers.py:498): assigned in FIRST-CALL order ------------------
    def assign_template_scopes(self, first_call: str = "inverse"):
        """``first_call='inverse'``: the graph traced ``loss``/``inverse`` first (train_noise_flow.py:302) so
        couplings get ``model/real_nvp_conv_template``, ``..._1``, ... in data->latent order.
        ``first_call='forward'``: only ``sample`` was traced (NoiseFlowWrapper.py:64), so the names are
        handed out in latent->data order."""
        if self._template_scopes_assigned:
            return
        cps = [l for l in self.layers if l.kind == "coupling"]
        order = cps if first_call == "inverse" else list(reversed(cps))
        for l in order:
            iso = bool(l.data.get("iso"))
            l.data["template"] = self.store.unique_scope("model", "real_nvp_conv_template_iso" if iso else "real_nvp_conv_template")
            mode = l.data.get("mode", COUPLING_X)
            self._create_template(l.data["template"], cin={COUPLING_X: 2, COUPLING_XY: 6, COUPLING_Y: 4}[mode],
                                  cout=8 if mode == COUPLING_Y else 4, iso=iso)
        self._template_scopes_assigned = True

    def _create_template(self, s: str, cin: int = 2, cout: int = 4, iso: bool = False):
        """real_nvp_conv_template (layers.py:452-498) / real_nvp_conv_template_iso (:501-547, conv2d_iso :616-648)."""
        st, w = self.store, self.width
        std = w / 512 * 0.05                                                                     # layers.py:598-599
        def conv(nm, shp):           # variables in the reference's creation order: conv (W, b | B1, B2, C1, C2), then its BatchNorm
            if iso:                                                                              # :631-647, init_sd = 0.05
                for v in ("B1", "B2"):
                    st.get("%s/%s/%s" % (s, nm, v), shp, lambda: st.rng.randn(*shp) * 0.05)
                for v in ("C1", "C2"):
                    st.get("%s/%s/%s" % (s, nm, v), (1, 1, 1, w), lambda: st.rng.randn(1, 1, 1, w) * 0.05)
            else:
                st.get("%s/%s/W" % (s, nm), shp, lambda: st.rng.randn(*shp) * std)
                st.get("%s/%s/b" % (s, nm), (1, 1, 1, w), 0.0)

        conv("l_1", (3, 3, cin, w))
        st.get(s + "/bn_nvp_conv_1/mean", (w,), 0.0, trainable=False)                            # layers.py:382-387
        st.get(s + "/bn_nvp_conv_1/var", (w,), 1.0, trainable=False)
        conv("l_2", (1, 1, w, w))
        st.get(s + "/bn_nvp_conv_2/mean", (w,), 0.0, trainable=False)
        st.get(s + "/bn_nvp_conv_2/var", (w,), 1.0, trainable=False)
        st.get(s + "/l_last/W", (3, 3, w + 1, cout), 0.0)                                        # layers.py:662-663
        st.get(s + "/l_last/b", (1, 1, 1, cout), 0.0)
        st.get(s + "/l_last/logs", (1, cout), 0.0)

    def create_scale_variables(self):
        """Scale-layer variables are created when the bijector is first called under scope 'model'."""
        for l in self.layers:
            if l.kind == "scale":
                scale_row(l.token, self.store, self.hps, cam=0.0, iso=100.0, nlf0=1.0, nlf1=1.0)

    # ---- views ------------------------------------------------------------------------------------------
    def get_layer_names(self) -> List[str]:
        return [l.name for l in self.layers]

    def conv1x1_matrices(self, l: LayerSpec):
        v, s, p = self.store.vars, l.data["vscope"], l.data["pname"]
        return lu_to_matrix(v["%s/P_matpar_lu_%s" % (s, p)], v["%s/L_vec_matpar_lu_%s" % (s, p)],
                            v["%s/U_vec_matpar_lu_%s" % (s, p)], v["%s/log_S_matpar_lu_%s" % (s, p)],
                            v["%s/sign_S_matpar_lu_%s" % (s, p)])

    def coupling_weights(self, l: LayerSpec, iso: float = 100.0) -> Dict[str, np.ndarray]:
        """Reference-shaped tensors of a coupling's net.  ISO-conditioned templates (conv2d_iso, layers.py:616-648):
        ``w = B1 * iso[0] + B2``, ``b = C1 * iso[0] + C2`` -- the effective weights for the call's ISO (fp32 arithmetic as TF)."""
        v, s = self.store.vars, l.data["template"]
        if s is None:
            raise RuntimeError("template scopes not assigned yet (call assign_template_scopes)")
        c = lambda a: np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        if l.data.get("iso"):
            i0 = np.float32(iso)
            l1w, l1b = v[s + "/l_1/B1"] * i0 + v[s + "/l_1/B2"], v[s + "/l_1/C1"] * i0 + v[s + "/l_1/C2"]
            l2w, l2b = v[s + "/l_2/B1"] * i0 + v[s + "/l_2/B2"], v[s + "/l_2/C1"] * i0 + v[s + "/l_2/C2"]
        else:
            l1w, l1b, l2w, l2b = v[s + "/l_1/W"], v[s + "/l_1/b"], v[s + "/l_2/W"], v[s + "/l_2/b"]
        return dict(l1_w=c(l1w), l1_b=c(l1b), bn1_mean=c(v[s + "/bn_nvp_conv_1/mean"]),
                    bn1_var=c(v[s + "/bn_nvp_conv_1/var"]), l2_w=c(l2w), l2_b=c(l2b),
                    bn2_mean=c(v[s + "/bn_nvp_conv_2/mean"]), bn2_var=c(v[s + "/bn_nvp_conv_2/var"]),
                    last_w=c(v[s + "/l_last/W"]), last_b=c(v[s + "/l_last/b"]), last_logs=c(v[s + "/l_last/logs"]),
                    rescaling_scale=float(v[l.scope + "/rescaling_scale0"]))

    def scale_table(self, l: LayerSpec, extra: Optional[List[Tuple[float, float, float, float]]] = None) -> np.ndarray:
        """[n_rows, 2] float32: rows 0..24 = standard (cam, iso) grid, then ``extra`` (cam, iso, nlf0, nlf1)."""
        rows = []
        for cam in range(5):
            for iso in ISO_VALS:
                rows.append(scale_row(l.token, self.store, self.hps, cam=float(cam), iso=iso, nlf0=0.0, nlf1=1.0))
        for (cam, iso, n0, n1) in (extra or []):
            rows.append(scale_row(l.token, self.store, self.hps, cam=cam, iso=iso, nlf0=n0, nlf1=n1))
        return np.asarray(rows, dtype=np.float32)


# ================================================================================================
# cond_utils.py: every scale layer reduces to scale = sqrt(a*y + b) (sdn*) or scale = g (gain*)
# ================================================================================================
def _iso_ladder(store: VariableStore, fmt: str, iso: float, init: float) -> float:
    """Nested tf.cond ladders over iso[0] (e.g. cond_utils.py:71-89): an unknown ISO takes the 800 branch."""
    vals = {int(v): float(store.get("model/" + fmt % int(v), (1,), init)[0]) for v in ISO_VALS}
    return vals[int(iso)] if float(iso) in ISO_VALS else vals[800]


def _iso_onehot(gain_params: np.ndarray, iso: float) -> float:
    """tf.where(tf.equal(iso_vals, iso)) -> one_hot -> reduce_sum (cond_utils.py:226-228): unknown ISO -> 0."""
    for k, v in enumerate(ISO_VALS):
        if float(iso) == v:
            return float(gain_params[k])
    return 0.0


def scale_row(token: str, store: VariableStore, hps, cam: float, iso: float, nlf0: float, nlf1: float):
    """(a, b) for sdn tokens, (g, 0) for gain tokens, for one (cam, iso[, nlf0, nlf1]) conditioning class."""
    g = store.get
    gain_init = float(getattr(hps, "gain_init", 0.0))
    if token in ("sdngain", "fitsdngain2"):                                                     # cond_utils.py:11-38 (ISO polynomials)
        e = lambda n: math.exp(float(g("model/" + n, (1,), -6.0)[0]))
        i0 = float(iso)
        if token == "sdngain":                                                                   # sdn_iso_model_params_3
            return e("p1") * i0 ** 2 + e("p2") * i0 + e("p3"), e("q1") * i0 ** 3 + e("q2") * i0 ** 2 + e("q3") * i0 + e("q4")
        return e("p2") * i0 + e("p3"), e("q2") * i0 ** 2 + e("q3") * i0 + e("q4")               # sdn_iso_model_params_2
    if token == "sdn":                                                                           # cond_utils.py:41-52
        return _sigmoid(g("model/b1", (1,), -3.0)[0]), _sigmoid(g("model/b2", (1,), 3.0)[0])
    if token == "sdn1":                                                                          # :55-97
        c = 1e-2
        r_gain = math.exp(c * _iso_ladder(store, "r_gain_param_%05d", iso, 0.0 / c)) * iso
        return _sigmoid(g("model/b1", (1,), -3.0)[0]) / r_gain, _sigmoid(g("model/b2", (1,), 3.0)[0])
    if token in ("sdn2", "sdn3"):                                                                # :100-162
        c = 1e-1
        gain = math.exp(c * _iso_ladder(store, "gain_param_%05d", iso, gain_init / c)) * iso
        b1, b2 = _sigmoid(g("model/b1", (1,), -3.0)[0]), _sigmoid(g("model/b2", (1,), 3.0)[0])
        if token == "sdn2":
            return b1, gain * b2                       # sqrt(gain*(b1*y/gain + b2))
        return gain * b1, gain * gain * b2             # gain*sqrt(b1*y/gain + b2)
    if token == "sdn4":                                                                          # :165-187
        s = "model/sdn_gain"
        g(s + "/gain_val", (1,), 1.0)
        gain = math.exp(_iso_onehot(g(s + "/gain_params", (5,), gain_init), iso)) * iso
        return math.exp(float(g(s + "/beta1", (1,), gain_init)[0])) / gain, math.exp(float(g(s + "/beta2", (1,), 0.0)[0]))
    if token in ("sdn5", "sdn6"):                                                                # :205-276
        (c_i, beta1_i, beta2_i, gain_params_i, cam_params_i) = hps.param_inits
        npc = 3 if token == "sdn5" else 1
        s = "model/sdn_gain"
        cam_params = g(s + "/cam_params", (npc, 5), lambda: np.asarray(cam_params_i)[:npc])
        if float(cam) not in (0.0, 1.0, 2.0, 3.0, 4.0):
            raise IndexError("camera id %r not in 0..4 (reference: tf.where(...)[0] on an empty tensor)" % (cam,))
        ocp = np.exp(c_i * cam_params[:, int(cam)].astype(np.float64))                           # :216-220
        g(s + "/gain_val", (1,), 1.0)                                                            # :223
        gsel = _iso_onehot(g(s + "/gain_params", (5,), lambda: np.asarray(gain_params_i)), iso)
        beta1 = float(g(s + "/beta1", (1,), beta1_i)[0])
        beta2 = float(g(s + "/beta2", (1,), beta2_i)[0])
        if token == "sdn5":
            gain = math.exp(c_i * gsel * ocp[2]) * iso                                           # :230
            return math.exp(c_i * beta1 * ocp[0]) / gain, math.exp(c_i * beta2 * ocp[1])         # :236-238
        gain = math.exp(c_i * gsel * ocp[0]) * iso                                               # :267
        return math.exp(c_i * beta1) / gain, math.exp(c_i * beta2)                               # :273-275
    if token == "camsdn":                                                                        # AffineCouplingCamSdn.py:47
        return float(nlf0), float(nlf1)
    if token == "gain":                                                                          # cond_utils.py:319-330
        return _sigmoid(g("model/g1", (1,), -3.0)[0]) * iso + _sigmoid(g("model/g2", (1,), 3.0)[0]), 0.0
    if token == "gain1":                                                                         # :333-351
        c = 1e-5
        return math.exp(c * float(g("model/g1", (1,), -5.0 / c)[0])) * iso + math.exp(c * float(g("model/g2", (1,), 0.0)[0])), 0.0
    if token == "gain2":                                                                         # :354-395
        c = 1e-1
        return math.exp(c * _iso_ladder(store, "gain_param_%05d", iso, gain_init / c)) * iso, 0.0
    if token == "gain3":                                                                         # :398-429
        c = 1e-5
        return math.exp(c * _iso_ladder(store, "gain_param_%05d", iso, -5.0 / c)), 0.0
    if token == "gain4":                                                                         # :432-440
        return float(g("model/sdn_gain/gain_val", (1,), 1.0)[0]), 0.0
    raise ValueError("unknown scale token %r" % token)


def std_row(cam, iso) -> Optional[int]:
    """Row of the standard table for (cam, iso), or None if either is outside the grid."""
    c, i = float(cam), float(iso)
    if c in (0.0, 1.0, 2.0, 3.0, 4.0) and i in ISO_VALS:
        return int(c) * 5 + ISO_VALS.index(i)
    return None
