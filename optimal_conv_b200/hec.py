"""ctypes binding of libhec.so -- the same C ABI (include/hec.h) a cgo shim binds.

Thin by design: every method is one C-ABI call.  There is no CPU fallback: if the CUDA
library is missing or no device is present, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEC_LIB", os.path.join(_HERE, "libhec.so"))  # HEC_LIB: a build variant to measure (tools/)

u64p = C.POINTER(C.c_uint64)
u64pp = C.POINTER(u64p)
vp = C.c_void_p

HEC_OK, HEC_E_INVAL, HEC_E_CUDA, HEC_E_LEVEL, HEC_E_NOKEY, HEC_E_SCALE, HEC_E_UNSUPPORTED, HEC_E_NOMEM = (
    0, -1, -2, -3, -4, -5, -6, -7)
CONV_FUSED, CONV_OPLEVEL = 0, 1

# every symbol include/hec.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "hec_version", "hec_ctx_create", "hec_ctx_destroy", "hec_last_error", "hec_sync", "hec_timer_start",
    "hec_timer_stop_ms", "hec_launch_count", "hec_host_register", "hec_host_unregister", "hec_pt_upload", "hec_pt_free", "hec_encode_coeffs", "hec_encode_coeffs_many", "hec_pt_download", "hec_ct_upload", "hec_ct_download",
    "hec_ct_copy_new", "hec_ct_level", "hec_ct_scale", "hec_ct_set_scale", "hec_ct_free", "hec_swk_upload",
    "hec_swk_drop", "hec_mul_pt_new", "hec_mult_by_const", "hec_rescale", "hec_set_scale", "hec_add", "hec_add_new",
    "hec_sub_new", "hec_add_pt", "hec_rlk_upload", "hec_mul_relin_new", "hec_sub", "hec_drop_level", "hec_mul_by_pow2", "hec_mult_by_i", "hec_conjugate", "hec_add_const",
    "hec_ptdiag_upload", "hec_ptdiag_free", "hec_linear_transform", "hec_coeffs_to_slots", "hec_slots_to_coeffs", "hec_sub_sum", "hec_mod_up", "hec_bootstrap_ctos", "hec_bootstrap_stoc", "hec_bootstrapp", "hec_mult_by_int_and_add", "hec_evaluate_poly", "hec_evaluate_cheby", "hec_eval_relu", "hec_eval_relu_many", "hec_rotate_gal", "hec_rotate_new", "hec_rotate_hoisted", "hec_galois_for_rotation",
    "hec_ntt", "hec_keyswitch", "hec_moddown", "hec_conv_then_pack", "hec_conv_bl", "hec_ext_ctxt", "hec_keep_ctxt", "hec_plan_create", "hec_plan_run",
    "hec_plan_run_host", "hec_plan_submit_host", "hec_plan_wait", "hec_plan_span_begin", "hec_plan_span_end_ms",
    "hec_plan_profile", "hec_plan_kernel_names", "hec_plan_is_deferred", "hec_plan_destroy", "hec_plan_cache_size", "hec_float_quotient_threshold", "hec_ext_double_ctxt", "hec_conv_bn_relu",
]


class BtpParams(C.Structure):
    """hec_btp_params (include/hec.h)"""
    _fields_ = [("prescale", C.c_double), ("postscale", C.c_double), ("sinescale", C.c_double), ("sqrt2pi", C.c_double),
                ("sc_fac", C.c_double), ("message_ratio", C.c_double), ("params_scale", C.c_double),
                ("sin_type", C.c_int), ("sin_rescal", C.c_int), ("arcsine_deg", C.c_int),
                ("sine_qi", C.POINTER(C.c_uint64)), ("n_sine_qi", C.c_int),
                ("cheby", C.POINTER(C.c_double)), ("n_cheby", C.c_int), ("cheby_a", C.c_double), ("cheby_b", C.c_double)]


class LayerArgs(C.Structure):
    """hec_layer_args (include/hec.h)"""
    _fields_ = [("n_conv", C.c_int), ("pt_ker", C.POINTER(C.c_void_p) * 2), ("pt_bias", C.c_void_p * 2),
                ("max_ob", C.c_int), ("norm", C.c_int * 2), ("out_scale", C.c_double), ("pt_idx", C.POINTER(C.c_void_p)),
                ("conv_flags", C.c_int), ("pt_pre", C.c_void_p), ("pt_shift2", C.c_void_p), ("pt_post", C.c_void_p),
                ("pow", C.c_double), ("alpha", C.c_double), ("iter", C.c_int), ("btp", C.POINTER(BtpParams)),
                ("ctos_mats", C.POINTER(C.c_void_p)), ("n_ctos", C.c_int), ("stoc_mats", C.POINTER(C.c_void_p)), ("n_stoc", C.c_int),
                ("move_kind", C.c_int), ("keep_mask", C.c_void_p * 2),
                ("n_r", C.c_int * 2), ("rots_r", C.POINTER(C.c_int) * 2), ("pts_r", C.POINTER(C.c_void_p) * 2),
                ("n_m", C.c_int * 2), ("rots_m", C.POINTER(C.c_int) * 2), ("pts_m", C.POINTER(C.c_void_p) * 2),
                ("min_scale", C.c_double)]


MOVE_KEEP, MOVE_EXT, MOVE_EXT_DOUBLE = 0, 1, 2


class HecError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libhec error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libhec.so; fails loudly if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # build in-tree with nvcc when the toolchain is present; there is nothing else to fall back to
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_hec_build", os.path.join(os.path.dirname(_HERE), "__graft_entry__.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build_libhec()
        except Exception as e:  # noqa: BLE001
            raise ImportError("%s is not built and could not be built (%s): run "
                              "`python -c 'import __graft_entry__ as g; g.build()'`" % (LIB_PATH, e))
    L = C.CDLL(LIB_PATH)
    L.hec_version.restype = C.c_char_p
    L.hec_ctx_create.argtypes = [C.POINTER(vp), C.c_int, u64p, C.c_int, u64p, C.c_int, C.c_int]
    L.hec_ctx_destroy.argtypes = [vp]
    L.hec_ctx_destroy.restype = None
    L.hec_last_error.argtypes = [vp]
    L.hec_last_error.restype = C.c_char_p
    L.hec_sync.argtypes = [vp]
    L.hec_timer_start.argtypes = [vp]
    L.hec_timer_stop_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.hec_launch_count.argtypes = [vp]
    L.hec_launch_count.restype = C.c_uint64
    L.hec_host_register.argtypes = [vp, vp, C.c_size_t]
    L.hec_host_unregister.argtypes = [vp, vp]
    L.hec_pt_upload.argtypes = [vp, C.c_int, u64pp, C.c_double, C.POINTER(vp)]
    L.hec_encode_coeffs.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.POINTER(vp)]
    L.hec_encode_coeffs_many.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(vp)]
    L.hec_pt_download.argtypes = [vp, vp, u64pp]
    L.hec_pt_free.argtypes = [vp, vp]
    L.hec_pt_free.restype = None
    L.hec_ct_upload.argtypes = [vp, C.c_int, u64pp, u64pp, C.c_double, C.POINTER(vp)]
    L.hec_ct_download.argtypes = [vp, vp, u64pp, u64pp]
    L.hec_ct_copy_new.argtypes = [vp, vp, C.POINTER(vp)]
    L.hec_ct_level.argtypes = [vp]
    L.hec_ct_scale.argtypes = [vp]
    L.hec_ct_scale.restype = C.c_double
    L.hec_ct_set_scale.argtypes = [vp, C.c_double]
    L.hec_ct_set_scale.restype = None
    L.hec_ct_free.argtypes = [vp, vp]
    L.hec_ct_free.restype = None
    L.hec_swk_upload.argtypes = [vp, C.c_uint64, C.c_int, u64pp]
    L.hec_swk_drop.argtypes = [vp, C.c_uint64]
    L.hec_mul_pt_new.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.hec_mult_by_const.argtypes = [vp, vp, C.c_double]
    L.hec_rescale.argtypes = [vp, vp, C.c_double]
    L.hec_set_scale.argtypes = [vp, vp, C.c_double]
    L.hec_add.argtypes = [vp, vp, vp, vp]
    L.hec_add_new.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.hec_sub_new.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.hec_add_pt.argtypes = [vp, vp, vp]
    L.hec_rlk_upload.argtypes = [vp, C.c_int, u64pp]
    L.hec_sub.argtypes = [vp, vp, vp, vp]
    L.hec_drop_level.argtypes = [vp, vp, C.c_int]
    L.hec_mul_by_pow2.argtypes = [vp, vp, C.c_int]
    L.hec_mult_by_i.argtypes = [vp, vp, C.c_int]
    L.hec_conjugate.argtypes = [vp, vp, vp]
    L.hec_add_const.argtypes = [vp, vp, C.c_double]
    L.hec_mult_by_int_and_add.argtypes = [vp, vp, C.c_int64, vp]
    L.hec_evaluate_poly.argtypes = [vp, vp, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
    L.hec_ptdiag_upload.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_int), u64pp, C.POINTER(vp)]
    L.hec_ptdiag_free.argtypes = [vp, vp]
    L.hec_ptdiag_free.restype = None
    L.hec_linear_transform.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.hec_coeffs_to_slots.argtypes = [vp, vp, C.POINTER(vp), C.c_int, C.POINTER(vp), C.POINTER(vp)]
    L.hec_slots_to_coeffs.argtypes = [vp, vp, vp, C.POINTER(vp), C.c_int, C.POINTER(vp)]
    L.hec_bootstrap_stoc.argtypes = [vp, vp, vp, C.POINTER(vp), C.c_int, C.POINTER(vp)]
    L.hec_bootstrapp.argtypes = [vp, vp, C.POINTER(BtpParams), C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
    L.hec_sub_sum.argtypes = [vp, vp, C.c_int]
    L.hec_mod_up.argtypes = [vp, vp, C.POINTER(vp)]
    L.hec_bootstrap_ctos.argtypes = [vp, vp, C.POINTER(BtpParams), C.POINTER(vp), C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_double)]
    L.hec_evaluate_cheby.argtypes = [vp, vp, C.POINTER(C.c_double), C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
    L.hec_eval_relu.argtypes = [vp, vp, C.c_double, C.c_double, C.POINTER(vp)]
    L.hec_eval_relu_many.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_double, C.c_double, C.POINTER(vp)]
    L.hec_mul_relin_new.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.hec_rotate_gal.argtypes = [vp, vp, C.c_uint64, vp]
    L.hec_rotate_new.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    L.hec_rotate_hoisted.argtypes = [vp, vp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.hec_galois_for_rotation.argtypes = [vp, C.c_int]
    L.hec_galois_for_rotation.restype = C.c_uint64
    L.hec_ntt.argtypes = [vp, C.c_int, C.c_int, u64p, u64p, C.c_int]
    L.hec_keyswitch.argtypes = [vp, C.c_int, u64pp, C.c_uint64, u64pp, u64pp]
    L.hec_moddown.argtypes = [vp, C.c_int, u64pp, u64pp, u64pp]
    L.hec_conv_then_pack.argtypes = [vp, vp, C.POINTER(vp), C.c_int, C.c_int, C.c_double, C.POINTER(vp), vp, C.c_int,
                                     C.POINTER(vp)]
    L.hec_conv_bl.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp), vp, C.POINTER(vp)]
    L.hec_ext_ctxt.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.c_int, C.c_double, C.POINTER(vp)]
    L.hec_keep_ctxt.argtypes = [vp, vp, vp, C.c_double, C.POINTER(vp)]
    L.hec_ext_double_ctxt.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int), C.POINTER(vp), C.c_int, C.POINTER(C.c_int), C.POINTER(vp),
                                      C.c_int, C.c_double, C.POINTER(vp)]
    L.hec_conv_bn_relu.argtypes = [vp, vp, vp, C.POINTER(LayerArgs), C.POINTER(vp)]
    L.hec_plan_create.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(vp), vp,
                                  C.c_int, C.POINTER(vp)]
    L.hec_plan_run.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.hec_plan_run_host.argtypes = [vp, u64pp, u64pp, u64pp, u64pp]
    L.hec_plan_submit_host.argtypes = [vp, u64pp, u64pp, u64pp, u64pp, C.POINTER(C.c_int)]
    L.hec_plan_wait.argtypes = [vp, C.c_int]
    L.hec_plan_span_begin.argtypes = [vp]
    L.hec_plan_span_end_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.hec_plan_profile.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    L.hec_plan_kernel_names.argtypes = [vp, C.c_char_p, C.c_int]
    L.hec_plan_is_deferred.argtypes = [vp]
    L.hec_plan_destroy.argtypes = [vp]
    L.hec_plan_destroy.restype = None
    L.hec_plan_cache_size.argtypes = [vp]
    L.hec_float_quotient_threshold.argtypes = [C.c_uint64]
    L.hec_float_quotient_threshold.restype = C.c_uint64
    _lib = L
    return L


def _rows(a):
    """per-limb pointer array for a contiguous [L][N] uint64 array"""
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"] and a.ndim == 2
    arr = (u64p * a.shape[0])()
    for i in range(a.shape[0]):
        arr[i] = a[i].ctypes.data_as(u64p)
    return arr


def _ptrs(addresses):
    arr = (u64p * len(addresses))()
    for i, x in enumerate(addresses):
        arr[i] = C.cast(C.c_void_p(int(x)), u64p)
    return arr


class Plaintext:
    def __init__(self, ctx, h, level=None):
        self.ctx, self.h, self.level = ctx, h, level

    def free(self):
        if self.h:
            self.ctx.L.hec_pt_free(self.ctx.h, self.h)
            self.h = None


class Ciphertext:
    """ckks.Ciphertext on the device.  `owned=False` marks handles that belong to a Plan."""

    def __init__(self, ctx, h, owned=True):
        self.ctx, self.h, self.owned = ctx, h, owned

    @property
    def level(self):
        return self.ctx.L.hec_ct_level(self.h)

    @property
    def scale(self):
        return self.ctx.L.hec_ct_scale(self.h)

    def set_scale(self, s):
        self.ctx.L.hec_ct_set_scale(self.h, s)

    def download(self):
        n = self.level + 1
        c0 = np.empty((n, self.ctx.N), dtype=np.uint64)
        c1 = np.empty((n, self.ctx.N), dtype=np.uint64)
        self.ctx._chk(self.ctx.L.hec_ct_download(self.ctx.h, self.h, _rows(c0), _rows(c1)))
        return c0, c1

    def free(self):
        if self.h and self.owned:
            self.ctx.L.hec_ct_free(self.ctx.h, self.h)
        self.h = None


class Context:
    """One evaluator (ckks.NewEvaluator, main.go:430-441) bound to one GPU."""

    def __init__(self, logN, Q, P, device=0):
        self.L = lib()
        self.logN, self.N, self.Q, self.P = logN, 1 << logN, list(Q), list(P)
        q = np.array(self.Q, dtype=np.uint64)
        p = np.array(self.P if self.P else [0], dtype=np.uint64)
        h = vp()
        rc = self.L.hec_ctx_create(C.byref(h), logN, q.ctypes.data_as(u64p), len(self.Q), p.ctypes.data_as(u64p),
                                   len(self.P), device)
        if rc != 0:
            raise HecError(rc, "hec_ctx_create failed (no CUDA device or unsupported parameters)")
        self.h = h

    def close(self):
        if self.h:
            self.L.hec_ctx_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise HecError(rc, self.L.hec_last_error(self.h).decode())

    def sync(self):
        self._chk(self.L.hec_sync(self.h))

    def timer_start(self):
        self._chk(self.L.hec_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        self._chk(self.L.hec_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def host_register(self, arr):
        """page-lock a numpy array in place (hec_host_register)"""
        self._chk(self.L.hec_host_register(self.h, arr.ctypes.data, arr.nbytes))

    def host_unregister(self, arr):
        self._chk(self.L.hec_host_unregister(self.h, arr.ctypes.data))

    def launch_count(self):
        return int(self.L.hec_launch_count(self.h))

    # ---- uploads ----
    def upload_pt(self, limbs, scale):
        limbs = np.ascontiguousarray(limbs, dtype=np.uint64)
        h = vp()
        self._chk(self.L.hec_pt_upload(self.h, limbs.shape[0] - 1, _rows(limbs), scale, C.byref(h)))
        return Plaintext(self, h, limbs.shape[0] - 1)

    def EncodeCoeffsNTT(self, values, level, scale):
        """NewPlaintext(params, level, scale); encoder.EncodeCoeffs(values, pt); encoder.ToNTT(pt) (conv.go:512-514)"""
        v = np.ascontiguousarray(values, dtype=np.float64)
        h = vp()
        self._chk(self.L.hec_encode_coeffs(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), v.shape[0], level, scale, C.byref(h)))
        return Plaintext(self, h, level)

    def EncodeCoeffsNTTMany(self, values, level, scale):
        """the plaintext loop of prep_Ker (conv.go:510-515): values [count][n] -> count plaintexts"""
        v = np.ascontiguousarray(values, dtype=np.float64)
        hs = (vp * v.shape[0])()
        self._chk(self.L.hec_encode_coeffs_many(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), v.shape[0], v.shape[1], level, scale, hs))
        return [Plaintext(self, vp(h), level) for h in hs]

    def download_pt(self, pt, level=None):
        level = pt.level if level is None else level
        out = np.empty((level + 1, self.N), dtype=np.uint64)
        self._chk(self.L.hec_pt_download(self.h, pt.h, _rows(out)))
        return out

    def upload_ct(self, c0, c1, scale):
        c0, c1 = np.ascontiguousarray(c0, dtype=np.uint64), np.ascontiguousarray(c1, dtype=np.uint64)
        h = vp()
        self._chk(self.L.hec_ct_upload(self.h, c0.shape[0] - 1, _rows(c0), _rows(c1), scale, C.byref(h)))
        return Ciphertext(self, h)

    def upload_swk(self, galEl, swk, max_level):
        """swk: [digits][2][nQ+nP][N] as Lattigo stores it."""
        swk = np.ascontiguousarray(swk, dtype=np.uint64)
        flat = swk.reshape(-1, self.N)
        self._chk(self.L.hec_swk_upload(self.h, galEl, max_level, _rows(flat)))

    def upload_rlk(self, rlk, max_level):
        """rlk: RelinearizationKey.Keys[0], [digits][2][nQ+nP][N] as Lattigo stores it."""
        flat = np.ascontiguousarray(rlk, dtype=np.uint64).reshape(-1, self.N)
        self._chk(self.L.hec_rlk_upload(self.h, max_level, _rows(flat)))

    # ---- evaluator ops (ckks.Evaluator subset) ----
    def Sub(self, a, b, out):
        self._chk(self.L.hec_sub(self.h, a.h, b.h, out.h))

    def DropLevel(self, ct, levels):
        self._chk(self.L.hec_drop_level(self.h, ct.h, levels))

    def MulByPow2(self, ct, pow2):
        self._chk(self.L.hec_mul_by_pow2(self.h, ct.h, pow2))

    def MultByi(self, ct):
        self._chk(self.L.hec_mult_by_i(self.h, ct.h, 0))

    def DivByi(self, ct):
        self._chk(self.L.hec_mult_by_i(self.h, ct.h, 1))

    def Conjugate(self, ct, out):
        self._chk(self.L.hec_conjugate(self.h, ct.h, out.h))

    def AddConst(self, ct, c):
        self._chk(self.L.hec_add_const(self.h, ct.h, c))

    def MultByIntAndAdd(self, ct, c, out):
        self._chk(self.L.hec_mult_by_int_and_add(self.h, ct.h, c, out.h))

    def EvaluatePoly(self, ct, coeffs, target_scale, eval_scale):
        arr = (C.c_double * len(coeffs))(*coeffs)
        h = vp()
        self._chk(self.L.hec_evaluate_poly(self.h, ct.h, arr, len(coeffs), target_scale, eval_scale, C.byref(h)))
        return Ciphertext(self, h)

    def upload_ptdiag(self, log_slots, n1, level, scale, diags):
        """diags: {k: (dQ [(level+1)][N], dP [nP][N])} -> device handle of a PtDiagMatrix"""
        keys = sorted(diags)
        flat = np.ascontiguousarray(np.concatenate([np.concatenate([diags[k][0][:level + 1], diags[k][1]]) for k in keys]), dtype=np.uint64)
        h = vp()
        self._chk(self.L.hec_ptdiag_upload(self.h, log_slots, n1, level, scale, len(keys), (C.c_int * len(keys))(*keys), _rows(flat), C.byref(h)))
        return h

    def free_ptdiag(self, h):
        self.L.hec_ptdiag_free(self.h, h)

    def LinearTransform(self, ct, mat):
        h = vp()
        self._chk(self.L.hec_linear_transform(self.h, ct.h, mat, C.byref(h)))
        return Ciphertext(self, h)

    def CoeffsToSlots(self, ct, mats):
        h0, h1 = vp(), vp()
        self._chk(self.L.hec_coeffs_to_slots(self.h, ct.h, (vp * len(mats))(*mats), len(mats), C.byref(h0), C.byref(h1)))
        return Ciphertext(self, h0), (Ciphertext(self, h1) if h1.value else None)

    def SubSum(self, ct, log_slots):
        self._chk(self.L.hec_sub_sum(self.h, ct.h, log_slots))

    def SlotsToCoeffs(self, ct0, ct1, mats):
        h = vp()
        self._chk(self.L.hec_slots_to_coeffs(self.h, ct0.h, ct1.h if ct1 is not None else None, (vp * len(mats))(*mats), len(mats), C.byref(h)))
        return Ciphertext(self, h)

    def ModUp(self, ct):
        h = vp()
        self._chk(self.L.hec_mod_up(self.h, ct.h, C.byref(h)))
        return Ciphertext(self, h)

    @staticmethod
    def _btp_params(b):
        co, lo, hi = b["cheby"]
        qi = (C.c_uint64 * len(b["sine_qi"]))(*b["sine_qi"])
        ca = (C.c_double * len(co))(*co)
        bp = BtpParams(b["prescale"], b["postscale"], b["sinescale"], b["sqrt2pi"], b["sc_fac"], b["message_ratio"], b["params_scale"],
                       b["sin_type"], b["sin_rescal"], 0, qi, len(b["sine_qi"]), ca, len(co), lo, hi)
        bp._keep = (qi, ca)   # the arrays must outlive the call
        return bp

    def Bootstrapp(self, ct, b, mats_inv, mats_fwd):
        bp = self._btp_params(b)
        h = vp()
        self._chk(self.L.hec_bootstrapp(self.h, ct.h, C.byref(bp), (vp * len(mats_inv))(*mats_inv), len(mats_inv),
                                        (vp * len(mats_fwd))(*mats_fwd), len(mats_fwd), C.byref(h)))
        return Ciphertext(self, h)

    def BootstrappConv_CtoS(self, ct, b, mats):
        """b: dict with the Bootstrapper fields (names of hec_btp_params, plus cheby = (coeffs, a, b)); mats: uploaded pDFTInv"""
        bp = self._btp_params(b)
        h0, h1, k = vp(), vp(), C.c_double()
        self._chk(self.L.hec_bootstrap_ctos(self.h, ct.h, C.byref(bp), (vp * len(mats))(*mats), len(mats), C.byref(h0), C.byref(h1), C.byref(k)))
        return Ciphertext(self, h0), (Ciphertext(self, h1) if h1.value else None), k.value

    def BootstrappConv_StoC(self, ct0, ct1, mats):
        h = vp()
        self._chk(self.L.hec_bootstrap_stoc(self.h, ct0.h, ct1.h if ct1 is not None else None, (vp * len(mats))(*mats), len(mats), C.byref(h)))
        return Ciphertext(self, h)

    def EvaluateCheby(self, ct, coeffs, target_scale, eval_scale):
        arr = (C.c_double * len(coeffs))(*coeffs)
        h = vp()
        self._chk(self.L.hec_evaluate_cheby(self.h, ct.h, arr, len(coeffs), target_scale, eval_scale, C.byref(h)))
        return Ciphertext(self, h)

    def evalReLU(self, ct, alpha, eval_scale):
        h = vp()
        self._chk(self.L.hec_eval_relu(self.h, ct.h, alpha, eval_scale, C.byref(h)))
        return Ciphertext(self, h)

    def evalReLUMany(self, cts, alpha, eval_scale):
        n = len(cts)
        ins = (vp * n)(*[ct.h for ct in cts])
        outs = (vp * n)()
        self._chk(self.L.hec_eval_relu_many(self.h, ins, n, alpha, eval_scale, outs))
        return [Ciphertext(self, vp(outs[i])) for i in range(n)]

    def MulRelinNew(self, a, b):
        h = vp()
        self._chk(self.L.hec_mul_relin_new(self.h, a.h, b.h, C.byref(h)))
        return Ciphertext(self, h)

    def MulNew(self, ct, pt):
        h = vp()
        self._chk(self.L.hec_mul_pt_new(self.h, ct.h, pt.h, C.byref(h)))
        return Ciphertext(self, h)

    def MultByConst(self, ct, c):
        self._chk(self.L.hec_mult_by_const(self.h, ct.h, c))

    def Rescale(self, ct, min_scale):
        self._chk(self.L.hec_rescale(self.h, ct.h, min_scale))

    def SetScale(self, ct, scale):
        self._chk(self.L.hec_set_scale(self.h, ct.h, scale))

    def Add(self, a, b, out):
        self._chk(self.L.hec_add(self.h, a.h, b.h, out.h))

    def AddNew(self, a, b):
        h = vp()
        self._chk(self.L.hec_add_new(self.h, a.h, b.h, C.byref(h)))
        return Ciphertext(self, h)

    def SubNew(self, a, b):
        h = vp()
        self._chk(self.L.hec_sub_new(self.h, a.h, b.h, C.byref(h)))
        return Ciphertext(self, h)

    def AddPt(self, ct, pt):
        self._chk(self.L.hec_add_pt(self.h, ct.h, pt.h))

    def CopyNew(self, ct):
        h = vp()
        self._chk(self.L.hec_ct_copy_new(self.h, ct.h, C.byref(h)))
        return Ciphertext(self, h)

    def RotateGal(self, ct, galEl, out):
        self._chk(self.L.hec_rotate_gal(self.h, ct.h, galEl, out.h))

    def RotateNew(self, ct, k):
        h = vp()
        self._chk(self.L.hec_rotate_new(self.h, ct.h, k, C.byref(h)))
        return Ciphertext(self, h)

    def RotateHoisted(self, ct, rotations):
        n = len(rotations)
        rots = (C.c_int * n)(*rotations)
        outs = (vp * n)()
        self._chk(self.L.hec_rotate_hoisted(self.h, ct.h, rots, n, outs))
        return {r: Ciphertext(self, vp(outs[i])) for i, r in enumerate(rotations)}

    def galois_for_rotation(self, k):
        return int(self.L.hec_galois_for_rotation(self.h, k))

    # ---- ring level ----
    def ntt(self, a, limb, ring=0, inverse=False):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        self._chk(self.L.hec_ntt(self.h, ring, limb, a.ctypes.data_as(u64p), out.ctypes.data_as(u64p), int(inverse)))
        return out

    def keyswitch(self, c1, galEl):
        c1 = np.ascontiguousarray(c1, dtype=np.uint64)
        d0, d1 = np.empty_like(c1), np.empty_like(c1)
        self._chk(self.L.hec_keyswitch(self.h, c1.shape[0] - 1, _rows(c1), galEl, _rows(d0), _rows(d1)))
        return d0, d1

    def moddown(self, accQ, accP):
        accQ, accP = np.ascontiguousarray(accQ, dtype=np.uint64), np.ascontiguousarray(accP, dtype=np.uint64)
        out = np.empty_like(accQ)
        self._chk(self.L.hec_moddown(self.h, accQ.shape[0] - 1, _rows(accQ), _rows(accP), _rows(out)))
        return out

    # ---- conv path ----
    def conv_then_pack(self, ct_in, pt_ker, norm, out_scale, pt_idx, pt_bias=None, flags=CONV_FUSED):
        """conv_then_pack (conv.go:522-546) [+ bias Add, eval.go:258]."""
        B = len(pt_ker)
        ker = (vp * B)(*[(p.h if p is not None else None) for p in pt_ker])
        idx = (vp * len(pt_idx))(*[(p.h if p is not None else None) for p in pt_idx])
        h = vp()
        self._chk(self.L.hec_conv_then_pack(self.h, ct_in.h, ker, B, norm, out_scale, idx,
                                            pt_bias.h if pt_bias is not None else None, flags, C.byref(h)))
        return Ciphertext(self, h)

    def conv_bl(self, ct_in, in_wid, ker_wid, rot_step, pt_taps, pt_bias=None):
        """evalConv_BN_BL_test's timed interval (eval.go:108-131); pt_taps[i][tap]."""
        flat = [p for row in pt_taps for p in row]
        arr = (vp * len(flat))(*[p.h for p in flat])
        h = vp()
        self._chk(self.L.hec_conv_bl(self.h, ct_in.h, in_wid, ker_wid, len(pt_taps), rot_step, arr,
                                     pt_bias.h if pt_bias is not None else None, C.byref(h)))
        return Ciphertext(self, h)

    def ext_ctxt(self, ct, r_idx, min_scale=None):
        """ext_ctxt (conv.go:347-371); r_idx: {rot: Plaintext}; min_scale=None skips the Rescale
        (the two halves of bsgs_ctxt / the first half of ext_double_ctxt)."""
        rots = list(r_idx)
        ra = (C.c_int * len(rots))(*rots)
        pa = (vp * len(rots))(*[r_idx[r].h for r in rots])
        h = vp()
        self._chk(self.L.hec_ext_ctxt(self.h, ct.h, len(rots), ra, pa, int(min_scale is not None),
                                      float(min_scale or 0.0), C.byref(h)))
        return Ciphertext(self, h)

    def keep_ctxt(self, ct, mask, min_scale):
        """keep_ctxt (conv.go:417-431)."""
        h = vp()
        self._chk(self.L.hec_keep_ctxt(self.h, ct.h, mask.h, min_scale, C.byref(h)))
        return Ciphertext(self, h)

    def ext_double_ctxt(self, ct, m_idx, r_idx, min_scale=None):
        """ext_double_ctxt (conv.go:374-414); min_scale=None: bsgs_ctxt (conv.go:303-344, no Rescale)."""
        def arrs(d):
            rots = list(d)
            return len(rots), (C.c_int * len(rots))(*rots), (vp * len(rots))(*[d[r].h for r in rots])
        nm, rm, pm = arrs(m_idx)
        nr, rr, pr = arrs(r_idx)
        h = vp()
        self._chk(self.L.hec_ext_double_ctxt(self.h, ct.h, nm, rm, pm, nr, rr, pr, int(min_scale is not None),
                                             float(min_scale or 0.0), C.byref(h)))
        return Ciphertext(self, h)

    def conv_bn_relu(self, pack_ctx, ct_input, *, pt_ker, pt_bias, norm, out_scale, pt_idx, pow, alpha, iter, btp, ctos_mats,
                     stoc_mats, min_scale, keep_mask=None, r_idx=None, m_idx=None, pt_pre=None, pt_shift2=None, pt_post=None,
                     conv_flags=CONV_FUSED):
        """evalConv_BNRelu_new (eval.go:272-575) as one call.  self: the main evaluator; pack_ctx: the pack evaluator.
        pt_ker / pt_bias / norm: lists of 1 or 2 convolutions; keep_mask: [Plaintext per half] (keep_ctxt), or
        r_idx (and m_idx): [{rot: Plaintext} per half] (ext_ctxt / ext_double_ctxt)."""
        a = LayerArgs()
        keep = []
        a.n_conv = len(pt_ker)
        a.max_ob = len(pt_ker[0])
        for k in range(a.n_conv):
            arr = (vp * a.max_ob)(*[(p.h if p is not None else None) for p in pt_ker[k]])
            keep.append(arr)
            a.pt_ker[k] = C.cast(arr, C.POINTER(C.c_void_p))
            a.pt_bias[k] = pt_bias[k].h if pt_bias[k] is not None else None
            a.norm[k] = norm[k]
        a.out_scale = out_scale
        idx = (vp * len(pt_idx))(*[(p.h if p is not None else None) for p in pt_idx])
        keep.append(idx)
        a.pt_idx = C.cast(idx, C.POINTER(C.c_void_p))
        a.conv_flags = conv_flags
        a.pt_pre = pt_pre.h if pt_pre is not None else None
        a.pt_shift2 = pt_shift2.h if pt_shift2 is not None else None
        a.pt_post = pt_post.h if pt_post is not None else None
        a.pow, a.alpha, a.iter, a.min_scale = pow, alpha, iter, min_scale
        bp = self._btp_params(btp)
        keep.append(bp)
        a.btp = C.pointer(bp)
        cm, sm = (vp * len(ctos_mats))(*ctos_mats), (vp * len(stoc_mats))(*stoc_mats)
        keep += [cm, sm]
        a.ctos_mats, a.n_ctos = C.cast(cm, C.POINTER(C.c_void_p)), len(ctos_mats)
        a.stoc_mats, a.n_stoc = C.cast(sm, C.POINTER(C.c_void_p)), len(stoc_mats)
        if keep_mask is not None:
            a.move_kind = MOVE_KEEP
            for ul, mk in enumerate(keep_mask):
                a.keep_mask[ul] = mk.h if mk is not None else None
        else:
            a.move_kind = MOVE_EXT_DOUBLE if m_idx is not None else MOVE_EXT
            for name, dicts in (("r", r_idx), ("m", m_idx)):
                if dicts is None:
                    continue
                for ul, d in enumerate(dicts):
                    rots = list(d)
                    ra, pa = (C.c_int * len(rots))(*rots), (vp * len(rots))(*[d[r].h for r in rots])
                    keep += [ra, pa]
                    getattr(a, "n_" + name)[ul] = len(rots)
                    getattr(a, "rots_" + name)[ul] = C.cast(ra, C.POINTER(C.c_int))
                    getattr(a, "pts_" + name)[ul] = C.cast(pa, C.POINTER(C.c_void_p))
        h = vp()
        self._chk(self.L.hec_conv_bn_relu(pack_ctx.h, self.h, ct_input.h, C.byref(a), C.byref(h)))
        return Ciphertext(self, h)

    def plan_cache_size(self):
        return int(self.L.hec_plan_cache_size(self.h))

    def plan(self, pt_ker, norm, in_scale, out_scale, pt_idx, pt_bias, batch):
        return Plan(self, pt_ker, norm, in_scale, out_scale, pt_idx, pt_bias, batch)


class Plan:
    """A prepared evalConv_BN over `batch` independent ciphertexts (CUDA graph)."""

    def __init__(self, ctx, pt_ker, norm, in_scale, out_scale, pt_idx, pt_bias, batch):
        self.ctx, self.batch = ctx, batch
        B = len(pt_ker)
        ker = (vp * B)(*[(p.h if p is not None else None) for p in pt_ker])
        idx = (vp * len(pt_idx))(*[(p.h if p is not None else None) for p in pt_idx])
        h = vp()
        ctx._chk(ctx.L.hec_plan_create(ctx.h, ker, B, norm, in_scale, out_scale, idx,
                                       pt_bias.h if pt_bias is not None else None, batch, C.byref(h)))
        self.h = h
        self._keep = (pt_ker, pt_idx, pt_bias)
        self._outs = (vp * batch)()

    def run(self, cts):
        """device-resident run; returns `batch` level-0 ciphertext handles (reused across runs)."""
        ins = (vp * self.batch)(*[c.h for c in cts])
        self.ctx._chk(self.ctx.L.hec_plan_run(self.h, ins, self._outs))
        return [Ciphertext(self.ctx, vp(self._outs[i]), owned=False) for i in range(self.batch)]

    def _host_ptrs(self, in_c0, in_c1, out_c0, out_c1):
        N = self.ctx.N

        def base(a):
            return a.data_ptr() if hasattr(a, "data_ptr") else a.ctypes.data

        b0, b1, o0, o1 = base(in_c0), base(in_c1), base(out_c0), base(out_c1)
        pin0 = _ptrs([b0 + (m * 2 + i) * N * 8 for m in range(self.batch) for i in range(2)])
        pin1 = _ptrs([b1 + (m * 2 + i) * N * 8 for m in range(self.batch) for i in range(2)])
        po0 = _ptrs([o0 + m * N * 8 for m in range(self.batch)])
        po1 = _ptrs([o1 + m * N * 8 for m in range(self.batch)])
        return pin0, pin1, po0, po1

    def run_host(self, in_c0, in_c1, out_c0, out_c1):
        """host-buffer run.  in_c0/in_c1: [batch][2][N] uint64 arrays (pinned for async copies);
        out_c0/out_c1: [batch][N].  Accepts numpy arrays or objects with data_ptr() (torch)."""
        self.ctx._chk(self.ctx.L.hec_plan_run_host(self.h, *self._host_ptrs(in_c0, in_c1, out_c0, out_c1)))

    def submit_host(self, in_c0, in_c1, out_c0, out_c1):
        """pipelined host-buffer run: returns a ticket; buffers must stay valid until wait(ticket)."""
        t = C.c_int()
        self.ctx._chk(self.ctx.L.hec_plan_submit_host(self.h, *self._host_ptrs(in_c0, in_c1, out_c0, out_c1), C.byref(t)))
        return t.value

    def wait(self, ticket):
        self.ctx._chk(self.ctx.L.hec_plan_wait(self.h, ticket))

    def span_begin(self):
        self.ctx._chk(self.ctx.L.hec_plan_span_begin(self.h))

    def span_end_ms(self):
        ms = C.c_float()
        self.ctx._chk(self.ctx.L.hec_plan_span_end_ms(self.h, C.byref(ms)))
        return ms.value

    def profile(self, cts):
        """[(kernel name, ms)] of one run with kernels launched one by one."""
        ins = (vp * self.batch)(*[c.h for c in cts])
        ms = (C.c_float * 256)()
        n = C.c_int()
        self.ctx._chk(self.ctx.L.hec_plan_profile(self.h, ins, ms, 256, C.byref(n)))
        buf = C.create_string_buffer(4096)
        self.ctx._chk(min(0, self.ctx.L.hec_plan_kernel_names(self.h, buf, 4096)))
        names = buf.value.decode().split(",")
        return [(names[i % len(names)], ms[i]) for i in range(n.value)]

    @property
    def deferred(self):
        return bool(self.ctx.L.hec_plan_is_deferred(self.h))

    def destroy(self):
        if self.h:
            for i in range(self.batch):
                if self._outs[i]:
                    self.ctx.L.hec_ct_free(self.ctx.h, self._outs[i])
                    self._outs[i] = None
            self.ctx.L.hec_plan_destroy(self.h)
            self.h = None
