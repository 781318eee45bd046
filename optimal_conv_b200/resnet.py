"""resnet.py -- the reference's encrypted ResNet inference chain over libhec: testResNet_crop_sparse
(test.go:76-366; CLI `resnet k 20 1`, BASELINE.json configs[4]) as a host program above the C ABI.

What runs here is the reference's layer sequence with the reference's shapes, level bookkeeping and data movement:
7 + 5 + 5 evalConv_BNRelu_new layers ("Conv_sparse") joined by two strided ones ("StrConv_sparse"), then the
reduce-mean + fully-connected layer as one evalConv_BN -- every layer ONE hec_conv_bn_relu call (conv on the pack
evaluator, BootstrappConv_CtoS, evalReLU, keep_ctxt / ext_double_ctxt, BootstrappConv_StoC, Rescale on the main one).
The float side of every layer is the reference's own: kernels and batch-norm vectors read from a Resnet_weights
directory (w{n}-conv.csv, w{n}-a.csv, w{n}-b.csv; hostprep.read_txt) or synthesised, reshaped by prep_Ker's code
(hostprep.prep_ker_coeffs) and encoded on the device (hec_encode_coeffs_many); the slot index maps come from the
ported generators of rot_util.go.

What does NOT exist outside the Go host, and is therefore synthetic here: the secret key and everything derived from
it (encryption of the image, rotation / relinearisation keys), the slot encoder (the 0/1 masks of the index maps as
plaintexts) and the bootstrapper's DFT factor matrices and sine polynomial.  They are replaced by seeded uniform
residues with the REAL supports -- the rotations of the real index maps, the diagonal patterns of the real factor
matrices (synth.dft_factor_specs), every modulus of the real chain -- so the device executes exactly the kernels,
launches, key sizes and memory footprint of a real inference, and the result is a ciphertext nobody can decrypt.
`run()` therefore reports time and memory per layer and the digest of the final ciphertext; logits need the Go host's
keys.  compare_logits() is the acceptance check of compare_final.py for result files produced with real keys.
"""
import math
import time

import numpy as np

from . import hostprep as hp
from . import params as PR
from . import synth

N = 1 << PR.LOGN


def layer_specs(depth=20, ker_wid=3, cf100=False):
    """the evalConv_BNRelu_new calls of testResNet_crop_sparse in order (test.go:152-262), then the final evalConv_BN"""
    blocks = {20: (7, 5, 5), 14: (5, 3, 3), 8: (3, 1, 1)}[depth]
    real_batch, norm = [16, 32, 64], [4, 8, 16]
    in_wids = [32, 16, 8]
    raw = [w - ker_wid // 2 for w in in_wids]
    init_pow, mid_pow, final_pow = (5.0, 5.0, {3: 7.0, 5: 6.0}.get(ker_wid, 5.0)) if cf100 else (6.0, 6.0, 6.0)
    L, n, pow_ = [], 0, init_pow
    for i in range(blocks[0]):
        L.append(dict(name="block1.%d" % (i + 1), kind="Conv_sparse", w=n, in_wid=in_wids[0], kp_wid=raw[0], real_ib=3 if i == 0 else real_batch[0],
                      real_ob=real_batch[0], norm=norm[0], log_sparse=2, pow=pow_))
        pow_, n = mid_pow, n + 1
    L.append(dict(name="block1to2", kind="StrConv_sparse", w=n, in_wid=in_wids[0], kp_wid=raw[1], real_ib=real_batch[0], real_ob=real_batch[1],
                  norm=norm[1], log_sparse=1, pow=pow_))
    n += 1
    for i in range(blocks[1]):
        L.append(dict(name="block2.%d" % (i + 1), kind="Conv_sparse", w=n, in_wid=in_wids[1], kp_wid=raw[1], real_ib=real_batch[1],
                      real_ob=real_batch[1], norm=norm[1], log_sparse=3, pow=pow_))
        n += 1
    L.append(dict(name="block2to3", kind="StrConv_sparse", w=n, in_wid=in_wids[1], kp_wid=raw[2], real_ib=real_batch[1], real_ob=real_batch[2],
                  norm=norm[2], log_sparse=2, pow=pow_))
    n += 1
    for i in range(blocks[2]):
        L.append(dict(name="block3.%d" % (i + 1), kind="Conv_sparse", w=n, in_wid=in_wids[2], kp_wid=raw[2], real_ib=real_batch[2],
                      real_ob=real_batch[2], norm=norm[2], log_sparse=4, pow=final_pow if i == blocks[2] - 1 else pow_))
        n += 1
    fc_wid = raw[2] + (1 - raw[2] % 2)
    L.append(dict(name="final_fc", kind="final", in_wid=in_wids[2], kp_wid=raw[2], ker_wid=fc_wid, real_ib=real_batch[2],
                  real_ob=100 if cf100 else 10, norm=norm[2]))
    return L


def synthetic_weights(spec, ker_wid, seed):
    """a layer's float weights with the shapes of its files: (ker_in, bn_a, bn_b)"""
    rng = np.random.default_rng(seed)
    if spec["kind"] == "final":
        k, ob, ib = spec["ker_wid"], spec["real_ob"], spec["real_ib"]
        fc = rng.uniform(-0.1, 0.1, ib * ob)
        ker = np.tile(fc, k * k)                                           # test.go:302-307: the FC kernel at every tap
        return ker, np.full(ob, 1.0 / (spec["kp_wid"] ** 2)), rng.uniform(-0.1, 0.1, ob)
    n = spec["real_ib"] * spec["real_ob"] * ker_wid * ker_wid
    return rng.uniform(-0.05, 0.05, n), rng.uniform(0.1, 0.3, spec["real_ob"]), rng.uniform(-0.1, 0.1, spec["real_ob"])


def load_weights(weight_dir, spec, ker_wid):
    """the same from a Resnet_weights directory (test.go:171-183, 296-313)"""
    if spec["kind"] == "final":
        k, ob, ib = spec["ker_wid"], spec["real_ob"], spec["real_ib"]
        fc = hp.read_txt(weight_dir + "final-fckernel.csv", ib * ob)
        return np.tile(fc, k * k), np.full(ob, 1.0 / (spec["kp_wid"] ** 2)), hp.read_txt(weight_dir + "final-fcbias.csv", ob)
    fk, fa, fb = hp.weight_files(weight_dir, spec["w"])
    return (hp.read_txt(fk, spec["real_ib"] * spec["real_ob"] * ker_wid * ker_wid), hp.read_txt(fa, spec["real_ob"]),
            hp.read_txt(fb, spec["real_ob"]))


def split_by_output_parity(ker_in, bn_a, bn_b, k2, real_ib, real_ob):
    """the two half kernels of "StrConv_sparse" (eval.go:347-366): output channels 2j and 2j+1"""
    K = np.asarray(ker_in).reshape(k2, real_ib, real_ob)
    return ((K[:, :, 0::2].reshape(-1), np.asarray(bn_a)[0::2], np.asarray(bn_b)[0::2]),
            (K[:, :, 1::2].reshape(-1), np.asarray(bn_a)[1::2], np.asarray(bn_b)[1::2]))


class _Tile:
    """a few seeded limbs per modulus, handed out cyclically: synthetic residues without generating gigabytes"""

    def __init__(self, moduli, n=4):
        self.lim = {q: [synth.uniform_mod(811 + 13 * i + j, N, q) for j in range(n)] for i, q in enumerate(moduli)}
        self.k = 0

    def limbs(self, moduli):
        self.k += 1
        return np.stack([self.lim[q][(self.k + i) % len(self.lim[q])] for i, q in enumerate(moduli)])


class Resnet:
    """contexts, keys, bootstrappers and masks of one network (newContext, main.go:43-330), then run() per image"""

    def __init__(self, hec, depth=20, ker_wid=3, device=0, weight_dir=None, cf100=False, log=None):
        self.hec, self.ker_wid, self.weight_dir = hec, ker_wid, weight_dir
        self.specs = layer_specs(depth, ker_wid, cf100)
        self.log = log or (lambda *a: None)
        Q, P = PR.Q_SET6, PR.P_ALL
        self.Q, self.P = Q, P
        self.pack = hec.Context(PR.LOGN, Q[:2], PR.P_PACK, device=device)   # cont.pack_evaluator (main.go:446-454)
        self.main = hec.Context(PR.LOGN, Q, P, device=device)               # cont.evaluator
        self.tile = _Tile(list(Q) + list(P))
        t0 = time.perf_counter()
        # ---- pack evaluator: monomials and the pack keys (gen_idxNlogs, conv.go:241-261)
        mono = np.zeros((PR.LOGN, N), dtype=np.uint64)
        for i in range(PR.LOGN):
            m = np.zeros(N, dtype=np.uint64)
            m[1 << i] = 1
            mono[i] = self.pack.ntt(m, 0)
        self.idx = [self.pack.upload_pt(mono[i:i + 1], 1.0) for i in range(PR.LOGN)]
        pm = list(Q[:2]) + list(PR.P_PACK)                                    # the level-0/1 slice of the alpha = 1 pack keys
        pk = np.stack([np.stack([self.tile.limbs(pm) for _ in range(2)]) for _ in range(2)])
        for j in range(PR.LOGN):
            self.pack.upload_swk((1 << (j + 1)) + 1, pk, 0)
        # ---- main evaluator: one synthetic key buffer stands in for every rotation key (sizes and levels are the real ones)
        beta = (len(Q) + len(P) - 1) // len(P)
        self.keybuf = np.stack([np.stack([self.tile.limbs(list(Q) + list(P)) for _ in range(2)]) for _ in range(beta)])
        self.have = {}
        self.main.upload_rlk(self.keybuf, 27)
        self.main.upload_swk(2 * N - 1, self.keybuf, 27)                     # conjugation
        # ---- per packing density: bootstrapper (factor matrices with the real diagonal supports) and masks
        self.btp, self.masks = {}, {}
        for ls in sorted({s["log_sparse"] for s in self.specs if "log_sparse" in s}):
            self.btp[ls] = self._bootstrapper(15 - ls)
        for s in self.specs:
            if s["kind"] == "Conv_sparse":
                self.masks[s["name"]] = self._mask(4)                        # ext_idx[in_wid][0] (gen_keep_vec_sparse), level of the ReLU output
            elif s["kind"] == "StrConv_sparse":
                m_idx, r_idx = hp.gen_comprs_sparse(N // 2, s["in_wid"], s["kp_wid"], s["log_sparse"], 0, 0)
                sq = float(np.sqrt(np.float64(Q[4])))
                for r in list(m_idx) + list(r_idx):
                    self._key(r, 4)
                self.masks[s["name"]] = ({r: self._mask(4, sq) for r in m_idx}, {r: self._mask(4, sq) for r in r_idx})
        self.main.sync()
        self.setup_s = time.perf_counter() - t0

    # synthetic stand-ins ------------------------------------------------------------------------------------------
    def _key(self, rot, max_level):
        g = self.main.galois_for_rotation(rot)
        if self.have.get(g, -1) < max_level:
            self.main.upload_swk(g, self.keybuf, max_level)
            self.have[g] = max_level

    def _mask(self, level, scale=None):
        return self.main.upload_pt(self.tile.limbs(self.Q[:level + 1]), float(self.Q[level]) if scale is None else scale)

    def _bootstrapper(self, log_slots):
        b = dict(synth.CTOS_FIELDS, sine_qi=self.Q[16:24])
        rng = np.random.default_rng(77)
        b["cheby"] = ([float(x) for x in rng.uniform(-1, 1, 64)], -25.0 / 4, 25.0 / 4)
        mats = {}
        for name, depth, top in (("ctos", 4, 27), ("stoc", 2, 3)):
            hs = []
            for n1, diags, lv in synth.dft_factor_specs(log_slots, depth, top):
                for d in diags:
                    if d % n1:
                        self._key(d % n1, lv)
                    if d // n1:
                        self._key((d // n1) * n1, lv)
                D = {d: (self.tile.limbs(self.Q[:lv + 1]), self.tile.limbs(self.P)) for d in diags}
                hs.append(self.main.upload_ptdiag(log_slots, n1, lv, float(self.Q[lv]), D))
            mats[name] = hs
        for i in range(log_slots, PR.LOGN - 1):                               # Bootstrapper.subSum and the repacking rotation
            self._key(1 << i, 27)
        return b, mats

    # one image ----------------------------------------------------------------------------------------------------
    def _weights(self, s, seed):
        if self.weight_dir:
            return load_weights(self.weight_dir, s, self.ker_wid)
        return synthetic_weights(s, self.ker_wid, seed)

    def _encode_conv(self, ker_in, bn_a, bn_b, in_wid, k, real_ib, real_ob, norm, out_scale):
        """prep_Ker + the bias plaintext of evalConv_BN (eval.go:231-243): float reshaping on the host, EncodeCoeffs + ToNTT on the device"""
        max_bat = N // (in_wid * in_wid)
        act = [i for i in range(max_bat) if i % norm == 0]
        coeffs = hp.prep_ker_coeffs(N, ker_in, bn_a, in_wid, k, real_ib, real_ob, norm, rows=act)
        pts = self.pack.EncodeCoeffsNTTMany(np.stack(coeffs), PR.ECD_LV, PR.SCALE)
        ker = [None] * max_bat
        for i, p in zip(act, pts):
            ker[i] = p
        bias = self.pack.EncodeCoeffsNTT(hp.bias_coeffs(N, bn_b, in_wid, norm), 0, out_scale)
        return ker, bias

    def run(self, image, seed=1, sync_layers=True):
        """image: in_wid^2 * 3 floats (test_image_{i}.csv).  Returns (final ciphertext handle, per-layer records).
        sync_layers=False: no synchronisation between layers -- a layer call returns once its launches are enqueued, so the
        float preparation of the next layer (prep_Ker's reshaping, on the host) runs while the device works; the records'
        eval_ms is then only the time the call took to return.  (Measured: 0.99 s per image against 0.70 s synchronised --
        same result, but with several layers in flight the buffers no longer land on the addresses of the previous image,
        so the job tables of every operation miss the context's cache and are staged again.)"""
        hec, k = self.hec, self.ker_wid
        s0 = self.specs[0]
        packed = hp.pack_image_sparse(image, s0["in_wid"], s0["kp_wid"], N // s0["in_wid"] ** 2, s0["norm"], N)
        # the client's job (EncodeCoeffs + Encrypt, test.go:152-155): a fresh encryption is two uniform-looking polynomials
        del packed
        ct = self.pack.upload_ct(self.tile.limbs(self.Q[:2]), self.tile.limbs(self.Q[:2]), PR.SCALE)
        rec = []
        for li, s in enumerate(self.specs):
            t0 = time.perf_counter()
            ker_in, bn_a, bn_b = self._weights(s, 1000 * seed + li)
            max_bat = N // (s["in_wid"] ** 2)
            flags = hec.CONV_FUSED
            if s["kind"] == "final":
                ker, bias = self._encode_conv(ker_in, bn_a, bn_b, s["in_wid"], s["ker_wid"], s["real_ib"], s["real_ob"], s["norm"], PR.SCALE)
                t1 = time.perf_counter()
                out = self.pack.conv_then_pack(ct, ker, s["norm"], PR.SCALE, self.idx, bias, flags)
                if sync_layers:
                    self.pack.sync()
            else:
                out_scale = float(2.0 ** round(math.log2(float(self.Q[0])) - (s["pow"] + 8)))   # eval.go:369
                b, mats = self.btp[s["log_sparse"]]
                kw = dict(out_scale=out_scale, pt_idx=self.idx, pow=s["pow"], alpha=0.0, iter=2, btp=b, ctos_mats=mats["ctos"],
                          stoc_mats=mats["stoc"], min_scale=PR.SCALE, conv_flags=flags)
                if s["kind"] == "Conv_sparse":
                    ker, bias = self._encode_conv(ker_in, bn_a, bn_b, s["in_wid"], k, s["real_ib"], s["real_ob"], s["norm"], out_scale)
                    t1 = time.perf_counter()
                    out = self.main.conv_bn_relu(self.pack, ct, pt_ker=[ker], pt_bias=[bias], norm=[s["norm"]],
                                                 keep_mask=[self.masks[s["name"]], None], **kw)
                else:
                    halves = split_by_output_parity(ker_in, bn_a, bn_b, k * k, s["real_ib"], s["real_ob"])
                    enc = [self._encode_conv(kk, a, bb, s["in_wid"], k, s["real_ib"], s["real_ob"] // 2, s["norm"] // 2, out_scale)
                           for kk, a, bb in halves]
                    xi = np.zeros(N)
                    xi[s["norm"] // 4] = 1.0                                                     # eval.go:372-376
                    shift = self.pack.EncodeCoeffsNTT(xi, 0, 1.0)
                    xi[:] = 0.0
                    if (s["in_wid"] - k // 2) % 2:                                               # eval.go:383-390
                        xi[0] = 1.0
                    else:
                        xi[N - (N // s["in_wid"] ** 2) * (s["in_wid"] + 1)] = -1.0
                    post = self.pack.EncodeCoeffsNTT(xi, 0, 1.0)
                    m_idx, r_idx = self.masks[s["name"]]
                    t1 = time.perf_counter()
                    out = self.main.conv_bn_relu(self.pack, ct, pt_ker=[e[0] for e in enc], pt_bias=[e[1] for e in enc],
                                                 norm=[s["norm"] // 2] * 2, m_idx=[m_idx, {}], r_idx=[r_idx, {}], pt_shift2=shift, pt_post=post, **kw)
                    ker = [p for e in enc for p in e[0]]
                    for p in (shift, post, enc[0][1], enc[1][1]):
                        p.free()
                    bias = None
                if sync_layers:
                    self.main.sync()
            t2 = time.perf_counter()
            for p in ker:
                if p is not None:
                    p.free()
            if bias is not None:
                bias.free()
            ct.free()
            ct = out
            rec.append({"layer": s["name"], "kind": s["kind"], "prep_ms": 1e3 * (t1 - t0), "eval_ms": 1e3 * (t2 - t1),
                        "level_out": out.level, "conv": "fused" if flags == hec.CONV_FUSED else "op-level"})
            self.log("%-10s %-14s prep %7.1f ms  eval %7.1f ms  -> level %d" % (s["name"], s["kind"], rec[-1]["prep_ms"], rec[-1]["eval_ms"], out.level))
        return ct, rec

    def close(self):
        self.main.close()
        self.pack.close()


def compare_logits(enc_dir, plain_dir, ker, max_num_samples, num_classes=10, tol=None):
    """compare_final.py:8-64 (compare_results): the encrypted run's class_result_ker{k}_{i}.csv files against the plain
    model's plain_prediction_{n}.csv ([n][classes]) and test_labels_{n}.csv, over the images for which a result file
    exists.  Returns the reference's three figures -- plain precision, encrypted precision, plain-vs-encrypted
    accordance (arg-max agreement, the acceptance criterion) -- the number compared, and the largest absolute logit
    difference; with `tol`, raises if that difference exceeds it."""
    import os
    plain = hp.read_txt(os.path.join(plain_dir, "plain_prediction_%d.csv" % max_num_samples)).reshape(max_num_samples, num_classes)
    labels = hp.read_txt(os.path.join(plain_dir, "test_labels_%d.csv" % max_num_samples)).reshape(max_num_samples)
    acc = true_acc = pl_true_acc = total = 0
    worst = 0.0
    for i in range(max_num_samples):
        f = os.path.join(enc_dir, "class_result_ker%d_%d.csv" % (ker, i))
        if not os.path.exists(f):
            continue
        total += 1
        res = hp.read_txt(f)[:num_classes]
        acc += int(np.argmax(res) == np.argmax(plain[i]))
        true_acc += int(np.argmax(res) == labels[i])
        pl_true_acc += int(np.argmax(plain[i]) == labels[i])
        worst = max(worst, float(np.max(np.abs(res - plain[i]))))
    if tol is not None and total and worst > tol:
        raise AssertionError("encrypted and plain logits differ by %.3g > %.3g" % (worst, tol))
    return {"plain_precision": pl_true_acc, "enc_precision": true_acc, "accordance": acc, "compared": total, "max_abs_diff": worst}
