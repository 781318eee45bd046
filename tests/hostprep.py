"""Host-side plaintext preparation of the reference, restated in numpy for the tests.

These routines stay in Go in a real deployment; the tests need them
to drive the evaluator with *meaningful* operands for the decrypt-and-compare check.
Each function cites the reference lines it follows.
"""
import numpy as np


def prep_input(raw, raw_w, w, N, norm):
    """prep_Input, non-transposed branch (main.go:1024-1031)."""
    B = N // (w * w)
    out = np.zeros(N)
    k = 0
    for i in range(w):
        for j in range(w):
            for b in range(B // norm):
                if i < raw_w and j < raw_w:
                    out[i * w * B + j * B + b * norm] = raw[k]
                    k += 1
    return out


def reshape_ker(ker_in, k_sz, out_batch):
    """reshape_ker, trans=false (conv.go:184-202)."""
    in_batch = len(ker_in) // (k_sz * out_batch)
    out = np.zeros((out_batch, k_sz * in_batch))
    for i in range(out_batch):
        for j in range(in_batch):
            for k in range(k_sz):
                out[i, j * k_sz + k] = ker_in[i + j * out_batch + k * out_batch * in_batch]
    return out


def encode_ker_final(ker_in, pos, i, w, B, k):
    """encode_ker_final (conv.go:206-237)."""
    vec = w * w * B
    out = np.zeros(vec)
    k_sz = k * k
    bias = pos * k_sz * B
    for j in range(B):
        for t in range(k_sz):
            out[(w * (t // k) + t % k) * B + j] = ker_in[i][(B - 1 - j) * k_sz + (k_sz - 1 - t) + bias]
    adj = (B - 1) + B * (w + 1) * (k - 1) // 2
    tmp = out[vec - adj:].copy()
    head = out[:adj].copy()
    body = out[adj:vec - adj].copy()
    res = np.empty(vec)
    res[:vec - 2 * adj] = body
    res[vec - 2 * adj:vec - adj] = tmp
    res[vec - adj:] = -head
    return res


def prep_ker_coeffs(N, ker_in, bn_a, w, k, real_ib, real_ob, norm):
    """prep_Ker up to (not including) EncodeCoeffs (conv.go:487-516): float coefficient
    vectors for the max_bat kernel plaintexts."""
    max_bat = N // (w * w)
    k_sz = k * k
    ker_rs = reshape_ker(ker_in, k_sz, real_ob)
    ker_rs = ker_rs * np.asarray(bn_a)[:, None]
    max_ker = np.zeros((max_bat, max_bat * k_sz))
    for i in range(real_ob):
        for j in range(real_ib):
            max_ker[norm * i, norm * j * k_sz:norm * j * k_sz + k_sz] = ker_rs[i, j * k_sz:(j + 1) * k_sz]
    return [encode_ker_final(max_ker, 0, i, w, max_bat, k) for i in range(max_bat)]


def bias_coeffs(N, bn_b, w, norm):
    """b_coeffs of evalConv_BN (eval.go:233-238)."""
    max_batch = N // (w * w)
    b = np.zeros(N)
    for i, v in enumerate(bn_b):
        for j in range(w * w):
            b[norm * i + j * max_batch] = v
    return b


def encode_coeffs(vals, scale, moduli):
    """EncodeCoeffs (L:ckks/encoder.go:655-664, utils.go:60-110): coefficient j <-
    floor(|v|*scale + 0.5) with sign, reduced per limb.  Not NTT'd."""
    out = np.empty((len(moduli), len(vals)), dtype=np.uint64)
    ints = [int(np.floor(abs(float(v)) * scale + 0.5)) * (1 if v >= 0 else -1) for v in vals]
    for i, q in enumerate(moduli):
        out[i] = np.array([x % q for x in ints], dtype=np.uint64)
    return out


def decode_coeffs(res, scale, moduli):
    """DecodeCoeffs at the level of `res` ([L][N] coefficient residues): centred CRT lift."""
    L = res.shape[0]
    if L == 1:
        q = moduli[0]
        v = res[0].astype(object)
        return np.array([float(x - q if x > q // 2 else x) for x in v]) / scale
    Qp = 1
    for q in moduli[:L]:
        Qp *= q
    acc = [0] * res.shape[1]
    for i in range(L):
        q = moduli[i]
        Qi = Qp // q
        inv = pow(Qi % q, -1, q)
        col = res[i].astype(object)
        for j in range(len(acc)):
            acc[j] = (acc[j] + int(col[j]) * inv % q * Qi) % Qp
    return np.array([float(x - Qp if x > Qp // 2 else x) for x in acc]) / scale


def post_process(cfs, raw_w, w):
    """post_process (main.go:1057-1070)."""
    B = len(cfs) // (w * w)
    out = np.zeros(raw_w * raw_w * B)
    for i in range(raw_w):
        for j in range(raw_w):
            for b in range(B):
                out[i * raw_w * B + B * j + b] = cfs[i * w * B + B * j + b]
    return out


def plain_conv_same(raw, ker_in, bn_a, bn_b, raw_w, k, B):
    """'SAME' cross-correlation + BN, HWIO kernel, HWC input/output order (the content of
    the reference's absent test_conv*_out_*.csv)."""
    x = np.asarray(raw).reshape(raw_w, raw_w, B)
    K = np.asarray(ker_in).reshape(k, k, B, B)
    pad = k // 2
    xp = np.zeros((raw_w + 2 * pad, raw_w + 2 * pad, B))
    xp[pad:pad + raw_w, pad:pad + raw_w] = x
    out = np.zeros((raw_w, raw_w, B))
    for r in range(k):
        for c in range(k):
            out += np.einsum("ijb,bo->ijo", xp[r:r + raw_w, c:c + raw_w], K[r, c])
    out = out * np.asarray(bn_a)[None, None, :] + np.asarray(bn_b)[None, None, :]
    return out.reshape(-1)
