"""Host-side float preparation around the conv path, as the reference's Go host does it:
kernel reshaping and coefficient layout (prep_Ker up to EncodeCoeffs, conv.go:184-237,487-516), input packing and result
extraction (main.go:1007-1070; test.go:134-150), the slot index maps of the between-layer helpers (rot_util.go), and the
whitespace-separated float files the programs exchange (main.go:971-1005).  Plain numpy: no device, no checker.

reshape_ker, encode_ker_final, prep_input (non-transposed) and post_process are pinned against the reference's compiled
routines (tests/golden/make_ref_eval_vectors.py::hostprep_case).  Each function cites the reference lines it follows.
"""
import numpy as np


# ---- files (main.go:971-1005) ------------------------------------------------------------------------------------
def read_txt(path, size=0):
    """readTxt: every whitespace-separated token is a float64; a non-zero `size` must match ("input size inconsistent!")."""
    with open(path) as f:
        vals = [float(t) for t in f.read().split()]
    if size != 0 and len(vals) != size:
        raise ValueError("input size inconsistent!")
    return np.array(vals, dtype=np.float64)


def write_txt(path, values):
    """writeTxt: one value per line, strconv.FormatFloat(v, 'e', -1, 64) -- the shortest representation that round-trips,
    in exponent form with at least two exponent digits."""
    def fmt(v):
        r = repr(float(v))
        if r in ("inf", "-inf", "nan"):
            return {"inf": "+Inf", "-inf": "-Inf", "nan": "NaN"}[r]
        mant, _, exp = r.partition("e")
        sign = "-" if mant.startswith("-") else ""
        mant = mant.lstrip("-")
        e = int(exp) if exp else 0
        ip, _, fp = mant.partition(".")
        digits = (ip + fp).lstrip("0")
        if not digits:
            return sign + "0e+00"
        # position of the decimal point relative to the first significant digit
        lead = len(ip.lstrip("0")) if ip.strip("0") else -(len(fp) - len(fp.lstrip("0")))
        e10 = e + lead - 1
        digits = digits.rstrip("0") or "0"
        m = digits[0] + ("." + digits[1:] if len(digits) > 1 else "")
        return "%s%se%s%02d" % (sign, m, "-" if e10 < 0 else "+", abs(e10))
    with open(path, "w") as f:
        for v in values:
            f.write(fmt(v) + "\n")


def weight_files(weight_dir, layer):
    """the per-layer files of a Resnet_weights directory (test.go:171-183): w{n}-conv.csv, w{n}-a.csv, w{n}-b.csv"""
    return tuple("%sw%d-%s.csv" % (weight_dir, layer, k) for k in ("conv", "a", "b"))


# ---- slot index maps of the between-layer helpers (rot_util.go) --------------------------------------------------------
def reverse_bits(num, bitwid):
    """reverseBits (conv.go:13-24)"""
    r = 0
    for i in range(bitwid):
        r |= ((num >> i) & 1) << (bitwid - 1 - i)
    return r


def gen_keep_vec(vec_size, in_wid, kp_wid, ul):
    """gen_keep_vec (rot_util.go:141-174): 0/1 slot mask keeping the kp_wid x kp_wid valid outputs of half `ul`"""
    logN = (2 * vec_size - 1).bit_length()
    idx = np.zeros(vec_size, dtype=np.int64)
    batch = 2 * vec_size // (in_wid * in_wid)
    if kp_wid < in_wid // 2:
        raise ValueError("keep width too small. less than in_wid/2")
    rows = range(in_wid // 2) if ul == 0 else range(kp_wid - in_wid // 2)
    for i in rows:
        for j in range(kp_wid):
            for b in range(batch):
                idx[reverse_bits(in_wid * batch * i + batch * j + b, logN - 1)] = 1
    return idx


def gen_keep_vec_sparse(vec_size, in_wid, kp_wid, log_sparse):
    """gen_keep_vec_sparse (rot_util.go:179-221): both halves in one ciphertext under sparse packing"""
    logN = (2 * vec_size - 1).bit_length()
    idx = np.zeros(vec_size, dtype=np.int64)
    batch = 2 * vec_size // (in_wid * in_wid)
    sp = 1 << log_sparse
    if sp == 1:
        raise ValueError("We do not support full packing in gen_keep_vec_sparse")
    if kp_wid < in_wid // 2:
        raise ValueError("keep width too small. less than in_wid/2")
    for i in range(in_wid // 2):
        for j in range(kp_wid):
            for b in range(batch // sp):
                idx[reverse_bits(in_wid * batch * i + batch * j + b * sp, logN - 1)] = 1
    for i in range(kp_wid - in_wid // 2):
        for j in range(kp_wid):
            for b in range(batch // sp):
                idx[reverse_bits(in_wid * batch * i + batch * j + b * sp, logN - 1) + vec_size // sp] = 1
    post = 2 * vec_size // sp
    for j in range(1, sp // 2):
        idx[post * j:post * (j + 1)] = idx[:post]
    return idx


def gen_comprs_sparse(vec_size, in_wid, kp_wid, log_sparse, ul, pos):
    """gen_comprs_sparse (rot_util.go:557-722): the two rotate-and-mask stages (m_idx, r_idx: {rotation: 0/1 slot mask})
    that compress a strided convolution's output, for sparse (log_sparse != 0) and full packing"""
    m_idx, r_idx = {}, {}
    batch = 2 * vec_size // (in_wid * in_wid * (1 << log_sparse))
    min_wid = in_wid // 2
    if in_wid % 2:
        raise ValueError("input wid not divisible by 2")
    lw = (in_wid - 1).bit_length()
    rb = reverse_bits
    if log_sparse != 0:
        if pos != 0:
            raise ValueError("No pos != 0 cases for log_sparse != 0")
        rep = 1 << (log_sparse - 1)
        for j in range(min_wid):
            tmp = np.zeros(vec_size, dtype=np.int64)
            for b in range(batch):
                for i in range(min_wid // 2):
                    for k in range(2):
                        if rb(j, lw - 1) < kp_wid and rb(i, lw - 2) + k * min_wid // 2 < kp_wid:
                            tmp[k * in_wid * min_wid * batch + in_wid * in_wid * b // 2 + in_wid * j // 2 + i] = 1
            blk = vec_size // rep
            for k in range(1, rep):
                tmp[k * blk:(k + 1) * blk] = tmp[:blk]
            m_idx[j * min_wid // 2] = tmp
        for b in range(batch):
            tmp = np.zeros(vec_size, dtype=np.int64)
            for j in range(min_wid):
                for i in range(min_wid // 2):
                    for k in range(2):
                        tmp[k * in_wid * min_wid * batch + b * in_wid * in_wid // 2 + j * min_wid // 2 + i] = 1
            blk = vec_size // rep
            for k in range(1, rep):
                tmp[k * blk:(k + 1) * blk] = tmp[:blk]
            r_idx[3 * b * min_wid * min_wid // 2] = tmp
        return m_idx, r_idx
    grp = 8 if batch > 8 * min_wid else (4 if batch > 4 * min_wid else 1)

    def keep(i, j):
        if rb(j, lw - 1) >= kp_wid:
            return False
        return rb(i, lw - 2) < kp_wid if ul == 0 else rb(i, lw - 2) + min_wid // 2 < kp_wid

    for j in range(min_wid):
        for bk in range(grp):
            tmp = np.zeros(vec_size, dtype=np.int64)
            for b in range(batch // grp):
                for i in range(min_wid // 2):
                    if keep(i, j):
                        tmp[grp * in_wid * min_wid * b + bk * min_wid * in_wid + min_wid * j + i] = 1
            m_idx[j * min_wid // 2 + (grp - 1) * bk * min_wid * min_wid // 2] = tmp
    for b in range(batch // grp):
        tmp = np.zeros(vec_size, dtype=np.int64)
        for bk in range(grp):
            for j in range(min_wid):
                for i in range(min_wid // 2):
                    tmp[grp * b * in_wid * min_wid + bk * min_wid * min_wid // 2 + j * min_wid // 2 + i] = 1
        r_idx[3 * b * grp * min_wid * min_wid // 2 - rb(pos, 2) * batch * min_wid * min_wid // 2] = tmp
    return m_idx, r_idx


def gen_comprs_fast(vec_size, in_wid, kp_wid, pos, ul):
    """gen_comprs_fast (rot_util.go:498-548): the two stages of the fast packing of a strided convolution's output under
    full packing (m_idx then r_idx, ext_double_ctxt)"""
    m_idx, r_idx = {}, {}
    batch = 2 * vec_size // (in_wid * in_wid)
    if kp_wid < in_wid // 2:
        raise ValueError("keep width too small. less than in_wid/2")
    pos = reverse_bits(pos, 2)
    min_wid = in_wid // 4
    if in_wid % 4:
        raise ValueError("input wid not divisible by 4")
    lw = (in_wid - 1).bit_length()
    for j in range(2 * min_wid):
        tmp = np.zeros(vec_size, dtype=np.int64)
        for b in range(batch):
            for i in range(min_wid):
                ok = reverse_bits(in_wid // 2 + j, lw) < kp_wid
                if ul == 1:
                    ok = ok and reverse_bits(min_wid + i, lw - 1) < kp_wid - in_wid // 2
                if ok:
                    tmp[2 * min_wid * in_wid * b + 2 * min_wid * j + i + in_wid * min_wid + min_wid] = 1
        m_idx[j * min_wid - 2 * min_wid * min_wid + min_wid] = tmp
    for b in range(batch):
        tmp = np.zeros(vec_size, dtype=np.int64)
        for j in range(2 * min_wid):
            for i in range(min_wid):
                tmp[2 * min_wid * in_wid * b + 3 * in_wid // 2 * min_wid + j * min_wid + i] = 1
        r_idx[3 * b * min_wid * in_wid // 2 - pos * min_wid * in_wid // 2 * batch + 3 * min_wid * in_wid // 2] = tmp
    return m_idx, r_idx


# ---- input / output layout --------------------------------------------------------------------------------------------
def pack_image_sparse(image, in_wid, raw_in_wid, max_batch, norm, N, channels=3):
    """the sparse packing of the network input (test.go:134-146): coefficient i*in_wid*max_batch + j*max_batch + b*norm"""
    out = np.zeros(N)
    k = 0
    for i in range(in_wid):
        for j in range(in_wid):
            for b in range(channels):
                if i < raw_in_wid and j < raw_in_wid:
                    out[i * in_wid * max_batch + j * max_batch + b * norm] = image[k]
                k += 1
    return out


def mat_one_norm(vec, batch, norm, sj, sk):
    """prt_mat_one_norm (main.go:920-939): the (sj, sk) position of every norm-th batch -- where the final reduce-mean +
    FC convolution leaves the logits"""
    mat_size = len(vec) // batch
    j = k = 1
    out = None
    for i in range(0, len(vec), batch):
        if j == sj and k == sk:
            out = np.array([vec[i + norm * idx] for idx in range(batch // norm)])
        k += 1
        if k * k > mat_size:
            k = 1
            j += 1
    return out


def prep_input(raw, raw_w, w, N, norm, trans=False):
    """prep_Input (main.go:1007-1041); trans: the (2i+1, 2j+1) placement of the transposed convolution."""
    B = N // (w * w)
    out = np.zeros(N)
    k = 0
    lim = w // 2 if trans else w
    for i in range(lim):
        for j in range(lim):
            for b in range(B // norm):
                if i < raw_w and j < raw_w:
                    if trans:
                        out[(2 * i + 1) * w * B + (2 * j + 1) * B + b * norm] = raw[k]
                    else:
                        out[i * w * B + j * B + b * norm] = raw[k]
                    k += 1
    return out


def reshape_ker(ker_in, k_sz, out_batch):
    """reshape_ker, trans=false (conv.go:184-202): out[i][j*k_sz + k] = ker_in[i + j*out_batch + k*out_batch*in_batch]"""
    ker_in = np.asarray(ker_in, dtype=np.float64)
    in_batch = len(ker_in) // (k_sz * out_batch)
    return np.ascontiguousarray(ker_in.reshape(k_sz, in_batch, out_batch).transpose(2, 1, 0)).reshape(out_batch, in_batch * k_sz)


def encode_ker_final(ker_in, pos, i, w, B, k):
    """encode_ker_final (conv.go:206-237): output channel i's kernel as the coefficient vector of its plaintext --
    tap t of input channel j at coefficient (w*(t/k) + t%k)*B + j, taps and channels reversed, then the whole vector
    rotated by `adj` positions with the wrapped part negated (multiplication by a monomial modulo X^N + 1)"""
    vec = w * w * B
    k_sz = k * k
    bias = pos * k_sz * B
    row = np.asarray(ker_in[i][bias:bias + B * k_sz], dtype=np.float64).reshape(B, k_sz)[::-1, ::-1]   # [j][t]
    t = np.arange(k_sz)
    out = np.zeros((w * w, B))
    out[w * (t // k) + t % k, :] = row.T
    out = out.reshape(vec)
    adj = (B - 1) + B * (w + 1) * (k - 1) // 2
    return np.concatenate([out[adj:vec - adj], out[vec - adj:], -out[:adj]])


def prep_ker_coeffs(N, ker_in, bn_a, w, k, real_ib, real_ob, norm, rows=None):
    """prep_Ker up to (not including) EncodeCoeffs (conv.go:487-516): the float coefficient vectors of the kernel
    plaintexts, for all max_bat output channels or for `rows` only (conv_then_pack reads every norm-th one)."""
    max_bat = N // (w * w)
    k_sz = k * k
    ker_rs = reshape_ker(ker_in, k_sz, real_ob) * np.asarray(bn_a, dtype=np.float64)[:, None]
    max_ker = np.zeros((max_bat, max_bat, k_sz))
    max_ker[:norm * real_ob:norm, :norm * real_ib:norm, :] = ker_rs.reshape(real_ob, real_ib, k_sz)
    max_ker = max_ker.reshape(max_bat, max_bat * k_sz)
    return [encode_ker_final(max_ker, 0, i, w, max_bat, k) for i in (range(max_bat) if rows is None else rows)]


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
