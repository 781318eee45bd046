// hec.cu -- libhec.so: context, generic evaluator ops, C ABI (see include/hec.h).
// The fused conv path lives in hec_conv.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include "hec_host.cuh"

using namespace hec;

void hec_plan_cache_clear(hec_ctx *c);                    // hec_conv.cu
void hec_plan_cache_evict(hec_ctx *c, uint64_t serial);   // hec_conv.cu

// =========================================================================================
// ring constants
// =========================================================================================
namespace hec {

// smallest primitive root >= 3 (search starts at 2 and increments before testing;
// L:ring/utils.go:69-90)
static u64 primitive_root(u64 q) {
    std::vector<u64> fac;
    u64 m = q - 1;
    for (u64 p = 2; p * p <= m; p += (p == 2 ? 1 : 2))
        if (m % p == 0) { fac.push_back(p); while (m % p == 0) m /= p; }
    if (m > 1) fac.push_back(m);
    for (u64 g = 3;; g++) {
        bool ok = true;
        for (u64 f : fac) if (powmod(g, (q - 1) / f, q) == 1) { ok = false; break; }
        if (ok) return g;
    }
}

void build_mod(HostMod &m, u64 q) {
    const u64 N = HEC_N;
    m.q = q;
    u64 inv = 1;
    for (int i = 0; i < 7; i++) inv *= 2 - q * inv; // Newton: q*inv = 1 mod 2^64
    m.qinv = inv;
    m.rmod = mform(1, q);
    auto shoup = [q](u64 w) { return (u64)(((u128)w << 64) / q); };
    m.ninv_w = invmod(N, q);
    m.ninv_s = shoup(m.ninv_w);
    m.gen = primitive_root(q);
    u64 psi = powmod(m.gen, (q - 1) / (2 * N), q);
    u64 psi_i = invmod(psi, q);
    // NttPsi[brev(j)] = psi^j and NttPsiInv[brev(j)] = psi^-j (L:ring/ring.go:118-200), plain residues
    std::vector<u64> tab(N), tabi(N);
    u64 a = 1, b = 1;
    for (u32 j = 0; j < N; j++) {
        u32 r = bitrev16(j);
        tab[r] = a;
        tabi[r] = b;
        a = mulmod(a, psi, q);
        b = mulmod(b, psi_i, q);
    }
    m.psi_half_mont = mform(tab[1], q);
    // device order: the 15 twiddles {ng*base + gi} of every `base`, in consumption order (hec_dev.cuh fwd4)
    m.psi.assign(N, make_ulonglong2(0, 0));
    m.psi_inv.assign(N, make_ulonglong2(0, 0));
    auto put = [&](size_t dst, u32 base, int j) {
        int lg = 0;
        while ((2 << lg) - 1 <= j) lg++;      // slot j = (2^lg - 1) + gi
        u32 ng = 1u << lg, gi = (u32)j - (ng - 1);
        u32 idx = ng * base + gi;
        m.psi[dst] = make_ulonglong2(tab[idx], shoup(tab[idx]));
        m.psi_inv[dst] = make_ulonglong2(tabi[idx], shoup(tabi[idx]));
    };
    for (int j = 0; j < 15; j++) put(HEC_TW_COLA + j, 1, j);
    for (u32 g = 0; g < 16; g++)
        for (int j = 0; j < 15; j++) put(HEC_TW_COLB + 15 * g + j, 16 + g, j);
    for (u32 blk = 0; blk < 256; blk++)
        for (int j = 0; j < 15; j++) put(HEC_TW_ROWA + 15 * blk + j, 256 + blk, j);
    for (u32 blk = 0; blk < 256; blk++)
        for (int j = 0; j < 15; j++)
            for (u32 p = 0; p < 16; p++) put(HEC_TW_ROWB + 240 * blk + 16 * j + p, 4096 + 16 * blk + p, j);
}

// moduli below 2^31 take the 32-bit transforms and the narrow basis-extension products (HEC_NO_SMALL=1: A/B switch)
static bool narrow_modulus(u64 q) {
    static const bool off = getenv("HEC_NO_SMALL") != nullptr;
    return !off && q < (1ull << 31);
}
static void build_modup(const hec_ctx *c, ModupTab &T, const std::vector<int> &src) {
    int n = (int)src.size(), nt = c->nQ + c->nP;
    T.n = n;
    T.smod = src;
    T.qib.resize(n);
    T.qisp.assign(nt, std::vector<u64>(n));
    T.qpjinv.assign(nt, std::vector<u64>(n + 1));
    for (int i = 0; i < n; i++) {
        u64 qi = c->q(src[i]), star = 1;
        for (int k = 0; k < n; k++) if (k != i) star = mulmod(star, c->q(src[k]) % qi, qi);
        T.qib[i] = mform(invmod(star, qi), qi);
        for (int t = 0; t < nt; t++) {
            u64 pt = c->q(t), s = 1;
            for (int k = 0; k < n; k++) if (k != i) s = mulmod(s, c->q(src[k]) % pt, pt);
            // targets below 2^31 (the narrow path of k_modup2): plain residue | (residue * 2^32 mod p) << 32
            T.qisp[t][i] = narrow_modulus(pt) ? (s | (mulmod(s, (1ull << 32) % pt, pt) << 32)) : mform(s, pt);
        }
    }
    for (int t = 0; t < nt; t++) {
        u64 pt = c->q(t), Qm = 1;
        for (int k = 0; k < n; k++) Qm = mulmod(Qm, c->q(src[k]) % pt, pt);
        u64 v = pt - Qm;
        T.qpjinv[t][0] = 0;
        for (int i = 1; i <= n; i++) { u64 s = T.qpjinv[t][i - 1] + v; T.qpjinv[t][i] = s >= pt ? s - pt : s; }
    }
}

} // namespace hec

u64 *hec_ctx::scratch(size_t limbs) {
    if (arena_top + limbs > arena_limbs) { // a sizing bug in the caller's reserve(): never hand out memory past the arena
        fprintf(stderr, "libhec: scratch arena overrun (%zu + %zu > %zu limbs)\n", arena_top, limbs, arena_limbs);
        abort();
    }
    u64 *p = arena + arena_top * HEC_N;
    arena_top += limbs;
    return p;
}

static int reserve(hec_ctx *c, size_t limbs) {
    c->scratch_reset();
    if (limbs <= c->arena_limbs) return HEC_OK;
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->arena) cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_limbs = 0;
    HEC_CUDA(c, cudaMalloc(&c->arena, limbs * HEC_N * sizeof(u64)));
    c->arena_limbs = limbs;
    return HEC_OK;
}

// =========================================================================================
// launch helpers
// =========================================================================================
static int check_launch(hec_ctx *c, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->fail(HEC_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return HEC_OK;
}

// launch one of the generic kernels (all of them start with HEC_PDL_SYNC); they carry the programmatic-stream-
// serialisation attribute (HEC_PDL=0 turns it off) so that the launch latency and CTA ramp of a kernel overlap the tail of its predecessor
#ifndef HEC_FWD16_DEFAULT
#define HEC_FWD16_DEFAULT 0
#endif
#ifndef HEC_DOT_BULK_DEFAULT
#define HEC_DOT_BULK_DEFAULT 0 // measured (profiles/r02b): k_dot already streams at 5.6 TB/s = 0.87 of the measured HBM peak; the staged variant ties on the key switch and loses 2-8 % on evalReLU / the tap sums
#endif
template <typename... KArgs, typename... Args>
static void launch_k_smem(hec_ctx *c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
    static const int pdl = getenv("HEC_PDL") ? atoi(getenv("HEC_PDL")) : 1; // measured: -4.7 % key switch, -6 % CtoS / evalReLU
    if (!pdl) {
        kern<<<grid, block, smem, c->stream>>>(KArgs(args)...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
static void launch_k(hec_ctx *c, void (*kern)(KArgs...), dim3 grid, dim3 block, Args &&...args) {
    launch_k_smem(c, kern, grid, block, 0, std::forward<Args>(args)...);
}

static int stage_cached(hec_ctx *c, std::vector<char> &h, char **dev);
// forward / inverse NTT of a list of limbs (in -> out; in == out allowed)
int hec_launch_ntt(hec_ctx *c, std::vector<LimbJob> &jobs, bool inverse) {
    // limbs of the same modulus next to each other: their CTAs run back to back and share the twiddle table in L2
    // (the row transforms read as many table bytes as data bytes).  The jobs are independent, so the order is free.
    static const int sort_by_mod = getenv("HEC_NTT_SORT") ? atoi(getenv("HEC_NTT_SORT")) : 1; // measured: -1.3 % key switch
    if (sort_by_mod) std::stable_sort(jobs.begin(), jobs.end(), [](const LimbJob &a, const LimbJob &b) { return a.mod < b.mod; });
    // both passes' job tables go to the device in one content-addressed block (stage_cached): an operation repeated on
    // the same buffers launches without a copy, and one launch per pass takes all limbs (no per-launch job limit)
    const size_t n = jobs.size();
    if (n == 0) return HEC_OK;
    std::vector<char> h(2 * n * sizeof(LimbJob));
    LimbJob *A = reinterpret_cast<LimbJob *>(h.data()), *B = A + n;
    for (size_t i = 0; i < n; i++) {
        const LimbJob &j = jobs[i];
        u64 *mid = (!inverse && j.mid) ? j.mid : j.out;
        A[i] = j;
        B[i] = j;
        A[i].out = mid;              // first pass: in -> mid (prologue, if any)
        A[i].flags = inverse ? 0 : (j.flags & (HEC_LJ_PRO | HEC_LJ_PRO2));
        B[i].in = mid;               // second pass: mid -> out (epilogue, if any)
        B[i].flags = inverse ? 0 : (j.flags & (HEC_LJ_EPI | HEC_LJ_ADD | HEC_LJ_ADDS));
        A[i].scatter_g = 0;
        if (inverse) B[i].scatter_g = 0;
    }
    // HEC_FWD16=1: the forward transform as one launch per limb list, a limb per 16-CTA cluster exchanging through
    // distributed shared memory (k_fwd16); 0: column pass + row pass through global memory
    static const int fwd16 = getenv("HEC_FWD16") ? atoi(getenv("HEC_FWD16")) : HEC_FWD16_DEFAULT;
    bool ext_flags = false; // PRO2 / ADDS exist in the two-pass kernels only
    for (const LimbJob &j : jobs) ext_flags = ext_flags || (j.flags & (HEC_LJ_PRO2 | HEC_LJ_ADDS));
    if (!inverse && fwd16 && !ext_flags) {
        static int ready = 0; // 1: usable, -1: this device / driver refuses the cluster shape
        if (!ready) {
            bool ok = cudaFuncSetAttribute(k_fwd16, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                      cudaFuncSetAttribute(k_fwd16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_F16_SMEM) == cudaSuccess;
            cudaGetLastError();
            ready = ok ? 1 : -1;
        }
        if (ready == 1) {
            std::vector<char> hj(n * sizeof(LimbJob));
            LimbJob *J = reinterpret_cast<LimbJob *>(hj.data());
            for (size_t i = 0; i < n; i++) { J[i] = jobs[i]; J[i].mid = nullptr; }
            char *dj = nullptr;
            int rc1 = stage_cached(c, hj, &dj);
            if (rc1) return rc1;
            static const int pdl = getenv("HEC_PDL") ? atoi(getenv("HEC_PDL")) : 1;
            for (size_t off = 0; off < n; off += 65535) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(HEC_TILES_PER_LIMB, (unsigned)std::min<size_t>(65535, n - off));
                cfg.blockDim = dim3(HEC_THREADS); cfg.dynamicSmemBytes = HEC_F16_SMEM; cfg.stream = c->stream;
                cudaLaunchAttribute at[2];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = HEC_TILES_PER_LIMB; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[1].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
                cudaLaunchKernelEx(&cfg, k_fwd16, reinterpret_cast<const LimbJob *>(dj) + off, (const ModC *)c->dmods);
                c->launches += 1;
            }
            return check_launch(c, "fwd16");
        }
    }
    char *dbuf = nullptr;
    int rc = stage_cached(c, h, &dbuf);
    if (rc) return rc;
    const LimbJob *dA = reinterpret_cast<const LimbJob *>(dbuf), *dB = dA + n;
    for (size_t off = 0; off < n; off += 65535) { // grid.y limit
        dim3 grid(HEC_TILES_PER_LIMB, (unsigned)std::min<size_t>(65535, n - off));
        if (!inverse) {
            if (ext_flags) {
                launch_k(c, k_col_fwd<true>, grid, dim3(HEC_THREADS), dA + off, c->dmods);
                launch_k(c, k_row_fwd<true>, grid, dim3(HEC_THREADS), dB + off, c->dmods);
            } else {
                launch_k(c, k_col_fwd<false>, grid, dim3(HEC_THREADS), dA + off, c->dmods);
                launch_k(c, k_row_fwd<false>, grid, dim3(HEC_THREADS), dB + off, c->dmods);
            }
        } else {
            launch_k(c, k_row_inv, grid, dim3(HEC_THREADS), dA + off, c->dmods);
            launch_k(c, k_col_inv, grid, dim3(HEC_THREADS), dB + off, c->dmods);
        }
        c->launches += 2;
    }
    return check_launch(c, "ntt");
}

template <int OP>
static int launch_ew(hec_ctx *c, std::vector<EwJob> &jobs) {
    for (size_t off = 0; off < jobs.size(); off += HEC_EWJOBS) {
        int n = (int)std::min<size_t>(HEC_EWJOBS, jobs.size() - off);
        EwJobs J;
        for (int i = 0; i < n; i++) J.j[i] = jobs[off + i];
        // at least ~4 CTAs per SM: a launch with few limbs spreads each limb over more (shorter) CTAs
        unsigned gx = 32;
        while (gx < 256 && gx * (unsigned)n < 592) gx *= 2;
        launch_k(c, k_ew<OP>, dim3(gx, n), dim3(256), J, c->dmods);
        c->launches += 1;
    }
    return check_launch(c, "ew");
}
static EwJob ewjob(const u64 *a, const u64 *b, u64 *out, int mod, u64 s0 = 0, u32 g = 0) {
    EwJob j;
    j.a = a; j.b = b; j.out = out; j.mod = mod; j.g = g; j.s0 = s0; j.s1 = 0;
    return j;
}

// device copy of a small host table, cached by content (hec_ctx::staged).  The tables hold buffer addresses, which
// repeat as long as the caller repeats the operation on the same ciphertexts / scratch layout.
static int stage_cached(hec_ctx *c, std::vector<char> &h, char **dev) {
    h.resize((h.size() + 7) & ~(size_t)7, 0);
    uint64_t k = 1469598103934665603ull;
    const uint64_t *w = reinterpret_cast<const uint64_t *>(h.data());
    for (size_t i = 0, n = h.size() / 8; i < n; i++) { k ^= w[i]; k *= 1099511628211ull; k ^= k >> 29; }
    auto it = c->staged.find(k);
    if (it != c->staged.end())
        for (auto &e : it->second)
            if (e.host.size() == h.size() && memcmp(e.host.data(), h.data(), h.size()) == 0) { *dev = e.dev; return HEC_OK; }
    const size_t need = (h.size() + 255) & ~(size_t)255;
    hec_ctx::StagedTab e;
    const size_t ring = c->stage_ring, SLAB = c->stage_slab_bytes;
    if (need > SLAB) { // an unusually large table gets a block of its own (freed when the ring wraps)
        HEC_CUDA(c, cudaMalloc(&e.dev, need));
        c->stage_big.push_back(e.dev);
    } else {
        if (!c->stage_cur || c->stage_slab_top + need > SLAB) {
            if (c->stage_cur) {
                HEC_CUDA(c, cudaEventRecord(c->stage_events[c->stage_idx], c->stream)); // every launch that reads this slab is in the stream by now
                c->stage_idx = (c->stage_idx + 1) % ring;
            }
            if (c->stage_idx == c->stage_slabs.size()) { // the ring is still growing
                char *slab = nullptr;
                cudaEvent_t ev = nullptr;
                HEC_CUDA(c, cudaMalloc(&slab, SLAB));
                HEC_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                c->stage_slabs.push_back(slab);
                c->stage_events.push_back(ev);
                c->stage_keys.emplace_back();
            } else {                                     // come round: retire what this slab held
                HEC_CUDA(c, cudaEventSynchronize(c->stage_events[c->stage_idx]));
                char *lo = c->stage_slabs[c->stage_idx], *hi = lo + SLAB;
                for (uint64_t key : c->stage_keys[c->stage_idx]) {
                    auto f = c->staged.find(key);
                    if (f == c->staged.end()) continue;
                    auto &v = f->second;
                    v.erase(std::remove_if(v.begin(), v.end(), [&](const hec_ctx::StagedTab &t) { return t.dev >= lo && t.dev < hi; }), v.end());
                    if (v.empty()) c->staged.erase(f);
                }
                c->stage_keys[c->stage_idx].clear();
                if (c->stage_idx == 0 && !c->stage_big.empty()) { // the oversized blocks go once per turn of the ring
                    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
                    for (auto itb = c->staged.begin(); itb != c->staged.end();) {
                        auto &v = itb->second;
                        v.erase(std::remove_if(v.begin(), v.end(), [&](const hec_ctx::StagedTab &t) {
                                    return std::find(c->stage_big.begin(), c->stage_big.end(), t.dev) != c->stage_big.end(); }), v.end());
                        itb = v.empty() ? c->staged.erase(itb) : std::next(itb);
                    }
                    for (char *pb : c->stage_big) cudaFree(pb);
                    c->stage_big.clear();
                }
            }
            c->stage_cur = c->stage_slabs[c->stage_idx];
            c->stage_slab_top = 0;
        }
        e.dev = c->stage_cur + c->stage_slab_top;
        c->stage_slab_top += need;
        c->stage_keys[c->stage_idx].push_back(k);
    }
    e.host = h;
    // pageable source: staged by the driver before the call returns
    HEC_CUDA(c, cudaMemcpyAsync(e.dev, e.host.data(), h.size(), cudaMemcpyHostToDevice, c->stream));
    *dev = e.dev;
    c->staged[k].push_back(std::move(e));
    return HEC_OK;
}

// launch the exact basis extensions `jobs` (one entry per target limb): entries that share their source digit are
// grouped so that k_modup2 computes y_i and v once per coefficient for all of the digit's targets; the grouped tables
// are staged into one stream-ordered device buffer
static int launch_modup(hec_ctx *c, std::vector<ModupJob> &jobs) {
    if (jobs.empty()) return HEC_OK;
    std::vector<Modup2Job> groups;
    std::vector<Modup2Target> targets;
    std::vector<size_t> first; // index of each group's first target
    for (const ModupJob &j : jobs) {
        bool same = !groups.empty() && groups.back().n == j.n && groups.back().ntargets < HEC_M2_MAXT;
        for (int s = 0; same && s < j.n; s++) same = groups.back().src[s] == j.src[s] && groups.back().smod[s] == j.smod[s];
        if (!same) {
            Modup2Job g;
            memset(&g, 0, sizeof g);
            g.n = j.n;
            for (int s = 0; s < j.n; s++) { g.src[s] = j.src[s]; g.smod[s] = j.smod[s]; g.qib[s] = j.qib[s]; }
            groups.push_back(g);
            first.push_back(targets.size());
        }
        Modup2Target t;
        memset(&t, 0, sizeof t);
        t.dst = j.dst; t.tmod = j.tmod;
        for (int s = 0; s < j.n; s++) t.qisp[s] = j.qisp[s];
        for (int v = 0; v <= j.n; v++) t.qpjinv[v] = j.qpjinv[v];
        targets.push_back(t);
        groups.back().ntargets++;
    }
    size_t tb = targets.size() * sizeof(Modup2Target), gb = groups.size() * sizeof(Modup2Job);
    char *dbuf = nullptr;
    for (size_t g = 0; g < groups.size(); g++) groups[g].first = first[g];
    std::vector<char> h(tb + gb);
    memcpy(h.data(), targets.data(), tb);
    memcpy(h.data() + tb, groups.data(), gb);
    int rc = stage_cached(c, h, &dbuf);
    if (rc) return rc;
    for (size_t off = 0; off < groups.size(); off += 65535) {
        unsigned n = (unsigned)std::min<size_t>(65535, groups.size() - off);
        launch_k(c, k_modup2, dim3(HEC_N / 256, n), dim3(256), reinterpret_cast<const Modup2Job *>(dbuf + tb) + off,
                 reinterpret_cast<const Modup2Target *>(dbuf), c->dmods);
        c->launches += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->fail(HEC_E_CUDA, std::string("modup: ") + cudaGetErrorString(e));
    return HEC_OK;
}
static ModupJob modup_job(const hec_ctx *c, const ModupTab &T, const u64 *src, size_t src_stride, int target, u64 *dst) {
    ModupJob j;
    memset(&j, 0, sizeof j);
    j.n = T.n;
    for (int s = 0; s < T.n; s++) {
        j.src[s] = src + s * src_stride;
        j.smod[s] = T.smod[s];
        j.qib[s] = T.qib[s];
        j.qisp[s] = T.qisp[target][s];
    }
    for (int v = 0; v <= T.n; v++) j.qpjinv[v] = T.qpjinv[target][v];
    j.dst = dst;
    j.tmod = target;
    (void)c;
    return j;
}

// =========================================================================================
// context
// =========================================================================================
extern "C" const char *hec_version(void) { return "libhec 0.1 (sm_100a)"; }

extern "C" int hec_ctx_create(hec_ctx **out, int logN, const uint64_t *Q, int nQ, const uint64_t *P, int nP, int device) {
    if (!out || !Q || nQ < 1 || nP < 0 || (nP > 0 && !P)) return HEC_E_INVAL;
    *out = nullptr;
    if (logN != HEC_LOGN) return HEC_E_UNSUPPORTED;
    if (nP > HEC_MAXA) return HEC_E_UNSUPPORTED;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device || device < 0) return HEC_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return HEC_E_CUDA;
    hec_ctx *c = new hec_ctx();
    c->device = device; c->nQ = nQ; c->nP = nP;
    int nm = nQ + nP;
    c->hm.resize(nm);
    for (int i = 0; i < nm; i++) {
        u64 q = i < nQ ? Q[i] : P[i - nQ];
        if ((q & 1) == 0 || q >= (1ull << 61) || (q - 1) % (2ull * HEC_N) != 0) { delete c; return HEC_E_INVAL; }
        build_mod(c->hm[i], q);
    }
    auto bail = [&](int code) { hec_ctx_destroy(c); return code; };
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(HEC_E_CUDA);
    // test hooks: a tiny ring makes the job-table cache wrap (and overflow into blocks of their own) within a few operations
    if (getenv("HEC_STAGE_SLAB_KB")) c->stage_slab_bytes = std::max<size_t>(4, (size_t)atol(getenv("HEC_STAGE_SLAB_KB"))) << 10;
    if (getenv("HEC_STAGE_RING")) c->stage_ring = std::max<size_t>(2, (size_t)atol(getenv("HEC_STAGE_RING")));
    {   // a pool of the context's own for the stream-ordered allocations (ciphertexts, plaintexts): freed buffers stay
        // cached in it instead of going back to the driver, and nothing of that outlives the context or touches the
        // device's default pool, which the host process may share
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&c->pool, &props) != cudaSuccess) return bail(HEC_E_CUDA);
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) return bail(HEC_E_CUDA);
    size_t tw = (size_t)nm * 2 * HEC_N;
    if (cudaMalloc(&c->dtables, tw * sizeof(ulonglong2)) != cudaSuccess) return bail(HEC_E_NOMEM);
    if (cudaMalloc(&c->dmods, nm * sizeof(ModC)) != cudaSuccess) return bail(HEC_E_NOMEM);
    std::vector<ModC> mc(nm);
    for (int i = 0; i < nm; i++) {
        ulonglong2 *psi = c->dtables + (size_t)i * 2 * HEC_N, *psi_inv = psi + HEC_N;
        if (cudaMemcpy(psi, c->hm[i].psi.data(), HEC_N * sizeof(ulonglong2), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(psi_inv, c->hm[i].psi_inv.data(), HEC_N * sizeof(ulonglong2), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(HEC_E_CUDA);
        mc[i].q = c->hm[i].q; mc[i].qinv = c->hm[i].qinv; mc[i].q2 = 2 * c->hm[i].q;
        mc[i].rmod = c->hm[i].rmod; mc[i].ninv_w = c->hm[i].ninv_w; mc[i].ninv_s = c->hm[i].ninv_s;
        mc[i].wn_w = mulmod(c->hm[i].psi_inv[HEC_TW_COLA].x, c->hm[i].ninv_w, c->hm[i].q);
        mc[i].wn_s = (u64)(((u128)mc[i].wn_w << 64) / c->hm[i].q);
        mc[i].psi = psi; mc[i].psi_inv = psi_inv;
        mc[i].tight = c->hm[i].q >= (1ull << 57) ? 1 : 0;
        mc[i].mu = c->hm[i].q > (1ull << 40) ? (u32)(((u128)1 << 64) / c->hm[i].q) : 0; mc[i].pad = 0;
        mc[i].small = narrow_modulus(c->hm[i].q) ? 1 : 0;
    }
    if (cudaMemcpy(c->dmods, mc.data(), nm * sizeof(ModC), cudaMemcpyHostToDevice) != cudaSuccess) return bail(HEC_E_CUDA);
    // RescaleParams (L:ring/ring.go:63-117)
    c->resc.resize(nQ);
    for (int L = 1; L < nQ; L++) {
        c->resc[L].resize(L);
        for (int i = 0; i < L; i++) {
            u64 qi = Q[i];
            c->resc[L][i] = mform(qi - invmod(Q[L] % qi, qi), qi);
        }
    }
    if (nP > 0) {
        c->alpha = nP;
        c->beta_full = (nQ + nP - 1) / nP;
        c->negpinv.resize(nQ);
        for (int i = 0; i < nQ; i++) {
            u64 qi = Q[i], Pm = 1;
            for (int j = 0; j < nP; j++) Pm = mulmod(Pm, P[j] % qi, qi);
            c->negpinv[i] = qi - mform(invmod(Pm, qi), qi); // [A] test_run 0x4e5049: params := qi - modDownParams[i]
        }
        std::vector<int> src;
        for (int j = 0; j < nP; j++) src.push_back(c->modP(j));
        build_modup(c, c->pq, src);
        // Decomposer tables (L:ring/ring_basis_extension.go:482-538)
        c->xalpha.assign(c->beta_full, nP);
        if (nQ % nP) c->xalpha[c->beta_full - 1] = nQ % nP;
        c->dec.resize(c->beta_full);
        for (int d = 0; d < c->beta_full; d++) {
            c->dec[d].resize(nP + 1);
            for (int nd = 2; nd <= c->xalpha[d]; nd++) {
                src.clear();
                for (int k = 0; k < nd; k++) src.push_back(d * nP + k);
                build_modup(c, c->dec[d][nd], src);
            }
        }
    }
    *out = c;
    return HEC_OK;
}

extern "C" void hec_ctx_destroy(hec_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    hec_plan_cache_clear(c);
    for (auto &kv : c->keys) { cudaFree(kv.second.buf); delete kv.second.kb; }
    for (char *p : c->stage_slabs) cudaFree(p);
    for (char *p : c->stage_big) cudaFree(p);
    for (cudaEvent_t ev : c->stage_events) cudaEventDestroy(ev);
    if (c->arena) cudaFree(c->arena);
    if (c->dtables) cudaFree(c->dtables);
    if (c->dmods) cudaFree(c->dmods);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->pool) cudaMemPoolDestroy(c->pool);
    delete c;
}
extern "C" const char *hec_last_error(const hec_ctx *c) { return c ? c->err.c_str() : "null context"; }
extern "C" int hec_sync(hec_ctx *c) {
    if (!c) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}
extern "C" int hec_timer_start(hec_ctx *c) {
    if (!c) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    HEC_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    return HEC_OK;
}
extern "C" int hec_timer_stop_ms(hec_ctx *c, float *ms) {
    if (!c || !ms) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    HEC_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    HEC_CUDA(c, cudaEventSynchronize(c->ev1));
    HEC_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return HEC_OK;
}
extern "C" uint64_t hec_launch_count(const hec_ctx *c) { return c ? c->launches : 0; }
// page-lock caller memory (e.g. the backing arrays of ring.Poly.Coeffs: Go's heap does not move objects)
// so that uploads / downloads / hec_plan_submit_host copy asynchronously at full PCIe rate
extern "C" int hec_host_register(hec_ctx *c, void *ptr, size_t bytes) {
    if (!c || !ptr || !bytes) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    HEC_CUDA(c, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return HEC_OK;
}
extern "C" int hec_host_unregister(hec_ctx *c, void *ptr) {
    if (!c || !ptr) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    HEC_CUDA(c, cudaHostUnregister(ptr));
    return HEC_OK;
}

// =========================================================================================
// handles
// =========================================================================================
int hec_ct_alloc(hec_ctx *c, int level, double scale, hec_ct **out) {
    hec_ct *ct = new hec_ct();
    ct->alloc = level + 1; ct->level = level; ct->scale = scale;
    // stream-ordered allocation: temporaries of the op-level path come from the device's memory pool and
    // are recycled without synchronising the stream
    if (cudaMallocFromPoolAsync(&ct->buf, (size_t)2 * ct->alloc * HEC_N * sizeof(u64), c->pool, c->stream) != cudaSuccess) {
        delete ct;
        return c->fail(HEC_E_NOMEM, "cudaMallocAsync ciphertext");
    }
    *out = ct;
    return HEC_OK;
}

extern "C" int hec_pt_upload(hec_ctx *c, int level, const uint64_t *const *limbs, double scale, hec_pt **out) {
    if (!c || !limbs || !out || level < 0 || level >= c->nQ) return c ? c->fail(HEC_E_INVAL, "pt_upload args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_pt *pt = new hec_pt();
    pt->level = level; pt->scale = scale; pt->serial = c->next_serial++;
    if (cudaMallocFromPoolAsync(&pt->buf, (size_t)(level + 1) * HEC_N * sizeof(u64), c->pool, c->stream) != cudaSuccess) { delete pt; return c->fail(HEC_E_NOMEM, "cudaMallocAsync plaintext"); }
    std::vector<EwJob> jobs;
    for (int i = 0; i <= level; i++) {
        u64 *d = pt->buf + (size_t)i * HEC_N;
        HEC_CUDA(c, cudaMemcpyAsync(d, limbs[i], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        u64 q = c->q(i);
        jobs.push_back(ewjob(d, nullptr, d, i, mform(c->hm[i].rmod, q))); // MFormLvl: * R^2 * R^-1
    }
    int rc = launch_ew<EW_TOMONT>(c, jobs);
    if (rc) return rc;
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = pt;
    return HEC_OK;
}
extern "C" void hec_pt_free(hec_ctx *c, hec_pt *pt) {
    if (!pt) return;
    if (c && !c->plan_cache.empty()) hec_plan_cache_evict(c, pt->serial); // a cached plan is identified by its plaintexts
    if (c) {
        cudaSetDevice(c->device);
        cudaFreeAsync(pt->buf, c->stream); // ordered after every kernel of this context that read it
    } else {
        cudaFree(pt->buf);
    }
    delete pt;
}

extern "C" int hec_ct_upload(hec_ctx *c, int level, const uint64_t *const *c0, const uint64_t *const *c1, double scale, hec_ct **out) {
    if (!c || !c0 || !c1 || !out || level < 0 || level >= c->nQ) return c ? c->fail(HEC_E_INVAL, "ct_upload args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *ct = nullptr;
    int rc = hec_ct_alloc(c, level, scale, &ct);
    if (rc) return rc;
    for (int i = 0; i <= level; i++) {
        HEC_CUDA(c, cudaMemcpyAsync(ct->limb(0, i), c0[i], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
        HEC_CUDA(c, cudaMemcpyAsync(ct->limb(1, i), c1[i], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    }
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = ct;
    return HEC_OK;
}
extern "C" int hec_ct_download(hec_ctx *c, const hec_ct *ct, uint64_t *const *c0, uint64_t *const *c1) {
    if (!c || !ct || !c0 || !c1) return c ? c->fail(HEC_E_INVAL, "ct_download args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    for (int i = 0; i <= ct->level; i++) {
        HEC_CUDA(c, cudaMemcpyAsync(c0[i], ct->limb(0, i), HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        HEC_CUDA(c, cudaMemcpyAsync(c1[i], ct->limb(1, i), HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    }
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}
extern "C" int hec_ct_copy_new(hec_ctx *c, const hec_ct *ct, hec_ct **out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *n = nullptr;
    int rc = hec_ct_alloc(c, ct->level, ct->scale, &n);
    if (rc) return rc;
    for (int p = 0; p < 2; p++)
        HEC_CUDA(c, cudaMemcpyAsync(n->limb(p, 0), ct->limb(p, 0), (size_t)(ct->level + 1) * HEC_N * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
    *out = n;
    return HEC_OK;
}
extern "C" int hec_ct_level(const hec_ct *ct) { return ct ? ct->level : -1; }
extern "C" double hec_ct_scale(const hec_ct *ct) { return ct ? ct->scale : 0.0; }
extern "C" void hec_ct_set_scale(hec_ct *ct, double s) { if (ct) ct->scale = s; }
extern "C" void hec_ct_free(hec_ctx *c, hec_ct *ct) {
    if (!ct) return;
    if (c) {
        cudaSetDevice(c->device);
        if (ct->owned) cudaFreeAsync(ct->buf, c->stream); // ordered after every kernel that used it
    } else if (ct->owned) {
        cudaFree(ct->buf);
    }
    delete ct;
}

extern "C" int hec_swk_upload(hec_ctx *c, uint64_t galEl, int max_level, const uint64_t *const *limbs) {
    if (!c || !limbs || c->nP == 0 || max_level < 0 || max_level >= c->nQ) return c ? c->fail(HEC_E_INVAL, "swk_upload args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_swk_drop(c, galEl);
    SwKey k;
    k.Lk = max_level + 1;
    k.ndig = (k.Lk + c->alpha - 1) / c->alpha;
    int kl = k.Lk + c->nP, full = c->nQ + c->nP;
    if (cudaMalloc(&k.buf, (size_t)k.ndig * 2 * kl * HEC_N * sizeof(u64)) != cudaSuccess) return c->fail(HEC_E_NOMEM, "cudaMalloc key");
    for (int d = 0; d < k.ndig; d++)
        for (int p = 0; p < 2; p++)
            for (int t = 0; t < kl; t++) {
                int srct = t < k.Lk ? t : c->nQ + (t - k.Lk);
                const u64 *src = (const u64 *)limbs[(size_t)(d * 2 + p) * full + srct];
                u64 *dst = k.buf + ((size_t)(d * 2 + p) * kl + t) * HEC_N;
                HEC_CUDA(c, cudaMemcpyAsync(dst, src, HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
            }
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    k.kb = new KeyBuf();
    k.kb->buf = k.buf; k.kb->serial = c->next_serial++;
    c->keys[galEl] = k;
    return HEC_OK;
}
extern "C" int hec_swk_drop(hec_ctx *c, uint64_t galEl) {
    if (!c) return HEC_E_INVAL;
    auto it = c->keys.find(galEl);
    if (it == c->keys.end()) return HEC_OK;
    cudaSetDevice(c->device);
    KeyBuf *kb = it->second.kb;
    c->keys.erase(it);
    if (!c->plan_cache.empty()) hec_plan_cache_evict(c, kb->serial);
    if (kb->plan_refs > 0) { kb->dead = true; return HEC_OK; } // a caller's plan still reads it: freed with that plan
    cudaStreamSynchronize(c->stream);
    cudaFree(kb->buf);
    delete kb;
    return HEC_OK;
}

// =========================================================================================
// generic evaluator ops
// =========================================================================================
extern "C" int hec_mul_pt_new(hec_ctx *c, const hec_ct *ct, const hec_pt *pt, hec_ct **out) {
    if (!c || !ct || !pt || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    int level = std::min(ct->level, pt->level);
    hec_ct *o = nullptr;
    int rc = hec_ct_alloc(c, level, ct->scale * pt->scale, &o);
    if (rc) return rc;
    std::vector<EwJob> jobs;
    for (int p = 0; p < 2; p++)
        for (int i = 0; i <= level; i++) jobs.push_back(ewjob(ct->limb(p, i), pt->buf + (size_t)i * HEC_N, o->limb(p, i), i));
    rc = launch_ew<EW_MULMONT>(c, jobs);
    if (rc) { hec_ct_free(c, o); return rc; }
    *out = o;
    return HEC_OK;
}

// scaleUpExact (L:ckks/utils.go:31-57): floor(|c*n| + 0.5) mod q with sign; prec-53
// big.Float arithmetic == IEEE double
static u64 scale_up_exact(double value, double n, u64 q) {
    bool neg = value < 0;
    volatile double x = neg ? -n * value : n * value;
    volatile double y = x + 0.5;
    double fl = floor(y);
    u64 res;
    if (fl < 18446744073709551616.0) res = (u64)fl % q;
    else {
        int e;
        double m = frexp(fl, &e);
        u64 mant = (u64)ldexp(m, 53);
        res = mant % q;
        for (int i = 0; i < e - 53; i++) res = (u64)(((u128)res * 2) % q);
    }
    return neg ? (res ? q - res : 0) : res;
}
// getConstAndScale, float64 case (L:ckks/evaluator.go:508-561)
double hec_const_limbs(const hec_ctx *c, int level, double constant, std::vector<u64> &k) {
    double up = 1.0;
    if (constant != 0) {
        double vi = (double)(int64_t)constant;
        if (constant - vi != 0) up = (double)c->q(level);
    }
    k.resize(level + 1);
    for (int i = 0; i <= level; i++) k[i] = constant != 0 ? scale_up_exact(constant, up, c->q(i)) : 0;
    return up;
}

extern "C" int hec_mult_by_const(hec_ctx *c, hec_ct *ct, double constant) {
    if (!c || !ct) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    std::vector<u64> k;
    double up = hec_const_limbs(c, ct->level, constant, k);
    std::vector<EwJob> jobs;
    for (int p = 0; p < 2; p++)
        for (int i = 0; i <= ct->level; i++) jobs.push_back(ewjob(ct->limb(p, i), nullptr, ct->limb(p, i), i, mform(k[i], c->q(i))));
    int rc = launch_ew<EW_MULSCALAR>(c, jobs);
    if (rc) return rc;
    ct->scale *= up;
    return HEC_OK;
}

// one DivRoundByLastModulusNTT on both polys (L:ring/ring_scaling.go:442-513) of every ciphertext of a batch
// at a common level (one launch sequence for all of them)
static int div_round_last_many(hec_ctx *c, const std::vector<hec_ct *> &cts) {
    int L = cts[0]->level, rc;
    size_t n = cts.size();
    if ((rc = reserve(c, n * (2 + 2 * (size_t)L)))) return rc;
    u64 qL = c->q(L), half = (qL - 1) >> 1;
    std::vector<u64 *> t(n), u(n);
    std::vector<LimbJob> nj;
    std::vector<EwJob> ej;
    for (size_t m = 0; m < n; m++) {
        t[m] = c->scratch(2); u[m] = c->scratch(2 * (size_t)L);
        for (int p = 0; p < 2; p++) {
            nj.push_back({cts[m]->limb(p, L), t[m] + (size_t)p * HEC_N, L, 0});
            ej.push_back(ewjob(t[m] + (size_t)p * HEC_N, nullptr, t[m] + (size_t)p * HEC_N, L, half));
        }
    }
    if ((rc = hec_launch_ntt(c, nj, true))) return rc;
    if ((rc = launch_ew<EW_CENTER>(c, ej))) return rc;
    ej.clear();
    nj.clear();
    for (size_t m = 0; m < n; m++)
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < L; i++) {
                u64 qi = c->q(i);
                u64 *ui = u[m] + ((size_t)p * L + i) * HEC_N;
                // (t mod q_i) + (q_i - half) in the transform's prologue, (u + 2q - ct) * (-q_L^-1) in its epilogue
                // (were an EW_REDUCE_ADD pass in front and an EW_SUBMUL pass behind)
                nj.push_back({t[m] + (size_t)p * HEC_N, cts[m]->limb(p, i), i, HEC_LJ_PRO | HEC_LJ_EPI, ui, cts[m]->limb(p, i),
                              c->resc[L][i], qi - half % qi});
            }
    if ((rc = hec_launch_ntt(c, nj, false))) return rc;
    for (size_t m = 0; m < n; m++) cts[m]->level = L - 1;
    return HEC_OK;
}

// Rescale(ct, minScale, ct) (L:ckks/evaluator.go:1291-1325) for a batch sharing level and scale
int hec_rescale_many(hec_ctx *c, const std::vector<hec_ct *> &cts, double min_scale) {
    if (cts.empty()) return HEC_OK;
    for (hec_ct *x : cts)
        if (x->level != cts[0]->level || x->scale != cts[0]->scale) return c->fail(HEC_E_INVAL, "batched Rescale needs a common level and scale");
    if (cts[0]->level == 0) return c->fail(HEC_E_LEVEL, "cannot Rescale: input Ciphertext already at level 0");
    while (cts[0]->level > 0 && cts[0]->scale / (double)c->q(cts[0]->level) >= min_scale / 2) {
        double s = cts[0]->scale / (double)c->q(cts[0]->level);
        int rc = div_round_last_many(c, cts);
        if (rc) return rc;
        for (hec_ct *x : cts) x->scale = s;
    }
    return HEC_OK;
}
extern "C" int hec_rescale(hec_ctx *c, hec_ct *ct, double min_scale) {
    if (!c || !ct) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    return hec_rescale_many(c, {ct}, min_scale);
}

extern "C" int hec_set_scale(hec_ctx *c, hec_ct *ct, double scale) {
    // L:ckks/evaluator.go:1194-1209
    if (!c || !ct) return c ? c->fail(HEC_E_INVAL, "set_scale args") : HEC_E_INVAL;
    int rc = hec_mult_by_const(c, ct, scale / ct->scale);
    if (rc) return rc;
    if ((rc = hec_rescale(c, ct, scale))) return rc;
    ct->scale = scale;
    return HEC_OK;
}

// Add / Sub (evaluateInPlace, L:ckks/evaluator.go:365-473): level = min, scale = max; when the scales differ
// by a factor whose floor is > 1 the smaller-scale operand is first multiplied by that integer (MultByConst),
// in place when it is also the receiver, else into scratch.
template <int OP>
static int addsub(hec_ctx *c, const hec_ct *a, const hec_ct *b, hec_ct *out) {
    int level = std::min(std::min(a->level, b->level), out->alloc - 1), L = level + 1, rc;
    const hec_ct *lo = nullptr; // the operand to scale up
    double factor = 1.0;
    if (a->scale > b->scale && floor(a->scale / b->scale) > 1) { lo = b; factor = floor(a->scale / b->scale); }
    else if (b->scale > a->scale && floor(b->scale / a->scale) > 1) { lo = a; factor = floor(b->scale / a->scale); }
    std::vector<const u64 *> xa(2 * L), xb(2 * L);
    for (int p = 0; p < 2; p++)
        for (int i = 0; i < L; i++) { xa[p * L + i] = a->limb(p, i); xb[p * L + i] = b->limb(p, i); }
    if (lo) {
        std::vector<u64> k;
        hec_const_limbs(c, level, factor, k);
        u64 *tmp = nullptr;
        if (lo != out) {
            if ((rc = reserve(c, 2 * (size_t)L))) return rc;
            tmp = c->scratch(2 * (size_t)L);
        }
        std::vector<EwJob> mj;
        std::vector<const u64 *> &x = lo == a ? xa : xb;
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < L; i++) {
                u64 *d = tmp ? tmp + (size_t)(p * L + i) * HEC_N : out->limb(p, i);
                mj.push_back(ewjob(lo->limb(p, i), nullptr, d, i, mform(k[i], c->q(i))));
                x[p * L + i] = d;
            }
        if ((rc = launch_ew<EW_MULSCALAR>(c, mj))) return rc;
    }
    std::vector<EwJob> jobs;
    for (int p = 0; p < 2; p++)
        for (int i = 0; i < L; i++) jobs.push_back(ewjob(xa[p * L + i], xb[p * L + i], out->limb(p, i), i));
    if ((rc = launch_ew<OP>(c, jobs))) return rc;
    out->level = level;
    out->scale = std::max(a->scale, b->scale);
    return HEC_OK;
}
extern "C" int hec_add(hec_ctx *c, const hec_ct *a, const hec_ct *b, hec_ct *out) {
    if (!c || !a || !b || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    return addsub<EW_ADD>(c, a, b, out);
}
extern "C" int hec_sub(hec_ctx *c, const hec_ct *a, const hec_ct *b, hec_ct *out) {
    if (!c || !a || !b || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    return addsub<EW_SUB>(c, a, b, out);
}
extern "C" int hec_add_new(hec_ctx *c, const hec_ct *a, const hec_ct *b, hec_ct **out) {
    if (!c || !a || !b || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *o = nullptr;
    int rc = hec_ct_alloc(c, std::min(a->level, b->level), a->scale, &o);
    if (rc) return rc;
    if ((rc = addsub<EW_ADD>(c, a, b, o))) { hec_ct_free(c, o); return rc; }
    *out = o;
    return HEC_OK;
}
extern "C" int hec_sub_new(hec_ctx *c, const hec_ct *a, const hec_ct *b, hec_ct **out) {
    if (!c || !a || !b || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *o = nullptr;
    int rc = hec_ct_alloc(c, std::min(a->level, b->level), a->scale, &o);
    if (rc) return rc;
    if ((rc = addsub<EW_SUB>(c, a, b, o))) { hec_ct_free(c, o); return rc; }
    *out = o;
    return HEC_OK;
}
// Add(ct, pt, ct): evaluateInPlace's scale matching as in addsub above (L:ckks/evaluator.go:365-473): when the scales
// differ by a factor whose floor is > 1 the smaller-scale operand is multiplied by that integer first -- the
// ciphertext in place (it is the receiver), the plaintext into scratch; the result carries the larger scale
extern "C" int hec_add_pt(hec_ctx *c, hec_ct *ct, const hec_pt *pt) {
    if (!c || !ct || !pt) return c ? c->fail(HEC_E_INVAL, "add_pt args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    const int level = std::min(ct->level, pt->level), L = level + 1;
    int rc;
    std::vector<const u64 *> pb(L);
    for (int i = 0; i < L; i++) pb[i] = pt->buf + (size_t)i * HEC_N;
    if (ct->scale > pt->scale && floor(ct->scale / pt->scale) > 1) {
        std::vector<u64> k;
        hec_const_limbs(c, level, floor(ct->scale / pt->scale), k);
        if ((rc = reserve(c, (size_t)L))) return rc;
        u64 *tmp = c->scratch((size_t)L);
        std::vector<EwJob> mj;
        for (int i = 0; i < L; i++) {
            mj.push_back(ewjob(pb[i], nullptr, tmp + (size_t)i * HEC_N, i, mform(k[i], c->q(i)))); // stays in Montgomery form
            pb[i] = tmp + (size_t)i * HEC_N;
        }
        if ((rc = launch_ew<EW_MULSCALAR>(c, mj))) return rc;
    } else if (pt->scale > ct->scale && floor(pt->scale / ct->scale) > 1) {
        std::vector<u64> k;
        hec_const_limbs(c, level, floor(pt->scale / ct->scale), k);
        std::vector<EwJob> mj;
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < L; i++) mj.push_back(ewjob(ct->limb(p, i), nullptr, ct->limb(p, i), i, mform(k[i], c->q(i))));
        if ((rc = launch_ew<EW_MULSCALAR>(c, mj))) return rc;
    }
    std::vector<EwJob> jobs;
    for (int i = 0; i < L; i++) jobs.push_back(ewjob(ct->limb(0, i), pb[i], ct->limb(0, i), i));
    if ((rc = launch_ew<EW_ADD_MONT>(c, jobs))) return rc;
    ct->level = level;
    ct->scale = std::max(ct->scale, pt->scale);
    return HEC_OK;
}

// ---- key switching -----------------------------------------------------------------------
// Every step below is written for a BATCH of independent key-switches at the same level (one
// launch sequence for all of them): the k^2 hoisted rotations of preConv_BL, the B/2-1 channel
// rotations of evalConv_BN_BL_test, or a single RotateGal (batch of one).

// ModDownSplitNTTPQ (L:ring/ring_basis_extension.go:247-291), in place on accQ[i] ([L][N]);
// accP[i] is [nP][N].  Scratch: L limbs per item.
// addQ (optional): per item a polynomial [L][N] added to the result (the "+ c0" of a rotation, L:ckks/evaluator.go:1590)
// outQ (optional): per item where the result goes instead of accQ (may be the same polynomial as addQ: each thread
// reads its element before it writes it)
// galQ (optional, with outQ pointing elsewhere than accQ / addQ): per item a Galois element g; the result is stored as
// sigma_g of itself (PermuteNTTWithIndexLvl folded into the last pass)
static int moddown_many(hec_ctx *c, int level, const std::vector<u64 *> &accQ, const std::vector<u64 *> &accP,
                        const std::vector<const u64 *> &addQ = {}, const std::vector<u64 *> &outQ = {},
                        const std::vector<u64> &galQ = {}) {
    int L = level + 1, nP = c->nP, rc;
    size_t n = accQ.size();
    std::vector<LimbJob> nj;
    for (size_t i = 0; i < n; i++)
        for (int j = 0; j < nP; j++) nj.push_back({accP[i] + (size_t)j * HEC_N, accP[i] + (size_t)j * HEC_N, c->modP(j), 0});
    if ((rc = hec_launch_ntt(c, nj, true))) return rc; // InvNTT on P (canonical residues of the lazy original)
    u64 *x = c->scratch(n * L);
    std::vector<ModupJob> mj;
    nj.clear();
    for (size_t i = 0; i < n; i++)
        for (int l = 0; l < L; l++) {
            u64 *xi = x + (i * L + l) * HEC_N;
            u64 *aq = accQ[i] + (size_t)l * HEC_N;
            mj.push_back(modup_job(c, c->pq, accP[i], HEC_N, c->modQ(l), xi));
            // NTT(xi), then (xi + 2q - accQ) * (-P^-1) in the transform's epilogue (was a separate EW_SUBMUL pass)
            const u64 *add = (i < addQ.size() && addQ[i]) ? addQ[i] + (size_t)l * HEC_N : nullptr;
            u64 *dst = (i < outQ.size() && outQ[i]) ? outQ[i] + (size_t)l * HEC_N : aq;
            u32 sg = 0;
            if (i < galQ.size() && galQ[i]) { // g^-1 mod 2N: g is odd, 2N a power of two (Newton)
                u64 g = galQ[i], inv = g;
                for (int it = 0; it < 6; it++) inv *= 2 - g * inv;
                sg = (u32)(inv & (2ull * HEC_N - 1));
            }
            nj.push_back({xi, dst, l, HEC_LJ_EPI | (add ? HEC_LJ_ADD : 0), xi, aq, c->negpinv[l], 0, add, sg});
        }
    if ((rc = launch_modup(c, mj))) return rc;
    return hec_launch_ntt(c, nj, false);
}

struct Decomp { u64 *D; int L, beta; const u64 *src; }; // src: the decomposed polynomial itself ([L][N], NTT domain)
static size_t decomp_limbs(const hec_ctx *c, int level) {
    int L = level + 1, beta = (L + c->alpha - 1) / c->alpha;
    return (size_t)L + (size_t)beta * (L + c->nP);
}
// DecomposeNTT (L:rlwe/keyswitch.go:94-141; DecomposeAndSplit L:ring/ring_basis_extension.go:543-664)
// of every c1[i] ([L][N], NTT domain).
static int decompose_many(hec_ctx *c, int level, const std::vector<const u64 *> &c1, std::vector<Decomp> &out) {
    int L = level + 1, nP = c->nP, alpha = c->alpha, rc;
    int beta = (L + alpha - 1) / alpha, W = L + nP;
    size_t n = c1.size();
    out.resize(n);
    std::vector<LimbJob> inv, fwd;
    std::vector<ModupJob> mj;
    for (size_t i = 0; i < n; i++) {
        u64 *cinv = c->scratch(L);
        u64 *D = c->scratch((size_t)beta * W);
        out[i].D = D; out[i].L = L; out[i].beta = beta; out[i].src = c1[i];
        for (int l = 0; l < L; l++) inv.push_back({c1[i] + (size_t)l * HEC_N, cinv + (size_t)l * HEC_N, l, 0});
        for (int d = 0; d < beta; d++) {
            int st = d * alpha, nd = std::min(c->xalpha[d], L - st);
            for (int t = 0; t < W; t++) {
                u64 *dst = D + ((size_t)d * W + t) * HEC_N;
                int mod = t < L ? c->modQ(t) : c->modP(t - L);
                if (t < L && t >= st && t < st + nd) continue; // in-digit limb: the inner product reads c1's NTT form in place
                if (nd == 1) { // copy path: the single limb reduced modulo the target in the transform's prologue
                    fwd.push_back({cinv + (size_t)st * HEC_N, dst, mod, HEC_LJ_PRO, nullptr, nullptr, 0, 0});
                    continue;
                }
                mj.push_back(modup_job(c, c->dec[d][nd], cinv + (size_t)st * HEC_N, HEC_N, mod, dst));
                fwd.push_back({dst, dst, mod, 0});
            }
        }
    }
    if ((rc = hec_launch_ntt(c, inv, true))) return rc;
    if (!mj.empty() && (rc = launch_modup(c, mj))) return rc;
    return hec_launch_ntt(c, fwd, false);
}
// KeyswitchHoisted (L:rlwe/keyswitch.go:234-304) for item i: inner product of decomposition dc[i] (or the
// shared dc[0]) with key[i], then mod-down.  d0[i], d1[i]: [L][N] outputs.  Scratch: (2 nP + 2 L) limbs per item.
// launch k_dot for a list of (a pointers, b pointers or none, out, modulus) sums; the pointer lists are
// staged into one stream-ordered device buffer
struct DotSpec { std::vector<const u64 *> a, b; u64 *out; int mod; };
static int launch_dot(hec_ctx *c, const std::vector<DotSpec> &specs) {
    size_t np = 0;
    for (auto &sp : specs) np += sp.a.size() + sp.b.size();
    size_t bytes = np * sizeof(u64 *) + specs.size() * sizeof(DotJob);
    char *dbuf = nullptr;
    std::vector<char> h(bytes);
    const u64 **hp = reinterpret_cast<const u64 **>(h.data());
    DotJob *hj = reinterpret_cast<DotJob *>(h.data() + np * sizeof(u64 *));
    size_t off = 0;
    for (size_t i = 0; i < specs.size(); i++) {
        const DotSpec &sp = specs[i];
        hj[i].a_off = (long long)off;
        for (auto x : sp.a) hp[off++] = x;
        hj[i].b_off = sp.b.empty() ? -1 : (long long)off;
        for (auto x : sp.b) hp[off++] = x;
        hj[i].out = sp.out; hj[i].mod = sp.mod; hj[i].T = (int)sp.a.size();
    }
    int rc = stage_cached(c, h, &dbuf);
    if (rc) return rc;
    // HEC_DOT_BULK=1: operand tiles staged by bulk asynchronous copies + mbarriers (k_dot_bulk); 0: per-thread loads (k_dot)
    static const int bulk = getenv("HEC_DOT_BULK") ? atoi(getenv("HEC_DOT_BULK")) : HEC_DOT_BULK_DEFAULT;
    static bool attr_set = false;
    if (bulk && !attr_set) {
        HEC_CUDA(c, cudaFuncSetAttribute(k_dot_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_DOT_BULK_SMEM));
        attr_set = true;
    }
    for (size_t off = 0; off < specs.size(); off += 65535) { // grid.y limit
        unsigned ny = (unsigned)std::min<size_t>(65535, specs.size() - off);
        const DotJob *dj = reinterpret_cast<const DotJob *>(dbuf + np * sizeof(u64 *)) + off;
        if (bulk)
            launch_k_smem(c, k_dot_bulk, dim3(HEC_N / HEC_DOT_TILE, ny), dim3(256), HEC_DOT_BULK_SMEM, dj,
                          reinterpret_cast<const u64 *const *>(dbuf), c->dmods);
        else
            launch_k(c, k_dot, dim3(32, ny), dim3(256), dj, reinterpret_cast<const u64 *const *>(dbuf), c->dmods);
        c->launches += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->fail(HEC_E_CUDA, std::string("k_dot: ") + cudaGetErrorString(e));
    return HEC_OK;
}

// SwitchKeysInPlaceNoModDown / KeyswitchHoistedNoModDown (L:rlwe/keyswitch.go:149-304): the inner product of
// decomposition dc[i] (or the shared dc[0]) with key[i], left in Q||P: accQ[2i+p] [L][N], accP[2i+p] [nP][N], p = 0,1
static int ks_mac_many(hec_ctx *c, int level, const std::vector<Decomp> &dc, const std::vector<const SwKey *> &key,
                       const std::vector<u64 *> &accQ, const std::vector<u64 *> &accP) {
    int L = level + 1, nP = c->nP, W = L + nP, rc;
    size_t n = key.size();
    int beta = dc[0].beta;
    for (size_t i = 0; i < n; i++)
        if (key[i]->Lk < L || key[i]->ndig < beta) return c->fail(HEC_E_NOKEY, "switching key slice does not cover this level");
    // one pass over all digits (k_dot): acc = sum_d digit_d * key_d, instead of one read-modify-write per digit
    std::vector<DotSpec> specs;
    for (size_t i = 0; i < n; i++) {
        const Decomp &D = dc.size() == 1 ? dc[0] : dc[i];
        int kl = key[i]->Lk + nP;
        for (int p = 0; p < 2; p++)
            for (int t = 0; t < W; t++) {
                DotSpec sp;
                int kt = t < L ? t : key[i]->Lk + (t - L);
                for (int d = 0; d < beta; d++) {
                    int st = d * c->alpha, nd = std::min(c->xalpha[d], L - st);
                    bool own = t < L && t >= st && t < st + nd; // limb of the digit itself: no lifted copy exists
                    sp.a.push_back(own ? D.src + (size_t)t * HEC_N : D.D + ((size_t)d * W + t) * HEC_N);
                    sp.b.push_back(key[i]->buf + ((size_t)(d * 2 + p) * kl + kt) * HEC_N);
                }
                sp.out = t < L ? accQ[2 * i + p] + (size_t)t * HEC_N : accP[2 * i + p] + (size_t)(t - L) * HEC_N;
                sp.mod = t < L ? c->modQ(t) : c->modP(t - L);
                specs.push_back(sp);
            }
    }
    (void)rc;
    return launch_dot(c, specs);
}
// add0 / add1 (optional): per item a polynomial added to d0 / d1, fused into the mod-down's last pass; acc_out
// (optional, with both): the sums go to add0 / add1 themselves (out += SwitchKeys(.), the relinearisation)
// out0 / out1 + gal (optional): the results go there, stored through the automorphism of gal[i] (a rotation's tail)
static int keyswitch_many(hec_ctx *c, int level, const std::vector<Decomp> &dc, const std::vector<const SwKey *> &key,
                          const std::vector<u64 *> &d0, const std::vector<u64 *> &d1, const std::vector<const u64 *> &add0 = {},
                          const std::vector<const u64 *> &add1 = {}, bool acc_out = false, const std::vector<u64 *> &out0 = {},
                          const std::vector<u64 *> &out1 = {}, const std::vector<u64> &gal = {}) {
    int nP = c->nP, rc;
    std::vector<u64 *> accQ, accP, outQ;
    std::vector<const u64 *> addQ;
    std::vector<u64> galQ;
    for (size_t i = 0; i < key.size(); i++) {
        u64 *ap = c->scratch(2 * (size_t)nP);
        accQ.push_back(d0[i]); accP.push_back(ap);
        accQ.push_back(d1[i]); accP.push_back(ap + (size_t)nP * HEC_N);
        addQ.push_back(i < add0.size() ? add0[i] : nullptr);
        addQ.push_back(i < add1.size() ? add1[i] : nullptr);
        if (acc_out) { outQ.push_back(const_cast<u64 *>(add0[i])); outQ.push_back(const_cast<u64 *>(add1[i])); }
        else if (!out0.empty()) { outQ.push_back(out0[i]); outQ.push_back(out1[i]); galQ.push_back(gal[i]); galQ.push_back(gal[i]); }
    }
    if ((rc = ks_mac_many(c, level, dc, key, accQ, accP))) return rc;
    return moddown_many(c, level, accQ, accP, addQ, outQ, galQ);
}
static size_t ks_limbs(const hec_ctx *c, int level) { return 2 * c->nP + 2 * (size_t)(level + 1); } // per item, after decompose

// permuteNTT tail (L:ckks/evaluator.go:1575-1597) for every item: d0 += c0, then PermuteNTTWithIndexLvl x2
static int finish_rotations(hec_ctx *c, int level, const std::vector<const hec_ct *> &ct, const std::vector<u64> &galEl,
                            const std::vector<u64 *> &d0, const std::vector<u64 *> &d1, const std::vector<hec_ct *> &out) {
    int L = level + 1, rc;
    std::vector<EwJob> perm; // d0 already carries + c0 (added in the mod-down's last pass, keyswitch_many add0)
    for (size_t i = 0; i < ct.size(); i++)
        for (int l = 0; l < L; l++) {
            perm.push_back(ewjob(d0[i] + (size_t)l * HEC_N, nullptr, out[i]->limb(0, l), l, 0, (u32)galEl[i]));
            perm.push_back(ewjob(d1[i] + (size_t)l * HEC_N, nullptr, out[i]->limb(1, l), l, 0, (u32)galEl[i]));
        }
    if ((rc = launch_ew<EW_PERMUTE>(c, perm))) return rc;
    for (size_t i = 0; i < ct.size(); i++) { out[i]->level = level; out[i]->scale = ct[i]->scale; }
    return HEC_OK;
}

// n independent rotations out[i] = sigma_{galEl[i]}(ct[i]) at a common level; `hoisted`: all ct[i] are the
// same ciphertext and share one decomposition (RotateHoisted, L:ckks/linear_transform.go:10-27).
int hec_rotate_many(hec_ctx *c, const std::vector<const hec_ct *> &ct, const std::vector<u64> &galEl,
                    const std::vector<hec_ct *> &out, bool hoisted) {
    size_t n = ct.size();
    if (n == 0) return HEC_OK;
    std::vector<const SwKey *> keys(n);
    int level = ct[0]->level;
    for (size_t i = 0; i < n; i++) {
        auto it = c->keys.find(galEl[i]);
        if (it == c->keys.end()) return c->fail(HEC_E_NOKEY, "rotation key for galEl " + std::to_string(galEl[i]) + " missing");
        keys[i] = &it->second;
        level = std::min(level, std::min(ct[i]->level, out[i]->alloc - 1));
    }
    int L = level + 1, rc;
    size_t ndec = hoisted ? 1 : n;
    if ((rc = reserve(c, ndec * decomp_limbs(c, level) + n * (ks_limbs(c, level) + 2 * (size_t)L)))) return rc;
    std::vector<u64 *> d0(n), d1(n);
    for (size_t i = 0; i < n; i++) { d0[i] = c->scratch(L); d1[i] = c->scratch(L); }
    std::vector<const u64 *> c1;
    for (size_t i = 0; i < ndec; i++) c1.push_back(ct[i]->limb(1, 0));
    std::vector<Decomp> dc;
    if ((rc = decompose_many(c, level, c1, dc))) return rc;
    std::vector<const u64 *> c0(n);
    bool alias = false; // RotateGal(ct, g, ct): the input's c0 is still being read while outputs land -> keep the gather pass
    for (size_t i = 0; i < n; i++) {
        c0[i] = ct[i]->limb(0, 0); // limbs of one polynomial are contiguous ([alloc][N])
        for (size_t k = 0; k < n; k++) alias = alias || out[i]->buf == ct[k]->buf;
    }
    static const int scatter = getenv("HEC_ROT_SCATTER") ? atoi(getenv("HEC_ROT_SCATTER")) : 1;
    if (scatter && !alias) {
        std::vector<u64 *> o0(n), o1(n);
        for (size_t i = 0; i < n; i++) { o0[i] = out[i]->limb(0, 0); o1[i] = out[i]->limb(1, 0); }
        if ((rc = keyswitch_many(c, level, dc, keys, d0, d1, c0, {}, false, o0, o1, galEl))) return rc;
        for (size_t i = 0; i < n; i++) { out[i]->level = level; out[i]->scale = ct[i]->scale; }
        return HEC_OK;
    }
    if ((rc = keyswitch_many(c, level, dc, keys, d0, d1, c0))) return rc;
    return finish_rotations(c, level, ct, galEl, d0, d1, out);
}

// MulRelinNew(ct, ct) (mulRelin ciphertext branch, L:ckks/evaluator.go:1398-1444): tensor product, then
// (c0, c1) += SwitchKeys(c2, rlk).  The relinearisation key is stored under the reserved id HEC_RLK_ID.
extern "C" int hec_rlk_upload(hec_ctx *c, int max_level, const uint64_t *const *limbs) {
    return hec_swk_upload(c, HEC_RLK_ID, max_level, limbs);
}
// a batch of independent products out[m] = MulRelinNew(a[m], b[m]) at a common level: one tensor launch,
// one batched decomposition / inner product / mod-down, one add
int hec_mul_relin_many(hec_ctx *c, const std::vector<const hec_ct *> &a, const std::vector<const hec_ct *> &b, std::vector<hec_ct *> &out) {
    auto it = c->keys.find(HEC_RLK_ID);
    if (it == c->keys.end()) return c->fail(HEC_E_NOKEY, "relinearisation key missing");
    size_t n = a.size();
    int level = std::min(a[0]->level, b[0]->level), L = level + 1, rc;
    for (size_t m = 0; m < n; m++)
        if (std::min(a[m]->level, b[m]->level) != level) return c->fail(HEC_E_INVAL, "batched MulRelin needs a common level");
    out.assign(n, nullptr);
    auto bail = [&](int e) { for (hec_ct *o : out) hec_ct_free(c, o); out.assign(n, nullptr); return e; };
    for (size_t m = 0; m < n; m++)
        if ((rc = hec_ct_alloc(c, level, a[m]->scale * b[m]->scale, &out[m]))) return bail(rc);
    if ((rc = reserve(c, n * (decomp_limbs(c, level) + ks_limbs(c, level) + 3 * (size_t)L)))) return bail(rc);
    std::vector<u64 *> c2(n), d0(n), d1(n);
    std::vector<TensorJob> tj;
    for (size_t m = 0; m < n; m++) {
        c2[m] = c->scratch(L); d0[m] = c->scratch(L); d1[m] = c->scratch(L);
        for (int i = 0; i < L; i++)
            tj.push_back({a[m]->limb(0, i), a[m]->limb(1, i), b[m]->limb(0, i), b[m]->limb(1, i), out[m]->limb(0, i), out[m]->limb(1, i),
                          c2[m] + (size_t)i * HEC_N, mform(c->hm[i].rmod, c->q(i)), i});
    }
    for (size_t off = 0; off < tj.size(); off += HEC_TNJOBS) {
        int k = (int)std::min<size_t>(HEC_TNJOBS, tj.size() - off);
        TensorJobs J;
        for (int i = 0; i < k; i++) J.j[i] = tj[off + i];
        launch_k(c, k_tensor, dim3(32, k), dim3(256), J, c->dmods);
        c->launches += 1;
    }
    if ((rc = check_launch(c, "tensor"))) return bail(rc);
    std::vector<const u64 *> src(c2.begin(), c2.end());
    std::vector<Decomp> dc;
    if ((rc = decompose_many(c, level, src, dc))) return bail(rc);
    std::vector<const SwKey *> keys(n, &it->second);
    // (c0, c1) += SwitchKeys(c2, rlk): the sums are formed in the last pass of the mod-down, straight into the output
    std::vector<const u64 *> o0(n), o1(n);
    for (size_t m = 0; m < n; m++) { o0[m] = out[m]->limb(0, 0); o1[m] = out[m]->limb(1, 0); }
    if ((rc = keyswitch_many(c, level, dc, keys, d0, d1, o0, o1, true))) return bail(rc);
    return HEC_OK;
}
// Rescale(MulRelinNew(a, b), min_scale) for a batch, the first division fused into the relinearisation's mod-down.
// Both end with a forward transform per limb: ModDownSplitNTTPQ with NTT(x_l), x_l the basis-extended P part, and
// divRoundByLastModulusNTT with NTT(r_l), r_l the centred remainder of the dropped limb.  On the kept limbs
//   ((accQ_l - NTT(x_l)) / P + d_l - NTT(r_l)) / q_T  =  (NTT(x_l + P r_l) - accQ_l) * (-(P q_T)^-1) + d_l / q_T
// (d = the tensor product's c0 / c1, T the dropped limb): ONE transform of x_l + P r_l instead of two.  The dropped limb
// itself is completed first (its mod-down as usual), transformed back and centred to give r.  Exact arithmetic modulo
// every q_l, so the result is bit for bit that of the two operations in sequence (L:ckks/evaluator.go:1398-1444,
// 1291-1325; L:ring/ring_scaling.go:442-513).  Further divisions, if the scale asks for them, run unfused.
// addend (optional): Rescale(Add(MulRelinNew(a, b), addend)) -- the sum is formed on the tensor product's c0 / c1 before the
// mod-down adds the key-switched part (exact, so the order is free).  The caller guarantees what hec_relin_rescale_fusable
// checks: levels and scales uniform over the batch, addend at the product's level or above with a scale the Add would
// not have to match by an integer factor.
static bool hec_relin_rescale_fusable(hec_ctx *c, const std::vector<const hec_ct *> &a, const std::vector<const hec_ct *> &b, double min_scale,
                                      const std::vector<const hec_ct *> *addend) {
    static const int fuse = getenv("HEC_RELIN_RESCALE") ? atoi(getenv("HEC_RELIN_RESCALE")) : 1;
    const size_t n = a.size();
    const int level = std::min(a[0]->level, b[0]->level);
    double s0 = a[0]->scale * b[0]->scale;
    bool same = fuse && level > 0;
    for (size_t m = 0; m < n && same; m++) same = std::min(a[m]->level, b[m]->level) == level && a[m]->scale * b[m]->scale == s0;
    if (same && addend) {
        const double sb = (*addend)[0]->scale;
        for (size_t m = 0; m < n && same; m++) same = (*addend)[m]->level >= level && (*addend)[m]->scale == sb;
        if (same && ((s0 > sb && floor(s0 / sb) > 1) || (sb > s0 && floor(sb / s0) > 1))) same = false; // evaluateInPlace would rescale one side
        s0 = std::max(s0, sb);
    }
    return same && s0 / (double)c->q(level) >= min_scale / 2;
}
int hec_mul_relin_rescale_many(hec_ctx *c, const std::vector<const hec_ct *> &a, const std::vector<const hec_ct *> &b, double min_scale,
                               std::vector<hec_ct *> &out, const std::vector<const hec_ct *> *addend = nullptr) {
    size_t n = a.size();
    int level = std::min(a[0]->level, b[0]->level), L = level + 1, rc;
    double s0 = a[0]->scale * b[0]->scale;
    if (!hec_relin_rescale_fusable(c, a, b, min_scale, addend)) {
        if (addend) return c->fail(HEC_E_INVAL, "fused MulRelin + Add + Rescale: operands do not qualify");
        if ((rc = hec_mul_relin_many(c, a, b, out))) return rc;
        rc = hec_rescale_many(c, out, min_scale);
        if (rc) { for (hec_ct *o : out) hec_ct_free(c, o); out.assign(n, nullptr); }
        return rc;
    }
    auto it = c->keys.find(HEC_RLK_ID);
    if (it == c->keys.end()) return c->fail(HEC_E_NOKEY, "relinearisation key missing");
    out.assign(n, nullptr);
    auto bail = [&](int e) { for (hec_ct *o : out) hec_ct_free(c, o); out.assign(n, nullptr); return e; };
    for (size_t m = 0; m < n; m++)
        if ((rc = hec_ct_alloc(c, level, s0, &out[m]))) return bail(rc);
    const int nP = c->nP, T = level;
    if ((rc = reserve(c, n * (decomp_limbs(c, level) + ks_limbs(c, level) + 3 * (size_t)L + 2 * (size_t)L + 4)))) return bail(rc);
    std::vector<u64 *> c2(n), d0(n), d1(n);
    std::vector<TensorJob> tj;
    for (size_t m = 0; m < n; m++) {
        c2[m] = c->scratch(L); d0[m] = c->scratch(L); d1[m] = c->scratch(L);
        for (int i = 0; i < L; i++)
            tj.push_back({a[m]->limb(0, i), a[m]->limb(1, i), b[m]->limb(0, i), b[m]->limb(1, i), out[m]->limb(0, i), out[m]->limb(1, i),
                          c2[m] + (size_t)i * HEC_N, mform(c->hm[i].rmod, c->q(i)), i});
    }
    for (size_t off = 0; off < tj.size(); off += HEC_TNJOBS) {
        int k = (int)std::min<size_t>(HEC_TNJOBS, tj.size() - off);
        TensorJobs J;
        for (int i = 0; i < k; i++) J.j[i] = tj[off + i];
        launch_k(c, k_tensor, dim3(32, k), dim3(256), J, c->dmods);
        c->launches += 1;
    }
    if ((rc = check_launch(c, "tensor"))) return bail(rc);
    if (addend) { // Add(prod, addend): on the tensor product's c0 / c1, before the key-switched part joins them
        std::vector<EwJob> aj;
        for (size_t m = 0; m < n; m++)
            for (int p = 0; p < 2; p++)
                for (int i = 0; i < L; i++) aj.push_back(ewjob(out[m]->limb(p, i), (*addend)[m]->limb(p, i), out[m]->limb(p, i), i));
        if ((rc = launch_ew<EW_ADD>(c, aj))) return bail(rc);
        s0 = std::max(s0, (*addend)[0]->scale);
    }
    std::vector<const u64 *> src(c2.begin(), c2.end());
    std::vector<Decomp> dc;
    if ((rc = decompose_many(c, level, src, dc))) return bail(rc);
    std::vector<const SwKey *> keys(n, &it->second);
    // inner product with the key, left in Q||P: item 2m + p = polynomial p of ciphertext m
    std::vector<u64 *> accQ, accP;
    for (size_t m = 0; m < n; m++) {
        u64 *ap = c->scratch(2 * (size_t)nP);
        accQ.push_back(d0[m]); accP.push_back(ap);
        accQ.push_back(d1[m]); accP.push_back(ap + (size_t)nP * HEC_N);
    }
    if ((rc = ks_mac_many(c, level, dc, keys, accQ, accP))) return bail(rc);
    // mod-down, first half: InvNTT on P, exact basis extension P -> every Q limb (coefficient domain)
    std::vector<LimbJob> nj;
    for (size_t i = 0; i < 2 * n; i++)
        for (int j = 0; j < nP; j++) nj.push_back({accP[i] + (size_t)j * HEC_N, accP[i] + (size_t)j * HEC_N, c->modP(j), 0});
    if ((rc = hec_launch_ntt(c, nj, true))) return bail(rc);
    u64 *x = c->scratch(2 * n * L), *vt = c->scratch(2 * n);
    std::vector<ModupJob> mj;
    for (size_t i = 0; i < 2 * n; i++)
        for (int l = 0; l < L; l++) mj.push_back(modup_job(c, c->pq, accP[i], HEC_N, c->modQ(l), x + (i * L + l) * HEC_N));
    if ((rc = launch_modup(c, mj))) return bail(rc);
    // the limb that is dropped: complete its mod-down (+ the tensor product's c0 / c1), transform back, centre
    nj.clear();
    for (size_t i = 0; i < 2 * n; i++) {
        u64 *xi = x + (i * L + T) * HEC_N;
        nj.push_back({xi, vt + i * HEC_N, T, HEC_LJ_EPI | HEC_LJ_ADD, xi, accQ[i] + (size_t)T * HEC_N, c->negpinv[T], 0,
                      out[i / 2]->limb((int)(i & 1), T), 0});
    }
    if ((rc = hec_launch_ntt(c, nj, false))) return bail(rc);
    nj.clear();
    const u64 qT = c->q(T), half = (qT - 1) >> 1;
    std::vector<EwJob> ej;
    for (size_t i = 0; i < 2 * n; i++) {
        nj.push_back({vt + i * HEC_N, vt + i * HEC_N, T, 0});
        ej.push_back(ewjob(vt + i * HEC_N, nullptr, vt + i * HEC_N, T, half));
    }
    if ((rc = hec_launch_ntt(c, nj, true))) return bail(rc);
    if ((rc = launch_ew<EW_CENTER>(c, ej))) return bail(rc);
    // the kept limbs: one transform of x_l + P r_l, both combines in its epilogue
    nj.clear();
    for (size_t i = 0; i < 2 * n; i++)
        for (int l = 0; l < T; l++) {
            const u64 ql = c->q(l);
            u64 Pm = 1;
            for (int j = 0; j < nP; j++) Pm = mulmod(Pm, c->q(c->modP(j)) % ql, ql);
            const u64 pq_inv = invmod(mulmod(Pm, qT % ql, ql), ql);
            u64 *xi = x + (i * L + l) * HEC_N;
            u64 *dl = out[i / 2]->limb((int)(i & 1), l); // d_l in, the result out (each thread reads its element before it writes it)
            LimbJob j = {xi, dl, l, HEC_LJ_PRO2 | HEC_LJ_EPI | HEC_LJ_ADDS, xi, accQ[i] + (size_t)l * HEC_N, mform(ql - pq_inv, ql),
                         ql - half % ql, dl, 0, vt + i * HEC_N, mform(Pm, ql), mform(invmod(qT % ql, ql), ql)};
            nj.push_back(j);
        }
    if ((rc = hec_launch_ntt(c, nj, false))) return bail(rc);
    for (size_t m = 0; m < n; m++) { out[m]->level = level - 1; out[m]->scale = s0 / (double)qT; }
    if ((rc = hec_rescale_many(c, out, min_scale))) return bail(rc);
    return HEC_OK;
}
extern "C" int hec_mul_relin_new(hec_ctx *c, const hec_ct *a, const hec_ct *b, hec_ct **out) {
    if (!c || !a || !b || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    std::vector<hec_ct *> o;
    int rc = hec_mul_relin_many(c, {a}, {b}, o);
    if (rc) return rc;
    *out = o[0];
    return HEC_OK;
}

extern "C" int hec_rotate_gal(hec_ctx *c, const hec_ct *ct, uint64_t galEl, hec_ct *out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    return hec_rotate_many(c, {ct}, {galEl}, {out}, false);
}

extern "C" uint64_t hec_galois_for_rotation(const hec_ctx *c, int k) {
    (void)c;
    // GaloisElementForColumnRotationBy (L:rlwe/params.go:308-312): 5^(k & (2N-1)) mod 2N
    u64 mask = 2ull * HEC_N - 1, e = (u64)(int64_t)k & mask, g = 1, b = 5;
    for (; e; e >>= 1) { if (e & 1) g = (g * b) & mask; b = (b * b) & mask; }
    return g;
}
extern "C" int hec_rotate_new(hec_ctx *c, const hec_ct *ct, int k, hec_ct **out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *o = nullptr;
    int rc = hec_ct_alloc(c, ct->level, ct->scale, &o);
    if (rc) return rc;
    if ((rc = hec_rotate_gal(c, ct, hec_galois_for_rotation(c, k), o))) { hec_ct_free(c, o); return rc; }
    *out = o;
    return HEC_OK;
}
// RotateHoisted (L:ckks/linear_transform.go:10-27): one DecomposeNTT shared by all rotations; the inner
// products, mod-downs and permutations of all rotations are batched into one launch sequence
extern "C" int hec_rotate_hoisted(hec_ctx *c, const hec_ct *ct, const int *rots, int n, hec_ct **outs) {
    if (!c || !ct || !rots || !outs || n < 0) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    int rc;
    for (int r = 0; r < n; r++) {
        outs[r] = nullptr;
        if (rots[r] != 0 && !c->keys.count(hec_galois_for_rotation(c, rots[r])))
            return c->fail(HEC_E_NOKEY, "rotation key for rotation " + std::to_string(rots[r]) + " missing");
    }
    std::vector<const hec_ct *> in;
    std::vector<u64> gal;
    std::vector<hec_ct *> out;
    auto fail = [&](int code) { // no half-filled output list
        for (int r = 0; r < n; r++) if (outs[r]) { hec_ct_free(c, outs[r]); outs[r] = nullptr; }
        return code;
    };
    for (int r = 0; r < n; r++) {
        if (rots[r] == 0) { if ((rc = hec_ct_copy_new(c, ct, &outs[r]))) return fail(rc); continue; } // cOut[0] = ctIn.CopyNew()
        if ((rc = hec_ct_alloc(c, ct->level, ct->scale, &outs[r]))) return fail(rc);
        in.push_back(ct); gal.push_back(hec_galois_for_rotation(c, rots[r])); out.push_back(outs[r]);
    }
    if ((rc = hec_rotate_many(c, in, gal, out, true))) return fail(rc);
    return HEC_OK;
}

// =========================================================================================
// ring-level test entry points (host in / host out)
// =========================================================================================
extern "C" int hec_ntt(hec_ctx *c, int ring, int limb, const uint64_t *in, uint64_t *out, int inverse) {
    if (!c || !in || !out || ring < 0 || ring > 1 || limb < 0 || limb >= (ring ? c->nP : c->nQ)) return c ? c->fail(HEC_E_INVAL, "ntt args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    int rc = reserve(c, 1);
    if (rc) return rc;
    u64 *d = c->scratch(1);
    HEC_CUDA(c, cudaMemcpyAsync(d, in, HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    std::vector<LimbJob> nj = {{d, d, ring ? c->modP(limb) : c->modQ(limb), 0}};
    if ((rc = hec_launch_ntt(c, nj, inverse != 0))) return rc;
    HEC_CUDA(c, cudaMemcpyAsync(out, d, HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}
extern "C" int hec_keyswitch(hec_ctx *c, int level, const uint64_t *const *c1, uint64_t galEl, uint64_t *const *d0, uint64_t *const *d1) {
    if (!c || !c1 || !d0 || !d1 || level < 0 || level >= c->nQ) return c ? c->fail(HEC_E_INVAL, "keyswitch args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    auto it = c->keys.find(galEl);
    if (it == c->keys.end()) return c->fail(HEC_E_NOKEY, "rotation key missing");
    int L = level + 1, rc;
    if ((rc = reserve(c, decomp_limbs(c, level) + ks_limbs(c, level) + 3 * (size_t)L))) return rc;
    u64 *x = c->scratch(L), *a0 = c->scratch(L), *a1 = c->scratch(L);
    for (int i = 0; i < L; i++) HEC_CUDA(c, cudaMemcpyAsync(x + (size_t)i * HEC_N, c1[i], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    std::vector<Decomp> dc;
    if ((rc = decompose_many(c, level, {x}, dc))) return rc;
    if ((rc = keyswitch_many(c, level, dc, {&it->second}, {a0}, {a1}))) return rc;
    for (int i = 0; i < L; i++) {
        HEC_CUDA(c, cudaMemcpyAsync(d0[i], a0 + (size_t)i * HEC_N, HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
        HEC_CUDA(c, cudaMemcpyAsync(d1[i], a1 + (size_t)i * HEC_N, HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    }
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}
extern "C" int hec_moddown(hec_ctx *c, int level, const uint64_t *const *accQ, const uint64_t *const *accP, uint64_t *const *out) {
    if (!c || !accQ || !accP || !out || c->nP == 0 || level < 0 || level >= c->nQ) return c ? c->fail(HEC_E_INVAL, "moddown args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    int L = level + 1, rc;
    if ((rc = reserve(c, 3 * (size_t)L + c->nP))) return rc;
    u64 *aq = c->scratch(L), *ap = c->scratch(c->nP);
    for (int i = 0; i < L; i++) HEC_CUDA(c, cudaMemcpyAsync(aq + (size_t)i * HEC_N, accQ[i], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    for (int j = 0; j < c->nP; j++) HEC_CUDA(c, cudaMemcpyAsync(ap + (size_t)j * HEC_N, accP[j], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    if ((rc = moddown_many(c, level, {aq}, {ap}))) return rc;
    for (int i = 0; i < L; i++) HEC_CUDA(c, cudaMemcpyAsync(out[i], aq + (size_t)i * HEC_N, HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}
