// hec_host.cuh -- host-side ring constants, context and handle types of libhec.
// Tables follow the reference's Lattigo fork: primitive-root choice and psi tables
// L:ring/ring.go:118-200, L:ring/utils.go:69-90; rescale / mod-down / mod-up constants
// L:ring/ring.go:63-117, L:ring/ring_basis_extension.go:43-145,482-538.
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <unordered_map>
#include <string>
#include <vector>
#include "hec_kernels.cuh"
#include "../../include/hec.h"

typedef unsigned __int128 u128;

namespace hec {

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
inline u64 powmod(u64 b, u64 e, u64 q) {
    u64 r = 1 % q;
    b %= q;
    for (; e; e >>= 1) { if (e & 1) r = mulmod(r, b, q); b = mulmod(b, b, q); }
    return r;
}
inline u64 invmod(u64 a, u64 q) { return powmod(a % q, q - 2, q); } // q prime
inline u64 mform(u64 a, u64 q) { return (u64)((((u128)(a % q)) << 64) % q); }
inline u32 bitrev16(u32 x) {
    u32 r = 0;
    for (int i = 0; i < 16; i++) r |= ((x >> i) & 1u) << (15 - i);
    return r;
}

struct HostMod {
    u64 q = 0, qinv = 0, rmod = 0, gen = 0;
    u64 ninv_w = 0, ninv_s = 0;          // N^-1 mod q and floor(N^-1 * 2^64 / q)
    u64 psi_half_mont = 0;               // MForm(psi^(N/2)) = NttPsi[1] as Lattigo stores it (MultByi / DivByi)
    std::vector<ulonglong2> psi, psi_inv; // (w, floor(w * 2^64 / q)) at index brev(j), w = psi^(+-j)
};
void build_mod(HostMod &m, u64 q);

// tables of one exact basis extension  S = {s_0..s_{n-1}}  ->  any target modulus
struct ModupTab {
    int n = 0;
    std::vector<int> smod;               // source modulus indices
    std::vector<u64> qib;                // [n]
    std::vector<std::vector<u64>> qisp;  // [target][n]
    std::vector<std::vector<u64>> qpjinv; // [target][n+1]
};

// device buffer of one switching key.  Plans hold a reference: dropping or replacing a key that a live plan's graph
// still reads only marks it dead; the memory goes when the last such plan does.
struct KeyBuf {
    u64 *buf = nullptr;
    int plan_refs = 0;
    bool dead = false;
    uint64_t serial = 0;
};
struct SwKey {
    u64 *buf = nullptr; // [ndig][2][Lk + nP][N], Montgomery (= kb->buf)
    int ndig = 0, Lk = 0;
    KeyBuf *kb = nullptr;
};

} // namespace hec

struct hec_ct {
    u64 *buf = nullptr; // [2][alloc][N]
    int alloc = 0, level = 0;
    double scale = 0;
    bool owned = true;
    u64 *limb(int c, int i) const { return buf + ((size_t)c * alloc + i) * HEC_N; }
};
struct hec_pt {
    u64 *buf = nullptr; // [level+1][N], Montgomery form (pt * R mod q)
    int level = 0;
    double scale = 0;
    uint64_t serial = 0; // unique per upload: identity of the contents for the plan cache
};
struct hec_plan;

struct hec_ctx {
    int device = 0, nQ = 0, nP = 0, alpha = 0, beta_full = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr; // the context's own stream-ordered pool
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<hec::HostMod> hm; // Q then P
    ModC *dmods = nullptr;
    ulonglong2 *dtables = nullptr;
    std::vector<std::vector<u64>> resc; // resc[L][i] = MForm(q_i - q_L^-1 mod q_i)
    std::vector<u64> negpinv;           // q_i - MForm(P^-1 mod q_i)
    hec::ModupTab pq;                   // P -> Q
    std::vector<int> xalpha;
    std::vector<std::vector<hec::ModupTab>> dec; // dec[d][nd]
    std::map<u64, hec::SwKey> keys;
    // scratch arena for the generic ops (stream-ordered reuse)
    u64 *arena = nullptr;
    size_t arena_limbs = 0, arena_top = 0;
    // small job / pointer tables of k_modup2 and k_dot, kept on the device and found again by content: the same
    // operation on the same buffers (a layer repeated per image, a benchmark loop) then launches without a copy
    // The device memory is a ring of slabs: when the ring comes round to a slab, the tables it held are dropped from
    // the cache (and the event recorded when the ring left it is waited for -- long past by then), so a long chain of
    // operations on ever new addresses (a network, image after image) keeps the most recent tables and never stalls
    // to start over.
    struct StagedTab { std::vector<char> host; char *dev = nullptr; };
    std::unordered_map<uint64_t, std::vector<StagedTab>> staged;
    std::vector<char *> stage_slabs;                 // ring of slabs the tables are carved from (bump allocation inside a slab)
    std::vector<cudaEvent_t> stage_events;           // recorded on the stream when the ring leaves a slab
    std::vector<std::vector<uint64_t>> stage_keys;   // hash keys of the tables living in each slab
    std::vector<char *> stage_big;                   // tables larger than a slab: blocks of their own, freed when the ring wraps
    size_t stage_idx = 0;                            // slab being filled
    size_t stage_slab_bytes = (size_t)16 << 20;      // slab size and number of slabs in the ring: 1 GiB in all (a network
    size_t stage_ring = 64;                          // layer chain stages thousands of distinct tables per image)
    char *stage_cur = nullptr;
    size_t stage_slab_top = 0;
    uint64_t launches = 0;
    uint64_t next_serial = 1;
    std::vector<hec_plan *> plan_cache; // plans hec_conv_then_pack built, most recently used first
    std::string err;

    int modQ(int i) const { return i; }
    int modP(int j) const { return nQ + j; }
    u64 q(int mod) const { return hm[mod].q; }
    u64 *scratch(size_t limbs);
    void scratch_reset() { arena_top = 0; }
    int fail(int code, const std::string &msg) { err = msg; return code; }
};

#define HEC_CUDA(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return (ctx)->fail(HEC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)
