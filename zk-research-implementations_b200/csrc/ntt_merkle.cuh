// ntt_merkle.cuh -- SURVEY section 8 (f-4), the last "next" row: the radix-2 number-theoretic transform of
// fft/src/fft.rs:6-60 and the Keccak Merkle tree of merkle_tree/src/merkle_tree.rs:31-214 on the device.
// (included by kernels.cuh inside namespace zkb; field-generic: the root of unity comes from the host)
#pragma once

// ================================================================== Keccak-256 of one or two field elements
// merkle_tree.rs:201-214: compute_hash(x) = Keccak256(fq_vec_to_bytes([x])), hash_pair(l, r) = Keccak256(bytes(l) || bytes(r)),
// both mapped back with F::from_le_bytes_mod_order.  sha3 0.10.8 Keccak256 = rate 136, ORIGINAL 0x01 padding.  32 or 64
// message bytes fit one block: one permutation per hash, one hash per thread (25 lanes in registers).
// (round constants: KECCAK_RC of the device transcript, kernels.cuh)
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ void keccak_f1600_thread(uint64_t* a) {
#pragma unroll 1
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5], d[5];
#pragma unroll
        for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
        for (int x = 0; x < 5; ++x) d[x] = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; ++i) a[i] ^= d[i % 5];
        // rho + pi
        uint64_t b[25];
        b[0] = a[0];
        b[10] = rotl64(a[1], 1);   b[20] = rotl64(a[2], 62);  b[5] = rotl64(a[3], 28);   b[15] = rotl64(a[4], 27);
        b[16] = rotl64(a[5], 36);  b[1] = rotl64(a[6], 44);   b[11] = rotl64(a[7], 6);   b[21] = rotl64(a[8], 55);  b[6] = rotl64(a[9], 20);
        b[7] = rotl64(a[10], 3);   b[17] = rotl64(a[11], 10); b[2] = rotl64(a[12], 43);  b[12] = rotl64(a[13], 25); b[22] = rotl64(a[14], 39);
        b[23] = rotl64(a[15], 41); b[8] = rotl64(a[16], 45);  b[18] = rotl64(a[17], 15); b[3] = rotl64(a[18], 21);  b[13] = rotl64(a[19], 8);
        b[14] = rotl64(a[20], 18); b[24] = rotl64(a[21], 2);  b[9] = rotl64(a[22], 61);  b[19] = rotl64(a[23], 56); b[4] = rotl64(a[24], 14);
#pragma unroll
        for (int y = 0; y < 25; y += 5)
#pragma unroll
            for (int x = 0; x < 5; ++x) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        a[0] ^= KECCAK_RC[round];
    }
}
// Montgomery residues in, Montgomery residue out; b == nullptr: compute_hash, else hash_pair
template <class F>
__device__ __forceinline__ Fe merkle_hash(const Fe& a_mont, const Fe* b_mont) {
    typedef Field<F> Fd;
    uint64_t s[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) s[i] = 0;
    const Fe a = Fd::from_mont(a_mont);
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = pack64(a.l[2 * k], a.l[2 * k + 1]);
    if (b_mont) {
        const Fe b = Fd::from_mont(*b_mont);
#pragma unroll
        for (int k = 0; k < 4; ++k) s[4 + k] = pack64(b.l[2 * k], b.l[2 * k + 1]);
        s[8] ^= 0x01ull;
    } else {
        s[4] ^= 0x01ull;
    }
    s[16] ^= 0x8000000000000000ull;  // last byte of the 136-byte rate block
    keccak_f1600_thread(s);
    Fe d;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d.l[2 * k] = lo32(s[k]);
        d.l[2 * k + 1] = hi32(s[k]);
    }
    return Fd::to_mont(Fd::mod_p(d));  // from_le_bytes_mod_order
}
__device__ __forceinline__ Fe ld_aos(const Fe* p, uint64_t i) {
    const uint4* q = reinterpret_cast<const uint4*>(p + i);
    const uint4 x = q[0], y = q[1];
    Fe r;
    r.l[0] = x.x; r.l[1] = x.y; r.l[2] = x.z; r.l[3] = x.w;
    r.l[4] = y.x; r.l[5] = y.y; r.l[6] = y.z; r.l[7] = y.w;
    return r;
}
__device__ __forceinline__ void st_aos(Fe* p, uint64_t i, const Fe& v) {
    uint4* q = reinterpret_cast<uint4*>(p + i);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// MerkleTree::new_with_inputs, :52-65: leaves[i] = compute_hash(input_i), the rest stay F::zero() (NOT hashed)
template <class F>
__global__ void __launch_bounds__(BLOCK) k_merkle_leaves(const Fe* __restrict__ inputs, uint64_t n_inputs, Fe* __restrict__ leaves, uint64_t n_leaves) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n_leaves; i += (uint64_t)gridDim.x * BLOCK)
        st_aos(leaves, i, i < n_inputs ? merkle_hash<F>(ld_aos(inputs, i), nullptr) : Field<F>::zero());
}
// one level, :67-75: next[i] = hash_pair(prev[2 i], prev[2 i + 1])
template <class F>
__global__ void __launch_bounds__(BLOCK) k_merkle_level(const Fe* __restrict__ prev, Fe* __restrict__ next, uint64_t n_next) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n_next; i += (uint64_t)gridDim.x * BLOCK) {
        const Fe r = ld_aos(prev, 2 * i + 1);
        st_aos(next, i, merkle_hash<F>(ld_aos(prev, 2 * i), &r));
    }
}
// The tree in one buffer: level 0 (leaves, 2^depth) then level 1 (2^(depth-1)) .. level depth (the root).
__host__ __device__ __forceinline__ uint64_t merkle_level_off(uint32_t depth, uint32_t level) { return (2ull << depth) - (2ull << (depth - level)); }
// update_leaf + recompute_path (:86-132) and create_proof (:138-183) walk one root path: a single thread.
// mode 0: update (data is the new leaf hash, or hashed first when !is_hash); mode 1: proof of `data` at leaf_id: status = 1 if
// compute_hash(data) is not the stored leaf ("Data does not match the leaf hash"), siblings[level] otherwise.
template <class F>
__global__ void k_merkle_path(Fe* tree, uint32_t depth, uint64_t leaf_id, Fe data, int is_hash, int mode, Fe* siblings, unsigned int* status) {
    if (threadIdx.x || blockIdx.x) return;
    Fe cur = is_hash ? data : merkle_hash<F>(data, nullptr);
    uint64_t index = leaf_id;
    if (mode == 1) {
        const Fe leaf = ld_aos(tree, leaf_id);
        bool same = true;
        for (int k = 0; k < 8; ++k) same &= leaf.l[k] == cur.l[k];
        *status = same ? 0u : 1u;
        if (!same) return;
    } else {
        st_aos(tree, leaf_id, cur);
    }
    for (uint32_t level = 0; level < depth; ++level) {
        const Fe sib = ld_aos(tree + merkle_level_off(depth, level), index ^ 1);
        if (mode == 1) {
            siblings[level] = sib;
        } else {
            cur = (index & 1) ? merkle_hash<F>(sib, &cur) : merkle_hash<F>(cur, &sib);
            st_aos(tree + merkle_level_off(depth, level + 1), index >> 1, cur);
        }
        index >>= 1;
    }
}

// ================================================================== NTT
// fft.rs:6-29: y[j] = sum_i c_i w^(i j), natural order in and out, w = the n-th root of unity ark-ff derives from the field's
// two-adic generator (host side).  Bit-reversed gather, then decimation-in-time stages s = 0 .. log n - 1: for every block of
// 2^(s+1) elements and k < 2^s:  (u, v) = (x[k], w^(k n / 2^(s+1)) x[k + 2^s])  ->  (u + v, u - v).
// Twiddles w^e, e < n / 2, from a two-level table: w^e = hi[e >> lo_bits] * lo[e & mask].
struct NttPows {
    Fe w[32];  // w^(2^i)
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_ntt_twiddles(TabRef lo, uint32_t lo_bits, TabRef hi, uint64_t n_hi, const __grid_constant__ NttPows pw) {
    const uint64_t n_lo = 1ull << lo_bits;
    for (uint64_t t = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; t < n_lo + n_hi; t += (uint64_t)gridDim.x * BLOCK) {
        const bool is_hi = t >= n_lo;
        const uint64_t e = is_hi ? (t - n_lo) << lo_bits : t;
        Fe v = Field<F>::one();
        for (int b = 0; b < 32; ++b)
            if ((e >> b) & 1) v = Field<F>::mul(v, pw.w[b]);
        st_fe(is_hi ? hi : lo, is_hi ? t - n_lo : t, v);
    }
}
struct NttArgs {
    TabRef in, data;   // first pass: in -> data (bit-reversed gather); later passes: data in place
    uint32_t log_n, s0, g;
    TabRef w_lo, w_hi;
    uint32_t lo_bits;
    int do_scale;      // last pass of an inverse transform: multiply by n^-1 (fft.rs:56-58)
    Fe scale;
};
constexpr int NTT_TILE_LOG = 9;  // 512 elements per CTA and pass, one butterfly per thread and stage
template <class F>
__device__ __forceinline__ Fe ntt_twiddle(const NttArgs& a, uint64_t e) {
    const Fe lo = ld_fe(a.w_lo, e & ((1ull << a.lo_bits) - 1));
    const uint64_t h = e >> a.lo_bits;
    return h ? Field<F>::mul(ld_fe(a.w_hi, h), lo) : lo;
}
__device__ __forceinline__ void sm_put(uint4* sm, uint32_t i, const Fe& v) {
    sm[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[512 + i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fe sm_get(const uint4* sm, uint32_t i) {
    const uint4 x = sm[i], y = sm[512 + i];
    Fe r;
    r.l[0] = x.x; r.l[1] = x.y; r.l[2] = x.z; r.l[3] = x.w;
    r.l[4] = y.x; r.l[5] = y.y; r.l[6] = y.z; r.l[7] = y.w;
    return r;
}
// One pass = stages [s0, s0 + g) on tiles of 2^g strided x C = 2^(9 - g) contiguous elements.  s0 == 0: the tile is a
// contiguous block of 2^g outputs gathered from the bit-reversed input positions.
template <class F>
__global__ void __launch_bounds__(BLOCK) k_ntt_pass(const __grid_constant__ NttArgs a) {
    typedef Field<F> Fd;
    __shared__ uint4 sm[1024];
    const uint32_t g = a.g, s0 = a.s0, cbits = s0 ? NTT_TILE_LOG - g : 0, C = 1u << cbits;
    const uint32_t tile_elems = 1u << (g + cbits);
    const uint64_t n = 1ull << a.log_n, tiles = n >> (g + cbits);
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        // global index of tile element (i, c): (hi << (s0 + g)) | (i << s0) | (lo0 + c)
        const uint64_t lo_chunks = s0 ? (1ull << s0) >> cbits : 1;
        const uint64_t hi = tile / lo_chunks, lo0 = (tile % lo_chunks) << cbits;
        const uint64_t base = (hi << (s0 + g)) | lo0;
        for (uint32_t t = threadIdx.x; t < tile_elems; t += BLOCK) {
            const uint32_t i = t >> cbits, c = t & (C - 1);
            const uint64_t j = base | ((uint64_t)i << s0) | c;
            Fe v;
            if (s0 == 0) v = ld_fe(a.in, __brevll(j) >> (64 - a.log_n));
            else v = ld_fe(a.data, j);
            sm_put(sm, t, v);
        }
        __syncthreads();
        for (uint32_t tt = 0; tt < g; ++tt) {
            for (uint32_t bf = threadIdx.x; bf < (tile_elems >> 1); bf += BLOCK) {
                const uint32_t c = bf & (C - 1), ib = bf >> cbits;
                const uint32_t i0 = ((ib >> tt) << (tt + 1)) | (ib & ((1u << tt) - 1)), i1 = i0 | (1u << tt);
                const uint32_t s = s0 + tt;
                const uint64_t k = ((uint64_t)(i0 & ((1u << tt) - 1)) << s0) | (lo0 + c);
                const uint64_t e = k << (a.log_n - s - 1);
                const Fe u = sm_get(sm, (i0 << cbits) | c);
                Fe v = sm_get(sm, (i1 << cbits) | c);
                if (e) v = Fd::mul(v, ntt_twiddle<F>(a, e));
                sm_put(sm, (i0 << cbits) | c, Fd::add(u, v));
                sm_put(sm, (i1 << cbits) | c, Fd::sub(u, v));
            }
            __syncthreads();
        }
        for (uint32_t t = threadIdx.x; t < tile_elems; t += BLOCK) {
            const uint32_t i = t >> cbits, c = t & (C - 1);
            Fe v = sm_get(sm, t);
            if (a.do_scale) v = Fd::mul(v, a.scale);
            st_fe(a.data, base | ((uint64_t)i << s0) | c, v);
        }
        __syncthreads();
    }
}
