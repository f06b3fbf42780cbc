// gbd_cluster_pcg_v4.cuh -- cluster-resident GBD-PCG, fourth generation: self-validating packets.
//
// fp32, n <= 16, power-of-two N >= 32 (the IIWA horizons 32 .. 512).  Same contract and the same
// floating-point operation order as the reference pcg<T,n,N> (GBD-PCG/include/pcg.cuh:54-218) ->
// bit-identical results.  Thread mapping, register-resident matrix rows and the redundant halo rows
// are those of v2 (gbd_cluster_pcg_v2.cuh); what changes is the exchange at the two per-iteration
// synchronisation points, which is the longest link of the dependent chain:
//
//  * every travelling value -- a knot row's dot partial (to all C CTAs), a boundary row element of
//    upsilon / r~ (to one neighbour) -- is ONE 8-byte packet {value, epoch} written with a single
//    64-bit store straight into the consumer's shared memory (st.relaxed.cluster.shared::cluster).
//    A packet validates itself: the consumer polls its own shared memory until the epoch word is
//    the one it expects.  No mbarrier, no staging through the producer's shared memory, no named
//    barrier, no shipping warp: the value leaves the producing thread the cycle it exists.
//  * all 16 lanes of a knot-row group end the per-knot GLASS tree holding the partial, and lane d
//    sends it to CTA d: one store instruction per warp covers all destinations.
//  * the consumer's poll doubles as the load of the N-way tree: lane l polls the PER partials
//    l, l + LW, l + 2 LW, ... (LW = N / PER), adds them in registers in the reference's stride order
//    (N/2, N/4, ... LW), then finishes with XOR-butterfly shuffles (strides LW/2 ... 1).  Every lane
//    ends with the total (a + b == b + a bit for bit), so there is no broadcast shuffle either:
//    N = 128 takes 3 register levels + 4 shuffle levels instead of 2 + 5 + broadcast.
//  * halo copies of r and p live in a register of the thread that owns them.
//
// Buffer reuse is safe without any extra handshake: a CTA can send phase A of iteration k+1 only after
// it has gathered phase B of iteration k, which every warp of every CTA sends only after it has read
// its phase-A packets of iteration k (true data dependence).  The prologue has its own buffers because
// its first exchange is neighbour-only and would break that chain across systems of a batch.
#pragma once
#include "gbd_cluster_pcg_v2.cuh"

namespace gbd {

template <bool WEAK>
__device__ __forceinline__ void st_packet(uint32_t cluster_addr, float v, uint32_t epoch)
{
    const uint64_t pk = ((uint64_t)epoch << 32) | (uint64_t)__float_as_uint(v);
    if constexpr (WEAK)   // A/B switch: SASS ST.E.64 instead of ST.E.64.STRONG.GPU
        asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(cluster_addr), "l"(pk) : "memory");
    else
        asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(cluster_addr), "l"(pk) : "memory");
}
__device__ __forceinline__ uint64_t ld_packet(uint32_t cta_addr)
{
    uint64_t pk;
    asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(pk) : "r"(cta_addr) : "memory");
    return pk;
}
__device__ __forceinline__ bool packet_ok(uint64_t pk, uint32_t epoch) { return (uint32_t)(pk >> 32) == epoch; }
__device__ __forceinline__ float packet_val(uint64_t pk) { return __uint_as_float((uint32_t)pk); }

// GLASS tree over CNT values held one per lane in lanes 0..CNT-1 of a G-lane group; EVERY lane of the
// group returns the total (same order as glass_tree_shfl: the serial tail is evaluated by all lanes).
template <uint32_t CNT, uint32_t G>
__device__ __forceinline__ float glass_tree_shfl_all(float x, uint32_t lane_in_group)
{
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t s = CNT;
#pragma unroll
    for (int lvl = 0; lvl < 8; ++lvl) {
        if (s > 3) {
            const uint32_t odd = s & 1u;
            s = (s - odd) / 2;
            const float y = __shfl_down_sync(FULL, x, s, G);
            float z = 0.f;
            if (odd) z = __shfl_sync(FULL, x, 2 * s, G);
            x = add_rn(x, y);
            if (odd && lane_in_group == 0) x = add_rn(x, z);
        }
    }
    const float x0 = __shfl_sync(FULL, x, 0, G), y1 = __shfl_sync(FULL, x, 1, G), y2 = __shfl_sync(FULL, x, 2, G);
    float tot = x0;
    if (s >= 2) tot = add_rn(tot, y1);
    if (s >= 3) tot = add_rn(tot, y2);
    return tot;
}

template <uint32_t n, uint32_t N, uint32_t C, uint32_t PER_ = 0>
struct ClusterPcg4 {
    using T = float;
    static_assert(n >= 2 && n <= 16, "v4 keeps a knot row in a 16-lane group");
    static_assert(is_pow2<N>::value && N >= 32, "the register tree needs a power-of-two knot count >= 32");
    static_assert(N % C == 0 && C >= 1 && C <= 16, "unsupported cluster shape");
    static constexpr uint32_t G = 16;
    static constexpr uint32_t R = N / C;
    static constexpr uint32_t NT = R * G;
    static_assert(R >= 2 && NT % 32 == 0 && NT <= 1024, "knot rows per CTA must fill whole warps");
    static constexpr uint32_t PER = PER_ ? PER_ : ((N / 32 > 8) ? N / 32 : 8);   // partials gathered per lane
    static constexpr uint32_t LW = N / PER;                      // distinct gather lanes (replicated over the warp)
    static_assert(LW >= 4 && LW <= 32, "gather width");
    static constexpr uint32_t W = 3 * n;
    static constexpr uint32_t TILE = 3 * n * n;
    static constexpr uint32_t XS = (n + 3) / 4 * 4;
    static constexpr uint32_t XLEN = (R + 2) * XS;
    static constexpr bool TMA_OK = (TILE * sizeof(T)) % 16 == 0;
    static constexpr size_t align16(size_t x) { return (x + 15) / 16 * 16; }
    // packet buffers (u64 each): 3 partial buffers [N], 4 halo buffers [from-left XS | from-right XS]
    static constexpr uint32_t PK_PART0 = 0, PK_PARTV = N, PK_PARTE = 2 * N;
    static constexpr uint32_t PK_HR = 3 * N, PK_H0 = PK_HR + 2 * XS, PK_HU = PK_H0 + 2 * XS, PK_HT = PK_HU + 2 * XS;
    static constexpr uint32_t PK_COUNT = PK_HT + 2 * XS;
    static constexpr size_t OFF_BAR = 0;
    static constexpr size_t OFF_PK = 16;
    static constexpr size_t OFF_XP = OFF_PK + sizeof(uint64_t) * PK_COUNT;
    static constexpr size_t OFF_XR = OFF_XP + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_S = OFF_XR + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_P = OFF_S + align16(sizeof(T) * R * TILE);
    static constexpr size_t SMEM_BYTES = OFF_P + align16(sizeof(T) * R * TILE);
};

// packet buffers cleared, tile mbarrier initialised; the caller follows it with a CTA barrier and cluster_sync()
template <uint32_t n, uint32_t N, uint32_t C, uint32_t PER_ = 0>
__device__ __forceinline__ void pcg_cluster_v4_init(unsigned char *smem_raw)
{
    using K = ClusterPcg4<n, N, C, PER_>;
    uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_PK);
    for (uint32_t i = threadIdx.x; i < K::PK_COUNT; i += blockDim.x) pk[i] = 0ull;      // epoch 0 is never sent
    if (threadIdx.x == 0) {
        mbar_init(reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR), 1);
        fence_mbar_init();
    }
}

// Solves systems first_sys, first_sys + sys_stride, ... < a.batch with this cluster.  Must be called by threads
// 0 .. NT-1 of every CTA of the cluster (and only by them: the intra-CTA barriers are named barriers over NT
// threads, so a launch may carry idle extra threads -- the drop-in pcg<T,n,N> runs under the caller's block size).
template <uint32_t n, uint32_t N, uint32_t C, bool WEAK = false, bool PROF = false, uint32_t PER_ = 0, bool EXACT_BLOCK = false>
__device__ __forceinline__ void pcg_cluster_v4_run(const PcgArgs<float> &a, unsigned char *smem_raw, uint32_t first_sys,
                                                   uint32_t sys_stride)
{
    using K = ClusterPcg4<n, N, C, PER_>;
    using T = float;
    constexpr uint32_t R = K::R, W = K::W, TILE = K::TILE, G = K::G, XS = K::XS, NT = K::NT, PER = K::PER, LW = K::LW;
    // EXACT_BLOCK: the launch carries exactly NT threads, so the plain CTA barrier (cheaper than a counted one) is safe
    auto cta_sync = [&]() { if constexpr (EXACT_BLOCK) __syncthreads(); else named_bar_sync(2, NT); };

    uint64_t *barT = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_PK);
    T *xp = reinterpret_cast<T *>(smem_raw + K::OFF_XP);   // p (prologue: lambda) rows, one halo row each side
    T *xr = reinterpret_cast<T *>(smem_raw + K::OFF_XR);   // r rows, one halo row each side
    T *sS = reinterpret_cast<T *>(smem_raw + K::OFF_S);
    T *sP = reinterpret_cast<T *>(smem_raw + K::OFF_P);

    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31u;
    const uint32_t j = t % G, k = t / G;            // lane inside the knot-row group, local knot row (< R)
    const bool is_row = j < n;
    const uint32_t jn = is_row ? j : 0;
    const uint32_t cr = cluster_ctarank();
    const uint32_t b = cr * R + k;
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    // halo duty: group 0 keeps the copy of the left neighbour's last row, group R-1 of the right one's first
    const bool lhalo = is_row && k == 0 && has_left, rhalo = is_row && k == R - 1 && has_right;
    const bool halo = lhalo || rhalo;

    const uint32_t pk_u = smem_u32(pk);
    // lane d of every knot-row group sends the group's partial to CTA d; halo lanes send their own boundary
    // element to the neighbour CTA.  (Pairing two packets per 16-byte store, or forwarding a CTA's partials
    // through one warp, was measured and is not faster: tools/micro/dsmem_exchange.cu, profiles/r01c_*.)
    const bool part_sender = j < C;
    const uint32_t peer_pk = map_to_cta(pk_u, part_sender ? j : cr) + 8u * b;
    const uint32_t nbr_pk = map_to_cta(pk_u, lhalo ? cr - 1 : (rhalo ? cr + 1 : cr)) + 8u * ((lhalo ? XS : 0u) + jn);
    const uint32_t my_halo_pk = pk_u + 8u * ((rhalo ? XS : 0u) + jn);
    const uint32_t my_part_pk = pk_u + 8u * (lane % LW);
    T *const halo_xr = xr + (rhalo ? (R + 1) * XS : 0u) + jn;
    T *const halo_xp = xp + (rhalo ? (R + 1) * XS : 0u) + jn;

    auto send_edge = [&](uint32_t halo_off, T edge, uint32_t ep) {
        if (halo) st_packet<WEAK>(nbr_pk + 8u * halo_off, edge, ep);
    };
    auto send_part = [&](uint32_t part_off, T prod, uint32_t ep) {
        const T partial = glass_tree_shfl_all<n, G>(prod, j);
        if (part_sender) st_packet<WEAK>(peer_pk + 8u * part_off, partial, ep);
    };
    uint32_t t_poll = 0;   // PROF: clock at the exit of the last poll loop
    // gather: poll the PER partials of this lane (+ the halo packet), N-way GLASS tree, total in every lane
    auto gather = [&](uint32_t part_off, uint32_t halo_off, uint32_t ep, T &edge) -> T {
        uint64_t q[PER], hq = 0;
        bool ok;
        SpinGuard guard;
        do {
            guard.tick();
            ok = true;
#pragma unroll
            for (uint32_t m = 0; m < PER; ++m) {
                q[m] = ld_packet(my_part_pk + 8u * (part_off + LW * m));
                ok = ok && packet_ok(q[m], ep);
            }
            if (halo) {
                hq = ld_packet(my_halo_pk + 8u * halo_off);
                ok = ok && packet_ok(hq, ep);
            }
        } while (!ok);
        if constexpr (PROF) asm volatile("mov.u32 %0, %%clock;" : "=r"(t_poll) : "l"(q[0]) : "memory");
        edge = packet_val(hq);
        T v[PER];
#pragma unroll
        for (uint32_t m = 0; m < PER; ++m) v[m] = packet_val(q[m]);
#pragma unroll
        for (uint32_t h = PER / 2; h >= 1; h /= 2) {
#pragma unroll
            for (uint32_t m = 0; m < PER / 2; ++m)
                if (m < h) v[m] = add_rn(v[m], v[m + h]);
        }
        T x = v[0];
#pragma unroll
        for (uint32_t s = LW / 2; s >= 1; s /= 2) x = add_rn(x, __shfl_xor_sync(0xffffffffu, x, s));
        return x;
    };

    uint32_t phT = 0, ep = 0;
    for (uint32_t sys = first_sys; sys < a.batch; sys += sys_stride) {
        const size_t moff = ((size_t)sys * N + (size_t)cr * R) * TILE;
        const size_t vbase = (size_t)sys * N * n;
        const T *gS = a.S + moff, *gP = a.Pinv + moff;
        const bool tma = K::TMA_OK && a.use_tma;

        if (tma) {
            if (t == 0) {
                fence_proxy_async();
                constexpr uint32_t total = (uint32_t)(sizeof(T) * R * TILE);
                constexpr uint32_t CHB = 16384;
                mbar_arrive_expect_tx(barT, 2 * total);
                for (uint32_t o = 0; o < total; o += CHB) {
                    const uint32_t len = total - o < CHB ? total - o : CHB;
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(sS) + o, reinterpret_cast<const unsigned char *>(gS) + o, len, barT);
                    tma_bulk_g2s(reinterpret_cast<unsigned char *>(sP) + o, reinterpret_cast<const unsigned char *>(gP) + o, len, barT);
                }
            }
        } else {
            for (uint32_t i = t; i < R * TILE; i += NT) { sS[i] = gS[i]; sP[i] = gP[i]; }
        }
        // lambda window: own knot rows plus one each side (absent neighbours read as zero); pads zeroed.
        // The r window's halo rows start at zero and are only ever written by their owner thread.
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) {
            const uint32_t row = i / XS, e = i % XS;
            const long kb = (long)(cr * R) + (long)row - 1;
            xp[i] = (e < n && kb >= 0 && kb < (long)N) ? a.lambda[vbase + (size_t)kb * n + e] : T(0);
            xr[i] = T(0);
        }
        T lam = T(0), gam = T(0);
        if (is_row) {
            lam = a.lambda[vbase + (size_t)b * n + j];
            gam = a.gamma[vbase + (size_t)b * n + j];
        }
        if (tma) mbar_wait(barT, phT);
        phT ^= 1u;
        cta_sync();

        // this thread's rows of S and Pinv live in registers for the whole solve; the tiles the reference
        // never reads (left of block row 0, right of block row N-1) are taken as zero
        T ms[W], mp[W];
        {
            const T *rowS = sS + k * TILE + jn, *rowP = sP + k * TILE + jn;
            const bool skip_l = b == 0, skip_r = b == N - 1;
#pragma unroll
            for (uint32_t c = 0; c < W; ++c) {
                const bool z = !is_row || (skip_l && c < n) || (skip_r && c >= 2 * n);
                ms[c] = z ? T(0) : rowS[c * n];
                mp[c] = z ? T(0) : rowP[c * n];
            }
        }
        const T *wp = xp + k * XS, *wr = xr + k * XS;      // 3-row windows [k-1 | k | k+1]
        T *own_p = xp + (k + 1) * XS + jn, *own_r = xr + (k + 1) * XS + jn;
        T edge;

        // ---- r = gamma - S*lambda ; exchange boundary rows of r            (pcg.cuh:118-126)
        T r = gam - chain_padded<T, n, XS>(ms, wp);
        if (is_row) *own_r = r;
        ++ep;
        send_edge(K::PK_HR, r, ep);
        if (halo) {
            uint64_t hq;
            SpinGuard guard;
            do { guard.tick(); hq = ld_packet(my_halo_pk + 8u * K::PK_HR); } while (!packet_ok(hq, ep));
            *halo_xr = packet_val(hq);
        }
        T rh = halo ? *halo_xr : T(0);                      // register copy of the neighbour's boundary r element
        cta_sync();
        // ---- r~ = Pinv*r ; p = r~ ; eta = r.r~                             (pcg.cuh:130-149)
        T rt = chain_padded<T, n, XS>(mp, wr);
        ++ep;
        send_edge(K::PK_H0, rt, ep);
        send_part(K::PK_PART0, mul_rn(r, rt), ep);
        T eta = gather(K::PK_PART0, K::PK_H0, ep, edge);
        T p = rt, ups = T(0);
        T ph = edge;                                        // register copy of the neighbour's boundary p element
        if (is_row) *own_p = p;
        if (halo) *halo_xp = ph;

        uint32_t iter = 0;
        uint8_t max_iter_exit = 1;
        // timeline build: %clock stamps of iterations 8..11, [iter][point][thread] (tools/timeline.py)
        auto stamp = [&](uint32_t pt, T dep) {
            if constexpr (PROF) {
                if (a.dbg && iter >= 8 && iter < 12) {
                    uint32_t c_;
                    asm volatile("mov.u32 %0, %%clock;" : "=r"(c_) : "f"(dep) : "memory");
                    a.dbg[((iter - 8) * 12 + pt) * (C * NT) + cr * NT + t] = c_;
                }
            }
        };
        auto stamp_poll = [&](uint32_t pt) {
            if constexpr (PROF) {
                if (a.dbg && iter >= 8 && iter < 12) a.dbg[((iter - 8) * 12 + pt) * (C * NT) + cr * NT + t] = t_poll;
            }
        };
        for (; iter < a.max_iter; ++iter) {
            cta_sync();
            stamp(0, p);
            // ---- upsilon = S*p ; v = p.upsilon                             (pcg.cuh:156-167)
            ups = chain_padded<T, n, XS>(ms, wp);
            stamp(1, ups);
            ++ep;
            send_edge(K::PK_HU, ups, ep);
            send_part(K::PK_PARTV, mul_rn(p, ups), ep);
            stamp(2, ups);
            const T alpha = eta / gather(K::PK_PARTV, K::PK_HU, ep, edge);      // :169
            stamp_poll(3);
            stamp(4, alpha);
            // ---- lambda += alpha p ; r -= alpha upsilon (own rows + halo copies)   (:172-176)
            lam = fma_rn(alpha, p, lam);
            r = fma_rn(-alpha, ups, r);
            if (is_row) *own_r = r;
            if (halo) { rh = fma_rn(-alpha, edge, rh); *halo_xr = rh; }
            cta_sync();
            stamp(5, r);
            // ---- r~ = Pinv*r ; eta' = r.r~                                 (:180-193)
            rt = chain_padded<T, n, XS>(mp, wr);
            stamp(6, rt);
            ++ep;
            send_edge(K::PK_HT, rt, ep);
            send_part(K::PK_PARTE, mul_rn(r, rt), ep);
            stamp(7, rt);
            const T eta_new = gather(K::PK_PARTE, K::PK_HT, ep, edge);
            stamp_poll(8);
            stamp(9, eta_new);
            if (abs_(eta_new) < a.exit_tol) { ++iter; max_iter_exit = 0; break; }   // :195
            const T beta = eta_new / eta;                                       // :199-200
            eta = eta_new;
            // ---- p = r~ + beta p (own rows + halo copies)                   (:203-206)
            p = fma_rn(beta, p, rt);
            if (is_row) *own_p = p;
            if (halo) { ph = fma_rn(beta, ph, edge); *halo_xp = ph; }
            stamp(10, p);
        }

        // ---- outputs                                                        (:212-215)
        if (is_row) {
            const size_t o = vbase + (size_t)b * n + j;
            a.lambda[o] = lam;
            if (a.r_out) a.r_out[o] = r;
            if (a.p_out) a.p_out[o] = p;
        }
        if (cr == 0 && t == 0) {
            store_result(a, sys, iter, max_iter_exit);
        }
        cta_sync();
        // batches: no CTA starts the next system's exchanges while a peer may still be reading this system's last packets /
        // phases (neighbour-only prologue exchanges do not order far CTAs).  Once per solve, off the iteration path; the
        // drop-in pcg<> (batch == 1, extra idle threads in the block) never gets here.
        if (a.batch > 1) cluster_sync();
    }
}

// C-ABI kernel: persistent clusters looping over a batch of systems
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB, bool WEAK = false, bool PROF = false, uint32_t PER_ = 0>
__global__ void __launch_bounds__(ClusterPcg4<n, N, C, PER_>::NT, MINB)
pcg_cluster_kernel_v4(const PcgArgs<float> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_cluster_v4_init<n, N, C, PER_>(smem_raw);
    __syncthreads();
    cluster_sync();   // all CTAs resident, packet buffers cleared, before any DSMEM traffic
    pcg_cluster_v4_run<n, N, C, WEAK, PROF, PER_, true>(a, smem_raw, cluster_idx(), cluster_count());
    cluster_sync();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
