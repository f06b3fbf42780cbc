// gbd_cluster_pcg_v5.cuh -- cluster-resident GBD-PCG, fifth generation: v3's thread mapping with v4's exchange.
//
// fp32, even n <= 16, power-of-two N >= 32.  Same contract and floating-point operation order as the reference
// pcg<T,n,N> (GBD-PCG/include/pcg.cuh:54-218) -> bit-identical results.
//
//  * thread mapping of gbd_cluster_pcg_v3.cuh: a knot row lives in an 8-lane group, lane j < n/2 owns matrix
//    rows j and j + n/2 of S and of Pinv in registers (two independent FMA chains per thread); the first level of
//    the per-knot GLASS tree (i, i + n/2) is one in-thread add.  CTAs are 8 * R threads: N = 128 runs as a
//    cluster of 8 CTAs x 128 threads -- half the polling warps of v4 at the same shape, and the block size the
//    reference launches pcg<> with (PCG_NUM_THREADS), so the body also runs behind the drop-in template.
//  * exchange of gbd_cluster_pcg_v4.cuh: every travelling value is one self-validating 8-byte {value, epoch}
//    packet stored straight into the consumer's shared memory and polled there; the poll doubles as the load of
//    the N-way tree (register adds for the strides >= LW, XOR-butterfly shuffles below).
//  * STAGE = true : tiles arrive by 1-D TMA bulk copies into shared memory and are lifted into registers
//    (C-ABI kernels, > 48 KB dynamic shared memory).  STAGE = false: each thread loads its rows straight from
//    global memory (drop-in kernel: the reference's launch site never opts in to > 48 KB).
#pragma once
#include "gbd_cluster_pcg_v3.cuh"
#include "gbd_cluster_pcg_v4.cuh"

namespace gbd {

template <uint32_t n, uint32_t N, uint32_t C, bool STAGE>
struct ClusterPcg5 {
    using T = float;
    static_assert(n % 2 == 0 && n >= 2 && n <= 16, "v5 needs an even block size <= 16");
    static_assert(is_pow2<N>::value && N >= 32, "the register tree needs a power-of-two knot count >= 32");
    static_assert(N % C == 0 && C >= 1 && C <= 8, "lane d of an 8-lane group sends to CTA d");
    static constexpr uint32_t H = n / 2;                 // active lanes per knot row; lane j owns rows j, j + H
    static constexpr uint32_t G = 8;
    static constexpr uint32_t R = N / C;
    static constexpr uint32_t NT = R * G;
    static_assert(R >= 2 && NT % 32 == 0 && NT <= 1024, "knot rows per CTA must fill whole warps");
    static constexpr uint32_t PER = (N / 32 > 8) ? N / 32 : 8;
    static constexpr uint32_t LW = N / PER;
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
    static constexpr size_t OFF_NEXT = 8;                // one packet {next system, solve sequence number} (work_counter mode)
    static constexpr size_t OFF_PK = 16;
    static constexpr size_t OFF_XP = OFF_PK + sizeof(uint64_t) * PK_COUNT;
    static constexpr size_t OFF_XR = OFF_XP + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_S = OFF_XR + align16(sizeof(T) * XLEN);
    static constexpr size_t OFF_P = OFF_S + (STAGE ? align16(sizeof(T) * R * TILE) : 0);
    static constexpr size_t SMEM_BYTES = OFF_P + (STAGE ? align16(sizeof(T) * R * TILE) : 0);
};

template <uint32_t n, uint32_t N, uint32_t C, bool STAGE>
__device__ __forceinline__ void pcg_cluster_v5_init(unsigned char *smem_raw)
{
    using K = ClusterPcg5<n, N, C, STAGE>;
    uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_PK);
    for (uint32_t i = threadIdx.x; i < K::PK_COUNT; i += blockDim.x) pk[i] = 0ull;      // epoch 0 is never sent
    if (threadIdx.x == 0) {
        *reinterpret_cast<uint64_t *>(smem_raw + K::OFF_NEXT) = 0ull;
        mbar_init(reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR), 1);
        fence_mbar_init();
    }
}

// Solves systems first_sys, first_sys + sys_stride, ... < a.batch with this cluster.  Called by threads 0 .. NT-1 of
// every CTA of the cluster, after pcg_cluster_v5_init + CTA barrier + cluster_sync.
// With a.work_counter (zero before the launch; sys_stride = number of clusters) only the FIRST system is fixed: while a
// solve runs, CTA 0 draws the cluster's next system from the counter and posts it to every CTA as one more packet, so
// clusters that meet short solves take more systems (iteration counts differ per system; a fixed stride leaves the
// tail of the batch to whichever clusters happened to draw the long ones).
template <uint32_t n, uint32_t N, uint32_t C, bool STAGE, bool EXACT_BLOCK>
__device__ __forceinline__ void pcg_cluster_v5_run(const PcgArgs<float> &a, unsigned char *smem_raw, uint32_t first_sys,
                                                   uint32_t sys_stride)
{
    using K = ClusterPcg5<n, N, C, STAGE>;
    using T = float;
    constexpr uint32_t R = K::R, W = K::W, TILE = K::TILE, G = K::G, H = K::H, XS = K::XS, NT = K::NT, PER = K::PER, LW = K::LW;
    auto cta_sync = [&]() { if constexpr (EXACT_BLOCK) __syncthreads(); else named_bar_sync(2, NT); };

    uint64_t *barT = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_BAR);
    uint64_t *pk = reinterpret_cast<uint64_t *>(smem_raw + K::OFF_PK);
    T *xp = reinterpret_cast<T *>(smem_raw + K::OFF_XP);
    T *xr = reinterpret_cast<T *>(smem_raw + K::OFF_XR);
    T *sS = reinterpret_cast<T *>(smem_raw + K::OFF_S);
    T *sP = reinterpret_cast<T *>(smem_raw + K::OFF_P);

    const uint32_t t = threadIdx.x;
    const uint32_t lane = t & 31u;
    const uint32_t j = t % G, k = t / G;               // lane in the knot-row group, local knot row (< R)
    const bool act = j < H;                            // lanes H..7 idle along
    const uint32_t j0 = act ? j : 0, j1 = j0 + H;      // the two matrix rows / vector elements of this thread
    const uint32_t cr = cluster_ctarank();
    const uint32_t b = cr * R + k;
    const bool has_left = cr > 0, has_right = cr + 1 < C;
    const bool lhalo = act && k == 0 && has_left, rhalo = act && k == R - 1 && has_right;
    const bool halo = lhalo || rhalo;

    const uint32_t pk_u = smem_u32(pk);
    const bool part_sender = j < C;
    const uint32_t peer_pk = map_to_cta(pk_u, part_sender ? j : cr) + 8u * b;
    const uint32_t nbr_pk = map_to_cta(pk_u, lhalo ? cr - 1 : (rhalo ? cr + 1 : cr)) + 8u * (lhalo ? XS : 0u);
    const uint32_t my_halo_pk = pk_u + 8u * (rhalo ? XS : 0u);
    const uint32_t my_part_pk = pk_u + 8u * (lane % LW);
    T *const halo_xr = xr + (rhalo ? (R + 1) * XS : 0u);
    T *const halo_xp = xp + (rhalo ? (R + 1) * XS : 0u);

    auto send_edge = [&](uint32_t halo_off, T e0, T e1, uint32_t ep) {
        if (halo) {
            st_packet<false>(nbr_pk + 8u * (halo_off + j0), e0, ep);
            st_packet<false>(nbr_pk + 8u * (halo_off + j1), e1, ep);
        }
    };
    // knot-row dot partial in GLASS order (level one (i, i + n/2) in-thread), sent to every CTA of the cluster
    auto send_part = [&](uint32_t part_off, T x0, T y0, T x1, T y1, uint32_t ep) {
        const T s = add_rn(mul_rn(x0, y0), mul_rn(x1, y1));
        const T partial = glass_tree_shfl_all<H, G>(s, j);
        if (part_sender) st_packet<false>(peer_pk + 8u * part_off, partial, ep);
    };
    // gather: poll the PER partials of this lane (+ the two halo packets), N-way GLASS tree, total in every lane
    auto gather = [&](uint32_t part_off, uint32_t halo_off, uint32_t ep, T &e0, T &e1) -> T {
        uint64_t q[PER], h0 = 0, h1 = 0;
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
                h0 = ld_packet(my_halo_pk + 8u * (halo_off + j0));
                h1 = ld_packet(my_halo_pk + 8u * (halo_off + j1));
                ok = ok && packet_ok(h0, ep) && packet_ok(h1, ep);
            }
        } while (!ok);
        e0 = packet_val(h0);
        e1 = packet_val(h1);
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

    const bool draw = a.work_counter != nullptr;
    const uint32_t next_u = smem_u32(smem_raw + K::OFF_NEXT);
    uint32_t phT = 0, ep = 0, seq = 0;
    for (uint32_t sys = first_sys; sys < a.batch;) {
        const size_t moff = ((size_t)sys * N + (size_t)cr * R) * TILE;
        const size_t vbase = (size_t)sys * N * n;
        const T *gS = a.S + moff, *gP = a.Pinv + moff;
        const bool tma = STAGE && K::TMA_OK && a.use_tma;

        if constexpr (STAGE) {
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
        }
        for (uint32_t i = t; i < (R + 2) * XS; i += NT) {
            const uint32_t row = i / XS, e = i % XS;
            const long kb = (long)(cr * R) + (long)row - 1;
            xp[i] = (e < n && kb >= 0 && kb < (long)N) ? a.lambda[vbase + (size_t)kb * n + e] : T(0);
            xr[i] = T(0);
        }
        T lam0 = T(0), lam1 = T(0), gam0 = T(0), gam1 = T(0);
        if (act) {
            const size_t o = vbase + (size_t)b * n;
            lam0 = a.lambda[o + j0]; lam1 = a.lambda[o + j1];
            gam0 = a.gamma[o + j0]; gam1 = a.gamma[o + j1];
        }

        // this thread's two rows of S and of Pinv live in registers for the whole solve; the tiles the
        // reference never reads (left of block row 0, right of block row N-1) are taken as zero
        T ms0[W], ms1[W], mp0[W], mp1[W];
        const bool skip_l = b == 0, skip_r = b == N - 1;
        if constexpr (STAGE) {
            if (tma) mbar_wait(barT, phT);
            phT ^= 1u;
            cta_sync();
            const T *rowS = sS + k * TILE, *rowP = sP + k * TILE;
#pragma unroll
            for (uint32_t c = 0; c < W; ++c) {
                const bool z = !act || (skip_l && c < n) || (skip_r && c >= 2 * n);
                ms0[c] = z ? T(0) : rowS[c * n + j0];
                ms1[c] = z ? T(0) : rowS[c * n + j1];
                mp0[c] = z ? T(0) : rowP[c * n + j0];
                mp1[c] = z ? T(0) : rowP[c * n + j1];
            }
        } else {
            const T *rowS = gS + (size_t)k * TILE, *rowP = gP + (size_t)k * TILE;
#pragma unroll
            for (uint32_t c = 0; c < W; ++c) {
                const bool z = !act || (skip_l && c < n) || (skip_r && c >= 2 * n);
                ms0[c] = z ? T(0) : __ldg(rowS + c * n + j0);
                ms1[c] = z ? T(0) : __ldg(rowS + c * n + j1);
                mp0[c] = z ? T(0) : __ldg(rowP + c * n + j0);
                mp1[c] = z ? T(0) : __ldg(rowP + c * n + j1);
            }
            cta_sync();
        }
        const T *wp = xp + k * XS, *wr = xr + k * XS;
        T *own_p = xp + (k + 1) * XS, *own_r = xr + (k + 1) * XS;
        T e0, e1;

        // ---- r = gamma - S*lambda ; exchange boundary rows of r            (pcg.cuh:118-126)
        T c0, c1;
        chain2_padded<n, XS>(ms0, ms1, wp, c0, c1);
        T r0 = gam0 - c0, r1 = gam1 - c1;
        if (act) { own_r[j0] = r0; own_r[j1] = r1; }
        ++ep;
        send_edge(K::PK_HR, r0, r1, ep);
        T rh0 = T(0), rh1 = T(0);                          // register copies of the neighbour's boundary r elements
        if (halo) {
            uint64_t h0, h1;
            SpinGuard guard;
            do {
                guard.tick();
                h0 = ld_packet(my_halo_pk + 8u * (K::PK_HR + j0));
                h1 = ld_packet(my_halo_pk + 8u * (K::PK_HR + j1));
            } while (!(packet_ok(h0, ep) && packet_ok(h1, ep)));
            rh0 = packet_val(h0); rh1 = packet_val(h1);
            halo_xr[j0] = rh0; halo_xr[j1] = rh1;
        }
        cta_sync();
        // ---- r~ = Pinv*r ; p = r~ ; eta = r.r~                             (pcg.cuh:130-149)
        T rt0, rt1;
        chain2_padded<n, XS>(mp0, mp1, wr, rt0, rt1);
        ++ep;
        send_edge(K::PK_H0, rt0, rt1, ep);
        send_part(K::PK_PART0, r0, rt0, r1, rt1, ep);
        T eta = gather(K::PK_PART0, K::PK_H0, ep, e0, e1);
        ++seq;
        if (draw && cr == 0 && t == 0) {
            // every CTA has entered this solve (its partials arrived), so it has consumed the previous post: the one
            // slot can be overwritten.  Posted now, read after the solve: the counter's round trip is off the path.
            const uint32_t nx = atomicAdd(a.work_counter, 1u) + sys_stride;
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) st_packet<false>(map_to_cta(next_u, c), __uint_as_float(nx), seq);
        }
        T p0 = rt0, p1 = rt1, u0 = T(0), u1 = T(0);
        T ph0 = e0, ph1 = e1;                              // register copies of the neighbour's boundary p elements
        if (act) { own_p[j0] = p0; own_p[j1] = p1; }
        if (halo) { halo_xp[j0] = ph0; halo_xp[j1] = ph1; }

        uint32_t iter = 0;
        uint8_t max_iter_exit = 1;
        for (; iter < a.max_iter; ++iter) {
            cta_sync();
            // ---- upsilon = S*p ; v = p.upsilon                             (pcg.cuh:156-167)
            chain2_padded<n, XS>(ms0, ms1, wp, u0, u1);
            ++ep;
            send_edge(K::PK_HU, u0, u1, ep);
            send_part(K::PK_PARTV, p0, u0, p1, u1, ep);
            const T alpha = eta / gather(K::PK_PARTV, K::PK_HU, ep, e0, e1);    // :169
            // ---- lambda += alpha p ; r -= alpha upsilon (own rows + halo copies)   (:172-176)
            lam0 = fma_rn(alpha, p0, lam0); lam1 = fma_rn(alpha, p1, lam1);
            r0 = fma_rn(-alpha, u0, r0); r1 = fma_rn(-alpha, u1, r1);
            if (act) { own_r[j0] = r0; own_r[j1] = r1; }
            if (halo) {
                rh0 = fma_rn(-alpha, e0, rh0); rh1 = fma_rn(-alpha, e1, rh1);
                halo_xr[j0] = rh0; halo_xr[j1] = rh1;
            }
            cta_sync();
            // ---- r~ = Pinv*r ; eta' = r.r~                                 (:180-193)
            chain2_padded<n, XS>(mp0, mp1, wr, rt0, rt1);
            ++ep;
            send_edge(K::PK_HT, rt0, rt1, ep);
            send_part(K::PK_PARTE, r0, rt0, r1, rt1, ep);
            const T eta_new = gather(K::PK_PARTE, K::PK_HT, ep, e0, e1);
            if (abs_(eta_new) < a.exit_tol) { ++iter; max_iter_exit = 0; break; }   // :195
            const T beta = eta_new / eta;                                       // :199-200
            eta = eta_new;
            // ---- p = r~ + beta p (own rows + halo copies)                   (:203-206)
            p0 = fma_rn(beta, p0, rt0); p1 = fma_rn(beta, p1, rt1);
            if (act) { own_p[j0] = p0; own_p[j1] = p1; }
            if (halo) {
                ph0 = fma_rn(beta, ph0, e0); ph1 = fma_rn(beta, ph1, e1);
                halo_xp[j0] = ph0; halo_xp[j1] = ph1;
            }
        }

        // ---- outputs                                                        (:212-215)
        if (act) {
            const size_t o = vbase + (size_t)b * n;
            a.lambda[o + j0] = lam0; a.lambda[o + j1] = lam1;
            if (a.r_out) { a.r_out[o + j0] = r0; a.r_out[o + j1] = r1; }
            if (a.p_out) { a.p_out[o + j0] = p0; a.p_out[o + j1] = p1; }
        }
        if (cr == 0 && t == 0) {
            store_result(a, sys, iter, max_iter_exit);
        }
        cta_sync();
        // batches: no CTA starts the next system's exchanges while a peer may still be reading this system's last packets /
        // phases (neighbour-only prologue exchanges do not order far CTAs).  Once per solve, off the iteration path; the
        // drop-in pcg<> (batch == 1, extra idle threads in the block) never gets here.
        if (a.batch > 1) cluster_sync();
        if (draw) {
            uint64_t q;
            uint32_t spins = 0;
            do {
                q = ld_packet(next_u);
                if (++spins > (1u << 26)) __trap();          // a lost post would otherwise hang the GPU
            } while (!packet_ok(q, seq));
            sys = __float_as_uint(packet_val(q));
        } else {
            sys += sys_stride;
        }
    }
}

// C-ABI kernel: persistent clusters looping over a batch of systems
template <uint32_t n, uint32_t N, uint32_t C, uint32_t MINB>
__global__ void __launch_bounds__(ClusterPcg5<n, N, C, true>::NT, MINB)
pcg_cluster_kernel_v5(const PcgArgs<float> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pcg_cluster_v5_init<n, N, C, true>(smem_raw);
    __syncthreads();
    cluster_sync();   // all CTAs resident, packet buffers cleared, before any DSMEM traffic
    pcg_cluster_v5_run<n, N, C, true, true>(a, smem_raw, cluster_idx(), cluster_count());
    cluster_sync();   // no CTA leaves while a peer may still write into its shared memory
}

}  // namespace gbd
