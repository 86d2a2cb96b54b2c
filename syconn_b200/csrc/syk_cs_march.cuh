// syk_cs_march.cuh -- tier 1 of the fused contact-site kernel for compile-time stencils (round-2 formulation).
//
// Same exact reformulation as syk_cs_fast.cuh (window histogram of block_processing_C.pyx:21-49 == separable box sums
// of id-indicator volumes over the few ids that occur near the tile, 8 ids per 32-bit word), with the sums taken in
// the order u -> v -> w and two CTA barriers per plane instead of four:
//   X(p)  relabel of input plane p (thread per 4-voxel quad): a quad that equals the quad of plane p-1 reuses its
//         compact word, anything else looks its ids up in a shared-memory hash whose 64-bit entries hold
//         {id, slot} (one load, no "pending" state).  The new compact word replaces the word of plane p-SU in the
//         in-place ring; where the two differ, the running u-sum U (4-bit fields, <= SU) of the voxel is updated.
//         In the same interval: arg-max of the boundary voxels of output plane p-SU (reads V of the previous step).
//   Y(p)  sliding v-sum of U (thread per column, 8-bit fields, ring of the last SV rows in registers) -> V;
//         boundary flags of plane p-1 (4 voxels per 32-bit operation on the compact indices); zeros / compaction of
//         the boundary voxels of output plane p-SU+1; slot ranks and recycling by one warp.
// Input planes arrive by TMA (cp.async.bulk.tensor.3d, box WP x VP x 1 uint32, zero fill outside the volume) into a
// ring of three raw planes, one plane ahead, completion on an mbarrier per buffer.
// Segments that ever need more than 8*GMAX ids near the marching plane are appended to `hard_list` and redone by the
// 64-id tier of syk_cs_fast.cuh, then by the generic kernel.  All arithmetic is integer and bit-identical to those.
#pragma once
#include <cuda.h>

#include "syk_cs_fast.cuh"

namespace csm {

using csfast::FastGeom;
using csfast::HASH;
using csfast::ld4;
using csfast::LU;
using csfast::nz_bytes;
using csfast::TV;
using csfast::TW;

template <int SU, int SV, int SW, int GMAX>
struct Cfg {
    static constexpr int VP = TV + SV - 1;
    static constexpr int WP = (TW + SW - 1 + 3) & ~3;
    static constexpr int PLANE = VP * WP;   // voxels of a haloed plane
    static constexpr int NQ = PLANE / 4;    // 4-voxel quads of a haloed plane
    static constexpr int QPR = WP / 4;
    static constexpr int KMAX = GMAX * 8;
    static constexpr int OU = SU / 2, OV = SV / 2, OW = SW / 2;
    static constexpr int CF = OU + 1;       // ring of boundary-flag planes
    static constexpr int NOQ = TV * TW / 4; // output quads per plane
    static constexpr int OPLANE = TV * WP;
    static constexpr int RAW_BYTES = (PLANE * 4 + 127) & ~127;
    static constexpr int NT = ((NQ + 31) / 32) * 32;  // one relabel quad per thread
    // dynamic shared memory (bytes, from a 128-byte aligned base)
    static constexpr int O_RAW = 0;
    static constexpr int O_COMP = O_RAW + 3 * RAW_BYTES;
    static constexpr int O_U = O_COMP + ((SU * PLANE + 15) & ~15);
    static constexpr int O_V = O_U + GMAX * PLANE * 4;
    static constexpr int O_CFLAG = O_V + GMAX * OPLANE * 8;
    static constexpr int O_ELIST = O_CFLAG + CF * TV * TW;
    static constexpr int TOTAL = O_ELIST + 2 * TV * TW * 2;
    static constexpr int DYN_BYTES = TOTAL + 128;  // slack for the manual 128-byte alignment
    static_assert(SU >= 5 && SU <= 15, "flags of plane p-OU must be complete one step ahead (OU >= 2); U has 4-bit fields");
    static_assert(SU * SV <= 255, "V has 8-bit fields");
    static_assert(SU * SV * SW <= 65535, "16-bit totals");
    static_assert(KMAX <= 32, "ranks / recycling are done by one warp");
    static_assert(((GMAX * WP + 31) / 32) * 32 + 32 + NOQ <= NT, "thread roles of interval Y do not fit");
    static_assert((WP * 4) % 16 == 0, "TMA box rows are multiples of 16 bytes");
};

template <int KMAX>
struct HashM {
    unsigned long long ent[HASH];  // id | (slot + 1) << 32; 0 = empty; high word 0 = id known, slot recycled
    unsigned ids[KMAX];            // id held by a slot
    int owner[KMAX];               // hash index owning the slot
    int lastp[KMAX];               // last plane in which the slot was seen
    __align__(16) unsigned short tb[KMAX];  // ((255 - rank by id) << 8) | slot, in the field order of the final sums
    unsigned freemask;             // bit s set: slot s is free
    unsigned seen[2];              // slots met by the relabel of plane p (double buffered)
    int n_edge[2];
    int ovf, newflag, ng;
    __align__(8) unsigned long long mbar[3];
};

template <typename H>
__device__ __forceinline__ int slot_take(H &h) {
    for (;;) {
        const unsigned m = *(volatile unsigned *)&h.freemask;
        if (m == 0u) return -1;
        const unsigned bit = m & (0u - m);
        if (atomicAnd(&h.freemask, ~bit) & bit) return __ffs((int)bit) - 1;
    }
}

// compact index (slot + 1) of id `lab` != 0; 0 when no slot is left (the segment is then given up)
template <typename H>
__device__ __noinline__ unsigned find_slot(H &h, unsigned lab) {
    unsigned i = (lab * 2654435761u) >> (32 - 8);  // HASH == 256
    for (int probes = 0; probes < 4 * HASH; ++probes) {
        const unsigned long long e = *(volatile unsigned long long *)&h.ent[i];
        const bool mine = (unsigned)e == lab;
        if (mine && (e >> 32)) return (unsigned)(e >> 32);
        if (mine || e == 0ull) {  // unknown id, or known with a recycled slot: take a slot, publish {id, slot} in one CAS
            const int s = slot_take(h);
            if (s < 0) break;
            const unsigned long long want = (unsigned long long)lab | ((unsigned long long)(s + 1) << 32);
            if (atomicCAS(&h.ent[i], e, want) == e) {
                h.ids[s] = lab;
                h.owner[s] = (int)i;
                h.newflag = 1;
                return (unsigned)(s + 1);
            }
            atomicOr(&h.freemask, 1u << s);  // somebody else was faster: look at the entry again
            continue;
        }
        i = (i + 1) & (HASH - 1);
    }
    h.ovf = 1;
    return 0u;
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CSM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CSM_DONE;\n"
        "bra CSM_WAIT;\n"
        "CSM_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

template <int SU, int SV, int SW, int GMAX, int MINB>
__global__ void __launch_bounds__((Cfg<SU, SV, SW, GMAX>::NT), MINB)
k_cs_march(const __grid_constant__ CUtensorMap tmap, unsigned long long *__restrict__ out, const FastGeom G,
           unsigned *__restrict__ hard_list, unsigned *__restrict__ hard_count) {
    using C = Cfg<SU, SV, SW, GMAX>;
    constexpr int VP = C::VP, WP = C::WP, PLANE = C::PLANE, NQ = C::NQ, QPR = C::QPR, KMAX = C::KMAX, NT = C::NT;
    constexpr int OU = C::OU, OV = C::OV, OW = C::OW, CF = C::CF, NOQ = C::NOQ, OPLANE = C::OPLANE;
    constexpr unsigned ALL_SLOTS = KMAX >= 32 ? 0xFFFFFFFFu : ((1u << KMAX) - 1u);
    constexpr bool PAIR_OK = 2 * SU * SV <= 255;
    using Hash = HashM<KMAX>;
    extern __shared__ unsigned char sm_raw[];
    unsigned char *sm = sm_raw + ((128u - ((unsigned)__cvta_generic_to_shared(sm_raw) & 127u)) & 127u);
    unsigned char *comp = sm + C::O_COMP;
    unsigned *U = reinterpret_cast<unsigned *>(sm + C::O_U);
    uint2 *V = reinterpret_cast<uint2 *>(sm + C::O_V);
    unsigned char *cflag = sm + C::O_CFLAG;
    unsigned short *elist = reinterpret_cast<unsigned short *>(sm + C::O_ELIST);
    __shared__ Hash H;

    const int tid = threadIdx.x;
    const unsigned raw_sa = (unsigned)__cvta_generic_to_shared(sm + C::O_RAW);
    const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(&H.mbar[0]);
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(bar_sa + 8u * i, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ring position / phase parity of the next TMA load to issue (li) and of the next plane to consume (ci)
    int li = 0, ci = 0, cpar = 0, inflight = 0;

    // thread roles: relabel quad (all threads < NQ); interval Y: v-sums (threads < GMAX*WP), ranks + recycling (the warp
    // after them), boundary flags + output compaction (the last NOQ threads)
    const int qb = tid / QPR, qc = (tid - qb * QPR) * 4;      // relabel quad: row, first column
    const int bg = tid / WP, bc = tid - bg * WP;              // v-sum: group, column
    constexpr int RANK0 = ((GMAX * WP + 31) / 32) * 32;       // first thread of the rank / recycling warp
    const int oq = tid - (NT - NOQ);                          // output quad of this thread (flags, compaction) or < 0
    const int oq_b = oq / (TW / 4), oq_c = (oq % (TW / 4)) * 4;

    for (long long seg = blockIdx.x; seg < G.nsegs; seg += gridDim.x) {
        const long long tw = seg % G.segs[2];
        const long long r0 = seg / G.segs[2];
        const long long tv = r0 % G.segs[1];
        const long long tu = r0 / G.segs[1];
        const long long u0 = tu * LU, v0 = tv * TV, w0 = tw * TW;  // output origin == input origin of the haloed block
        const int NP = (int)min((long long)(LU + SU - 1), G.n[0] - u0);
        __syncthreads();  // the previous segment is done with shared memory (also orders the mbarrier init)
        {   // ---- segment init: zero the compact ring and U; reset the hash ----
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 *p4 = reinterpret_cast<uint4 *>(sm + C::O_COMP);
            constexpr int n4 = (C::O_V - C::O_COMP) / 16;
            for (int i = tid; i < n4; i += NT) p4[i] = z;
            for (int i = tid; i < HASH; i += NT) H.ent[i] = 0ull;
            if (tid < KMAX) H.lastp[tid] = -(1 << 20);
            if (tid == 0) {
                H.freemask = ALL_SLOTS;
                H.seen[0] = H.seen[1] = 0u;
                H.n_edge[0] = H.n_edge[1] = 0;
                H.ovf = 0;
                H.newflag = 0;
                H.ng = 0;
                // first plane of the segment
                mbar_expect_tx(bar_sa + 8u * li, PLANE * 4);
                tma_load_3d(raw_sa + (unsigned)(li * C::RAW_BYTES), &tmap, (int)w0, (int)v0, (int)u0, bar_sa + 8u * li);
            }
            li = li == 2 ? 0 : li + 1;
            inflight = 1;
        }
        // ---- per-thread invariants of the segment ----
        unsigned mL = 0u, mR = 0u, mVlo = 0u, mVhi = 0u, dIn = 0u;
        if (oq >= 0) {
            const long long cv = v0 + oq_b + OV;
            mVlo = cv > 0 ? 0xFFFFFFFFu : 0u;
            mVhi = cv + 1 < G.n[1] ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long cw = w0 + oq_c + e + OW;
                if (cw > 0) mL |= 0xFFu << (8 * e);
                if (cw + 1 < G.n[2]) mR |= 0xFFu << (8 * e);
                if (v0 + oq_b < G.on[1] && w0 + oq_c + e < G.on[2]) dIn |= 0x80u << (8 * e);
            }
        }
        const long long d_rowoff = (v0 + oq_b) * G.ost[1] + (w0 + oq_c) * G.ost[2];
        __syncthreads();

        bool aborted = false;
        int rp = 0, rf = 0;     // p % SU (compact ring), p % CF (flag ring)
        int pbuf = 0;           // raw buffer of plane p - 1
        for (int p = 0; p <= NP; ++p, rp = (rp + 1 == SU ? 0 : rp + 1), rf = (rf + 1 == CF ? 0 : rf + 1)) {
            // ======================= interval X(p) =======================
            if (tid == 0 && p + 1 < NP) {  // plane p + 1 -> the buffer that held plane p - 2 (last read in X(p - 1))
                mbar_expect_tx(bar_sa + 8u * li, PLANE * 4);
                tma_load_3d(raw_sa + (unsigned)(li * C::RAW_BYTES), &tmap, (int)w0, (int)v0, (int)(u0 + p + 1), bar_sa + 8u * li);
            }
            if (p + 1 < NP) {
                li = li == 2 ? 0 : li + 1;
                inflight += 1;
            }
            // D(p - 1): boundary voxels of output plane uo = p - SU: final sum along w (16-bit fields) and arg-max
            {
                const int pd = p - 1, uo = pd - SU + 1;
                if (uo >= 0 && u0 + uo < G.on[0]) {
                    const int rfd = rf ? rf - 1 : CF - 1;  // pd % CF
                    const unsigned char *cf = cflag + (rfd + 1 == CF ? 0 : rfd + 1) * (TV * TW);  // flags of plane pd - OU
                    unsigned long long *orow = out + (u0 + uo) * G.ost[0];
                    const int ne = H.n_edge[pd & 1];
                    const unsigned short *el = elist + (pd & 1) * (TV * TW);
                    const int NG = H.ng;
                    for (int e = NT - 1 - tid; e < ne; e += NT) {  // from the last thread down
                        const int i = el[e];
                        const int b = i / TW, c = i - b * TW;
                        const int jc = cf[i] & 0x7F;
                        const int gc = (jc - 1) >> 3, nc = (jc - 1) & 7;
                        unsigned best = 0u;
                        for (int g = 0; g < NG; ++g) {
                            const uint2 *sp = V + g * OPLANE + b * WP + c;
                            unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
                            int r = 0;
                            if (PAIR_OK) {  // two v-sums fit an 8-bit field: add them before widening
#pragma unroll
                                for (; r + 1 < SW; r += 2) {
                                    const uint2 a = sp[r], d = sp[r + 1];
                                    const unsigned l = a.x + d.x, h = a.y + d.y;
                                    c0 += l & 0x00FF00FFu;
                                    c1 += (l >> 8) & 0x00FF00FFu;
                                    c2 += h & 0x00FF00FFu;
                                    c3 += (h >> 8) & 0x00FF00FFu;
                                }
                            }
#pragma unroll
                            for (; r < SW; ++r) {
                                const uint2 a = sp[r];
                                c0 += a.x & 0x00FF00FFu;         // slots 0, 4
                                c1 += (a.x >> 8) & 0x00FF00FFu;  // slots 2, 6
                                c2 += a.y & 0x00FF00FFu;         // slots 1, 5
                                c3 += (a.y >> 8) & 0x00FF00FFu;  // slots 3, 7
                            }
                            if ((c0 | c1 | c2 | c3) == 0u) continue;
                            if (g == gc) {  // the centre id does not compete
                                const unsigned keep = (nc & 4) ? 0x0000FFFFu : 0xFFFF0000u;
                                const int r4 = nc & 3;
                                if (r4 == 0) c0 &= keep;
                                else if (r4 == 1) c2 &= keep;
                                else if (r4 == 2) c1 &= keep;
                                else c3 &= keep;
                            }
                            const uint4 t = reinterpret_cast<const uint4 *>(H.tb)[g];
                            const unsigned k0 = max(__byte_perm(t.x, c0, 0x5410), __byte_perm(t.x, c0, 0x7632));
                            const unsigned k1 = max(__byte_perm(t.y, c1, 0x5410), __byte_perm(t.y, c1, 0x7632));
                            const unsigned k2 = max(__byte_perm(t.z, c2, 0x5410), __byte_perm(t.z, c2, 0x7632));
                            const unsigned k3 = max(__byte_perm(t.w, c3, 0x5410), __byte_perm(t.w, c3, 0x7632));
                            best = max(best, max(max(k0, k1), max(k2, k3)));
                        }
                        unsigned long long res = 0ull;
                        if (best >> 16) {
                            const unsigned center = H.ids[jc - 1], key = H.ids[best & 0xFFu];
                            res = center > key ? (((unsigned long long)key << 32) + center) : (((unsigned long long)center << 32) + key);
                        }
                        orow[(v0 + b) * G.ost[1] + (w0 + c) * G.ost[2]] = res;
                    }
                }
            }
            if (p == NP) break;  // that was the last output plane of the segment
            // A(p): relabel + u-sum update
            mbar_wait(bar_sa + 8u * ci, (unsigned)cpar);
            {
                const unsigned *rawc = reinterpret_cast<const unsigned *>(sm + C::O_RAW + ci * C::RAW_BYTES);
                const unsigned *rawp = reinterpret_cast<const unsigned *>(sm + C::O_RAW + pbuf * C::RAW_BYTES);
                unsigned seen = 0u;
                if (tid < NQ) {
                    const uint4 a = reinterpret_cast<const uint4 *>(rawc)[tid];
                    unsigned *ring = reinterpret_cast<unsigned *>(comp + rp * PLANE) + tid;
                    unsigned cn;
                    bool same = false;
                    if (p > 0) {
                        const uint4 q = reinterpret_cast<const uint4 *>(rawp)[tid];
                        same = a.x == q.x && a.y == q.y && a.z == q.z && a.w == q.w;
                    }
                    if (same) {
                        cn = reinterpret_cast<const unsigned *>(comp + (rp ? rp - 1 : SU - 1) * PLANE)[tid];
                    } else {
                        unsigned j0 = 0u, j1 = 0u, j2 = 0u, j3 = 0u;
                        if (a.x != 0u) j0 = find_slot(H, a.x);
                        if (a.y == a.x) j1 = j0;
                        if (a.z == a.x) j2 = j0;
                        if (a.w == a.x) j3 = j0;
                        unsigned need = 0u;  // the other distinct ids of the quad (a boundary crosses it) share one call site
                        if (a.y != 0u && a.y != a.x) need |= 2u;
                        if (a.z != 0u && a.z != a.x && a.z != a.y) need |= 4u;
                        if (a.w != 0u && a.w != a.x && a.w != a.y && a.w != a.z) need |= 8u;
                        while (need) {
                            const unsigned lab = (need & 2u) ? a.y : (need & 4u) ? a.z : a.w;
                            const unsigned j = find_slot(H, lab);
                            if (a.y == lab) j1 = j;
                            if (a.z == lab) j2 = j;
                            if (a.w == lab) j3 = j;
                            need &= need - 1u;
                        }
                        cn = j0 | (j1 << 8) | (j2 << 16) | (j3 << 24);
                    }
                    const unsigned co = *ring;   // plane p - SU (zero before the ring is full)
                    *ring = cn;
                    // slots present in this quad
                    {
                        const unsigned b0 = cn & 0xFFu;
                        if (cn == b0 * 0x01010101u) {
                            if (b0) seen = 1u << (b0 - 1u);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const unsigned j = (cn >> (8 * e)) & 0xFFu;
                                if (j) seen |= 1u << (j - 1u);
                            }
                        }
                    }
                    if (cn != co) {
                        const unsigned n0 = cn & 0xFFu, o0 = co & 0xFFu;
                        if (cn == n0 * 0x01010101u && co == o0 * 0x01010101u) {  // both quads uniform: 16-byte updates
                            if (n0) {
                                const int g = (int)(n0 - 1u) >> 3;
                                const unsigned d = 1u << (4u * ((n0 - 1u) & 7u));
                                uint4 *up = reinterpret_cast<uint4 *>(U + g * PLANE) + tid;
                                uint4 u4 = *up;
                                u4.x += d; u4.y += d; u4.z += d; u4.w += d;
                                *up = u4;
                            }
                            if (o0) {
                                const int g = (int)(o0 - 1u) >> 3;
                                const unsigned d = 1u << (4u * ((o0 - 1u) & 7u));
                                uint4 *up = reinterpret_cast<uint4 *>(U + g * PLANE) + tid;
                                uint4 u4 = *up;
                                u4.x -= d; u4.y -= d; u4.z -= d; u4.w -= d;
                                *up = u4;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const unsigned jn = (cn >> (8 * e)) & 0xFFu, jo = (co >> (8 * e)) & 0xFFu;
                                if (jn != jo) {
                                    if (jn) U[((int)(jn - 1u) >> 3) * PLANE + tid * 4 + e] += 1u << (4u * ((jn - 1u) & 7u));
                                    if (jo) U[((int)(jo - 1u) >> 3) * PLANE + tid * 4 + e] -= 1u << (4u * ((jo - 1u) & 7u));
                                }
                            }
                        }
                    }
                }
                seen = __reduce_or_sync(0xffffffffu, seen);
                if ((tid & 31) == 0 && seen) atomicOr(&H.seen[p & 1], seen);
            }
            pbuf = ci;
            ci = ci == 2 ? 0 : ci + 1;
            cpar ^= (ci == 0);
            inflight -= 1;
            __syncthreads();
            // ======================= interval Y(p) =======================
            if (H.ovf) { aborted = true; break; }
            if (tid == 0) {
                H.n_edge[(p + 1) & 1] = 0;  // D(p - 1) has read it
                H.seen[(p + 1) & 1] = 0u;
            }
            if (tid < GMAX * WP) {
                // B(p): V[g][b][c] = sum over rows b .. b+SV-1 of U[g][.][c], 8-bit fields {even slots | odd slots}
                const unsigned used = ~(*(volatile unsigned *)&H.freemask) & ALL_SLOTS;
                if (used >> (8 * bg)) {
                    const unsigned *up = U + bg * PLANE + bc;
                    uint2 *vp = V + bg * OPLANE + bc;
                    unsigned rx[SV], ry[SV];
                    unsigned ax = 0u, ay = 0u;
#pragma unroll
                    for (int r = 0; r < VP; ++r) {
                        const unsigned x = up[r * WP];
                        const unsigned lo = x & 0x0F0F0F0Fu, hi = (x >> 4) & 0x0F0F0F0Fu;
                        if (r >= SV) {  // the row that leaves the window first: fields never wrap
                            ax -= rx[r % SV];
                            ay -= ry[r % SV];
                        }
                        ax += lo;
                        ay += hi;
                        rx[r % SV] = lo;
                        ry[r % SV] = hi;
                        if (r >= SV - 1) vp[(r - (SV - 1)) * WP] = make_uint2(ax, ay);
                    }
                }
            } else if (tid >= RANK0 && tid < RANK0 + 32) {
                // one warp: last-seen planes, recycling of the slots whose id left the SU-plane window, ranks by id
                const int s = tid - RANK0;
                const unsigned seen = H.seen[p & 1];
                unsigned used = ~H.freemask & ALL_SLOTS;
                bool freed = false;
                if (s < KMAX && ((used >> s) & 1u)) {
                    if ((seen >> s) & 1u) H.lastp[s] = p;
                    else if (H.lastp[s] + SU <= p) freed = true;  // all its U fields are zero again
                }
                const unsigned fm = __ballot_sync(0xffffffffu, freed);
                if (freed) H.ent[H.owner[s]] = (unsigned long long)H.ids[s];  // id known, no slot
                used &= ~fm;
                if (s == 0) {
                    if (fm) atomicOr(&H.freemask, fm);
                    H.ng = used ? ((32 - __clz((int)used) + 7) >> 3) : 0;
                }
                if (H.newflag && s < KMAX && ((used >> s) & 1u)) {
                    const unsigned me = H.ids[s];
                    int r = 0;
                    for (unsigned m = used; m; m &= m - 1u) r += H.ids[__ffs((int)m) - 1] < me;
                    const int n = s & 7;
                    H.tb[(s & ~7) | ((n & 1) << 2) | (n & 2) | (n >> 2)] = (unsigned short)(((255 - r) << 8) | s);
                }
                __syncwarp();
                if (s == 0) H.newflag = 0;
            } else if (oq >= 0) {
                // C1: boundary flags of plane p - 1 (needs the compact planes p - 2, p - 1, p)
                if (p >= 2 && p - 1 >= OU && p - 1 <= LU - 1 + OU) {
                    const int rp1 = rp ? rp - 1 : SU - 1, rp2 = rp1 ? rp1 - 1 : SU - 1;
                    const unsigned char *c0 = comp + rp2 * PLANE, *c1 = comp + rp1 * PLANE, *c2 = comp + rp * PLANE;
                    const long long cu = u0 + p - 1;
                    const unsigned mUlo = cu > 0 ? 0xFFFFFFFFu : 0u, mUhi = cu + 1 < G.n[0] ? 0xFFFFFFFFu : 0u;
                    const int rowo = (oq_b + OV) * WP, col = oq_c + OW;
                    const unsigned Cc = ld4(c1 + rowo, col);
                    unsigned D = ((Cc ^ ld4(c0 + rowo, col)) & mUlo) | ((Cc ^ ld4(c2 + rowo, col)) & mUhi);
                    D |= ((Cc ^ ld4(c1 + rowo - WP, col)) & mVlo) | ((Cc ^ ld4(c1 + rowo + WP, col)) & mVhi);
                    D |= ((Cc ^ ld4(c1 + rowo, col - 1)) & mL) | ((Cc ^ ld4(c1 + rowo, col + 1)) & mR);
                    reinterpret_cast<unsigned *>(cflag + (rf ? rf - 1 : CF - 1) * (TV * TW))[oq] = Cc | (nz_bytes(D) & nz_bytes(Cc));
                }
                // zeros for the non-boundary voxels of output plane p - SU + 1 and the list of its boundary voxels
                const int uo = p - SU + 1;
                if (dIn != 0u && uo >= 0 && u0 + uo < G.on[0]) {
                    const unsigned char *cf = cflag + (rf + 1 == CF ? 0 : rf + 1) * (TV * TW);  // plane p - OU; CF == OU + 1
                    const unsigned E = reinterpret_cast<const unsigned *>(cf)[oq] & dIn;
                    unsigned long long *o4 = out + (u0 + uo) * G.ost[0] + d_rowoff;
                    if (E == 0u && dIn == 0x80808080u && G.out_vec) {
                        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                        reinterpret_cast<uint4 *>(o4)[0] = z;
                        reinterpret_cast<uint4 *>(o4)[1] = z;
                    } else {
                        const int n = __popc(E);
                        int base = n ? atomicAdd(&H.n_edge[p & 1], n) : 0;
                        unsigned short *el = elist + (p & 1) * (TV * TW);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (E & (0x80u << (8 * e))) el[base++] = (unsigned short)(oq * 4 + e);
                            else if (dIn & (0x80u << (8 * e))) o4[e * G.ost[2]] = 0ull;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (aborted && tid == 0) hard_list[atomicAdd(hard_count, 1u)] = (unsigned)seg;
        // drain: the load of a plane that will not be consumed (abort) must land before its buffer is reused
        while (inflight > 0) {
            mbar_wait(bar_sa + 8u * ci, (unsigned)cpar);
            ci = ci == 2 ? 0 : ci + 1;
            cpar ^= (ci == 0);
            inflight -= 1;
        }
    }
}

}  // namespace csm
