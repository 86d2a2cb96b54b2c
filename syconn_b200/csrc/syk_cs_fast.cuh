// syk_cs_fast.cuh -- fast path of the fused contact-site kernel (detect_seg_boundaries + process_block_nonzero).
//
// Exact reformulation of the reference's per-voxel window histogram (block_processing_C.pyx:21-49) as separable
// box sums of id-indicator volumes, evaluated only for the ids that actually occur near the tile:
//   * a CTA owns a TV x TW cross-section of the output and marches along the slowest axis (u) over a segment of LU
//     output planes; every input plane is read once per CTA (halo re-reads hit L2);
//   * the ids met on the way get compact 8-bit slots (shared-memory hash); a slot is recycled once its id has left the
//     su-plane window, so KMAX bounds the ids NEAR the marching plane, not per segment.  Only the compact planes of the
//     last su+1 input planes are kept in shared memory (1 byte per voxel);
//   * 8 ids share one 32-bit word as 4-bit indicator fields, so one integer add advances 8 box sums at once.  Per
//     plane: sliding sum along v of the entering plane and of the plane that leaves the window (window sv <= 15 ->
//     4-bit fields), difference added to the running sum over the last su planes (<= 255 -> 8-bit fields), and --
//     only for boundary voxels, which are compacted first -- the final sum along w in 16-bit fields, arg-max with
//     the reference tie-break (smallest id);
//   * the 6-neighbourhood boundary test runs on the compact 8-bit indices, four voxels per 32-bit operation;
//   * segments that ever hold more than KMAX ids in the window are appended to a device list and redone by the
//     generic kernel.
// All arithmetic is integer; results are bit-identical to the generic path.
#pragma once
#include <cuda.h>

#include "syk_common.cuh"

namespace csfast {

#ifdef SYK_NG_HIST
__device__ unsigned long long g_ids_hist[65];
#endif

#ifndef SYK_LU
#define SYK_LU 64
#endif
#ifndef SYK_TV
#define SYK_TV 16
#endif
constexpr int TV = SYK_TV;  // multiple of 8
constexpr int TW = 32;
constexpr int LU = SYK_LU;
constexpr int HASH = 256;
constexpr int NOQ = TV * TW / 4;  // output quads per plane (128)
constexpr int MAX_NQUAD = (TV + 16) * 48 / 4;  // host guarantees VP*WP/4 <= MAX_NQUAD (384 for TV = 16)
constexpr int MAX_WP = 48;        // host guarantees WP <= MAX_WP
// Two tiers of the same kernel: tier 1 keeps at most 24 ids near the marching plane (small shared-memory footprint,
// many CTAs per SM); the few segments that need more are redone by tier 2 (64 ids), then by the generic kernel.
#ifndef SYK_NT1
#define SYK_NT1 192
#endif
#ifndef SYK_MINB1
#define SYK_MINB1 5
#endif
constexpr int GMAX_T1 = 3, NT_T1 = SYK_NT1;
constexpr int GMAX_T2 = 8, NT_T2 = 320;

struct FastGeom {
    long long n[3];    // input extents (internal axes u, v, w)
    long long ist[3];  // input strides (elements)
    long long ost[3];  // output strides (elements)
    long long on[3];   // output extents
    int sten[3], off[3];
    int VP, WP;        // haloed plane dims: TV + sv - 1, TW + sw - 1 rounded up to a multiple of 4
    int CF;            // centre-flag ring depth
    int CR;            // compact-plane ring depth (su + 1)
    long long segs[3];
    long long nsegs;
    int elem_bytes;
    int vec4;          // 16-byte aligned uint32 rows: quads are loaded with one LDG.128
    int pair_ok;       // 2 * su * sv <= 255: two ring sums may be added in 8-bit fields before widening
    int out_vec;       // output rows 16-byte aligned and contiguous along w: zero runs are stored 16 bytes at a time
    int tma;           // vec4 && a tensor map describes the input: planes arrive as ONE cp.async.bulk.tensor box each
    // explicit edge mask (process_block_nonzero(edges, arr, ...), block_processing_C.pyx:66): replaces the boundary test;
    // nullptr = fused detect_seg_boundaries
    const void *edges;
    long long est[3];  // edge-volume strides (elements) along u, v, w
    int edge_bytes;    // 1 or 4
};

struct FastSmem {  // offsets in bytes into dynamic shared memory
    int raw, comp, ssum, cflag, elist, total;
};

__host__ __device__ __forceinline__ int fast_align16(int bytes) { return (bytes + 15) & ~15; }  // regions are 16-byte aligned
__host__ __device__ __forceinline__ int fast_align128(int bytes) { return (bytes + 127) & ~127; }  // TMA destinations
__host__ __device__ __forceinline__ FastSmem fast_layout(int VP, int WP, int CR, int CF, int GMAX) {
    FastSmem L;
    const int plane = VP * WP;
    const int oplane = TV * WP;
    L.raw = 0;
    L.comp = L.raw + 2 * fast_align128(plane * 4);          // two raw planes: the next one streams in (TMA box / cp.async)
    L.ssum = L.comp + fast_align16(CR * plane);            // ring of compact-index planes, 1 byte per voxel
    L.cflag = L.ssum + fast_align16(GMAX * oplane * 8);    // uint2 {even slots, odd slots} in 8-bit fields
    L.elist = L.cflag + fast_align16(CF * TV * TW);
    L.total = L.elist + fast_align16(TV * TW * 2) + 128;    // + slack: the kernel aligns its base to 128 bytes
    return L;
}
inline FastSmem fast_layout(const FastGeom &G, int GMAX) { return fast_layout(G.VP, G.WP, G.CR, G.CF, GMAX); }

constexpr unsigned SL_PENDING = 0xFFFFFFFFu;  // key inserted, slot being allocated
constexpr unsigned SL_DEAD = 0xFFFFFFFEu;     // id left the su-plane window: its slot was recycled
template <int GMAX>
struct HashT {
    static constexpr int KMAX = GMAX * 8;  // <= 64 (slot masks are 64 bit), <= 127 (index + boundary flag in a byte)
    static constexpr unsigned long long ALL_SLOTS = (KMAX >= 64) ? ~0ull : ((1ull << KMAX) - 1ull);
    unsigned keys[HASH];
    unsigned sl[HASH];             // compact slot of the key | SL_PENDING | SL_DEAD
    int lastp[HASH];               // last plane in which the key was seen
    unsigned ids[KMAX];            // id held by a slot
    int owner[KMAX];               // hash index owning the slot
    // ((255 - rank by id) << 8) | slot: tie-break key of the arg-max (smallest id wins).  The 8 entries of a group are
    // stored in the order of the 16-bit count fields {0,4 | 2,6 | 1,5 | 3,7} so that one 16-byte load pairs them up.
    __align__(16) unsigned short tb[KMAX];
    unsigned lut[GMAX][KMAX + 8];  // indicator word of compact index j in group g
    unsigned long long freemask;   // bit s set: slot s is free (lowest free slot is handed out first)
    int ovf, newflag, n_edge;
};

// hand out the lowest free compact slot to hash entry h (id lab); raises H.ovf when none is left
template <typename Hash>
__device__ __forceinline__ void slot_alloc(Hash &H, unsigned h, unsigned lab) {
    for (;;) {
        const unsigned long long m = *(volatile unsigned long long *)&H.freemask;
        if (m == 0ull) {
            H.ovf = 1;
            return;
        }
        const unsigned long long bit = m & (0ull - m);
        if (atomicAnd(&H.freemask, ~bit) & bit) {
            const int s = __ffsll((long long)bit) - 1;
            H.ids[s] = lab;
            H.owner[s] = (int)h;
            H.sl[h] = (unsigned)s;
            H.newflag = 1;
            return;
        }
    }
}

// find the hash index of `lab` (!= 0) at plane p, inserting it (and allocating a compact slot) when it is new or
// when its slot was recycled; slots become visible to other threads after the next barrier
template <typename Hash>
__device__ __noinline__ int hash_find_insert(Hash &H, unsigned lab, int p) {
    unsigned h = (lab * 2654435761u) >> (32 - 8);  // HASH == 256
    for (int probes = 0; probes < HASH; ++probes) {
        const unsigned cur = H.keys[h];
        if (cur == lab) {
            H.lastp[h] = p;
            if (H.sl[h] == SL_DEAD && atomicCAS(&H.sl[h], SL_DEAD, SL_PENDING) == SL_DEAD) slot_alloc(H, h, lab);
            return (int)h;
        }
        if (cur == 0u) {
            const unsigned prev = atomicCAS(&H.keys[h], 0u, lab);
            if (prev == 0u) {
                H.lastp[h] = p;
                slot_alloc(H, h, lab);
                return (int)h;
            }
            if (prev == lab) {
                H.lastp[h] = p;
                return (int)h;
            }
        }
        h = (h + 1) & (HASH - 1);
    }
    H.ovf = 1;  // hash full
    return 0;
}

// four consecutive bytes starting at byte column `col` of a row of 8-bit indices (row is 4-byte aligned)
__device__ __forceinline__ unsigned ld4(const unsigned char *row, int col) {
    const unsigned *w = reinterpret_cast<const unsigned *>(row) + (col >> 2);
    const unsigned sh = (unsigned)(col & 3) * 8u;
    const unsigned lo = w[0];
    if (sh == 0u) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ unsigned nz_bytes(unsigned x) {  // 0x80 in every byte of x that is non-zero
    return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}

// 16-byte asynchronous copy global -> shared (L2 only); bytes past src_bytes are zero-filled
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// TMA plane loads: one cp.async.bulk.tensor.3d box (WP x VP x 1 uint32, zero fill outside the volume) per input plane,
// completion on an mbarrier per raw buffer
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
        "CSF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CSF_DONE;\n"
        "bra CSF_WAIT;\n"
        "CSF_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// indicator word of compact index j in the group whose first slot is gsh / 4 - 1 ... : 1 << 4 * (j - 1 - 8 g) when j lies in
// group g, else 0.  gsh = 4 + 32 g.  PTX shl clamps shift amounts >= 32 (result 0), and 4 j - gsh wraps to a huge amount
// for j below the group (background included), so no table and no branch are needed.
#ifndef SYK_VSUM_LUT
__device__ __forceinline__ unsigned ind_word(unsigned j, unsigned gsh) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(1u), "r"(4u * j - gsh));
    return r;
}
#define SYK_IND(j) ind_word((unsigned)(j), gsh)
#else
#define SYK_IND(j) lut[(j)]
#endif

// sliding v-sum (window sv) of the indicator words of one column of a compact plane: acc[q], q = 0..7
__device__ __forceinline__ void vsum8(const unsigned char *col, int WP, int sv, const unsigned *lut, unsigned gsh, unsigned (&acc)[8]) {
    unsigned head[8];
    unsigned a = 0u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        head[r] = SYK_IND(col[r * WP]);
        if (r < sv) a += head[r];
    }
    for (int r = 8; r < sv; ++r) a += SYK_IND(col[r * WP]);
    acc[0] = a;
#pragma unroll
    for (int q = 1; q < 8; ++q) {
        a += SYK_IND(col[(q + sv - 1) * WP]) - head[q - 1];
        acc[q] = a;
    }
}

// seg_list == nullptr: all G.nsegs segments; otherwise the *seg_count segments named in seg_list (tier 2).
// SU/SV/SW != 0: stencil known at compile time (must equal G.sten) -- plane pitches, ring depths and the shared-memory
// layout fold into immediates and the window loops unroll.
// EDGES: explicit edge mask (G.edges) instead of the fused boundary test -- a separate instantiation so that detect_cs pays nothing
// TMA: input planes arrive as one cp.async.bulk.tensor box each (compile-time: the LDG / cp.async bookkeeping costs registers)
template <bool VEC4, int GMAX, int NT, int MINB, int SU = 0, int SV = 0, int SW = 0, bool EDGES = false, bool TMA = false>
__global__ void __launch_bounds__(NT, MINB)
k_cs_fast(const void *__restrict__ arr, unsigned long long *__restrict__ out, FastGeom G, FastSmem Lrt,
          const unsigned *__restrict__ seg_list, const unsigned *__restrict__ seg_count,
          unsigned *__restrict__ hard_list, unsigned *__restrict__ hard_count, const __grid_constant__ CUtensorMap tmap) {
    using Hash = HashT<GMAX>;
    constexpr int KMAX = Hash::KMAX;
    constexpr unsigned long long ALL_SLOTS = Hash::ALL_SLOTS;
    constexpr int NQMAX = SU > 0 ? ((TV + SV - 1) * ((TW + SW - 1 + 3) & ~3)) / 4 : MAX_NQUAD;  // quads per haloed plane
    constexpr int MAXQ = (NQMAX + NT - 1) / NT;                 // relabel quads (4 voxels along w) per thread
    constexpr int MAXIT = (GMAX * (TV / 8) * MAX_WP + NT - 1) / NT;  // v-pass items (8 outputs each) per thread
    constexpr bool FIX = SU > 0;
    const int su = FIX ? SU : G.sten[0], sv = FIX ? SV : G.sten[1], sw = FIX ? SW : G.sten[2];
    const int ou = FIX ? SU / 2 : G.off[0], ov = FIX ? SV / 2 : G.off[1], ow = FIX ? SW / 2 : G.off[2];
    const int VP = FIX ? TV + SV - 1 : G.VP, WP = FIX ? ((TW + SW - 1 + 3) & ~3) : G.WP;
    const int CR = FIX ? SU + 1 : G.CR, CF = FIX ? SU / 2 + 1 : G.CF;
    const bool pair_ok = FIX ? (2 * SU * SV <= 255) : (G.pair_ok != 0);
    const FastSmem L = FIX ? fast_layout(VP, WP, CR, CF, GMAX) : Lrt;
    extern __shared__ __align__(16) unsigned char sm_base[];
    unsigned char *sm = sm_base + ((128u - ((unsigned)__cvta_generic_to_shared(sm_base) & 127u)) & 127u);
    unsigned *raw = reinterpret_cast<unsigned *>(sm + L.raw);  // raw[(p & 1) * rawpitch ...]: input plane p
    unsigned char *comp = sm + L.comp;
    uint2 *ssum = reinterpret_cast<uint2 *>(sm + L.ssum);
    unsigned char *cflag = sm + L.cflag;
    unsigned short *elist = reinterpret_cast<unsigned short *>(sm + L.elist);
    __shared__ Hash H;

    const int tid = threadIdx.x;
    const int plane = VP * WP, oplane = TV * WP;
    const int nquad = plane >> 2, qpr = WP >> 2;  // quads per plane / per row
    const int rawpitch = fast_align128(plane * 4) >> 2;  // words between the two raw planes
    static_assert(!TMA || VEC4, "TMA planes need 16-byte aligned uint32 rows");
    __shared__ __align__(8) unsigned long long mbar[2];
    const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(&mbar[0]);
    const unsigned raw_sa = (unsigned)__cvta_generic_to_shared(raw);
    if (TMA && tid == 0) {
        mbar_init(bar_sa, 1u);
        mbar_init(bar_sa + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned par0 = 0u, par1 = 0u;  // phase parity of the next completion of raw buffer 0 / 1 (uniform over the CTA)
    int pend = -1;                  // raw buffer with a TMA load in flight that nobody has waited for yet

    for (int i = tid; i < GMAX * (KMAX + 8); i += NT) {
        const int g = i / (KMAX + 8), j = i - g * (KMAX + 8);
        const int s = j - 1;
        H.lut[g][j] = (j != 0 && (s >> 3) == g) ? (1u << ((s & 7) * 4)) : 0u;
    }

    // Segments are handed out dynamically after the first wave (hard_count[2] / [3] are the queue heads of tier 1 / tier 2):
    // ~7 segments per CTA of unequal length (the last u segment is short, volume edges clip tiles) leave a static round
    // robin with a long tail.
    const long long nwork = seg_list ? (long long)(*seg_count) : G.nsegs;
    unsigned *queue = hard_count + (seg_list ? 3 : 2);
    __shared__ unsigned s_next;
    for (long long wi = blockIdx.x; wi < nwork;) {
        const long long seg = seg_list ? (long long)seg_list[wi] : wi;
        const long long tw = seg % G.segs[2];
        const long long r0 = seg / G.segs[2];
        const long long tv = r0 % G.segs[1];
        const long long tu = r0 / G.segs[1];
        const long long u0 = tu * LU, v0 = tv * TV, w0 = tw * TW;  // output origin == input origin of the haloed block
        const int NP = (int)min((long long)(LU + su - 1), G.n[0] - u0);  // input planes of the segment (the last one is short)
        __syncthreads();
        if (tid == 0) s_next = gridDim.x + atomicAdd(queue, 1u);  // read after the barrier that ends the segment init
        // ---- segment init: zero running sums / hash ----
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 *p4 = reinterpret_cast<uint4 *>(sm + L.ssum);
            const int n4 = (L.cflag - L.ssum) / 16;
            for (int i = tid; i < n4; i += NT) p4[i] = z;
            for (int i = tid; i < HASH; i += NT) {
                H.keys[i] = 0u;
                H.sl[i] = SL_PENDING;
                H.lastp[i] = -(1 << 20);
            }
            if (tid == 0) {
                H.freemask = ALL_SLOTS;
                H.ovf = 0;
                H.newflag = 0;
                H.n_edge = 0;
            }
        }
        // ---- per-thread invariants of the segment (no divisions inside the plane loop) ----
        long long qoff[MAXQ];   // global element offset of the quad at plane u0 (without the u term)
        unsigned qok[MAXQ];     // bit k: voxel k of the quad lies inside the volume (v, w bounds)
#pragma unroll
        for (int k = 0; k < MAXQ; ++k) {
            const int q = tid + k * NT;
            qok[k] = 0u;
            qoff[k] = 0;
            if (q < nquad) {
                const int b = q / qpr, c = (q - b * qpr) * 4;
                const long long gv = v0 + b, gw = w0 + c;
                qoff[k] = gv * G.ist[1] + gw * G.ist[2];
                if (gv < G.n[1])
                    for (int e = 0; e < 4; ++e) qok[k] |= (gw + e < G.n[2]) ? (1u << e) : 0u;
            }
        }
        int itoff[MAXIT], itg[MAXIT];  // v-pass items: offset inside a plane (row vc*8, column c), group
#pragma unroll
        for (int k = 0; k < MAXIT; ++k) {
            const int it = tid + k * NT;
            const int c = it % WP;
            const int r = it / WP;
            itoff[k] = (r % (TV / 8)) * 8 * WP + c;
            itg[k] = r / (TV / 8);
        }
        // output-quad invariants (threads 0..NOQ-1): byte masks of existing neighbours / of in-bounds outputs
        const int oq_b = tid / (TW / 4), oq_c = (tid % (TW / 4)) * 4;
        unsigned mL = 0u, mR = 0u, mVlo = 0u, mVhi = 0u, mIn = 0u;
        if (tid < NOQ) {
            const long long cv = v0 + oq_b + ov;
            mVlo = cv > 0 ? 0xFFFFFFFFu : 0u;
            mVhi = cv + 1 < G.n[1] ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long cw = w0 + oq_c + e + ow;
                if (cw > 0) mL |= 0xFFu << (8 * e);
                if (cw + 1 < G.n[2]) mR |= 0xFFu << (8 * e);
                if (v0 + oq_b < G.on[1] && w0 + oq_c + e < G.on[2]) mIn |= 0x80u << (8 * e);
            }
        }
        const long long out_rowoff = (v0 + oq_b) * G.ost[1] + (w0 + oq_c) * G.ost[2];

        auto load_quad = [&](int k, long long gu, uint4 &v) {
            v = make_uint4(0u, 0u, 0u, 0u);
            if (qok[k] == 0u || gu >= G.n[0]) return;
            const long long o = gu * G.ist[0] + qoff[k];
            if (VEC4) {
                if (qok[k] == 15u) {
                    v = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned *>(arr) + o));
                    return;
                }
            }
            unsigned t[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                t[e] = 0u;
                if (qok[k] & (1u << e)) {
                    const long long a = o + e * G.ist[2];
                    t[e] = G.elem_bytes == 8 ? (unsigned)__ldg((const unsigned long long *)arr + a) : __ldg((const unsigned *)arr + a);
                }
            }
            v = make_uint4(t[0], t[1], t[2], t[3]);
        };

        // zeros for the non-boundary voxels of output plane p - su + 1 and the compacted list of its boundary voxels
        // (threads NT-NOQ .. NT-1, so that it overlaps the boundary test of threads 0 .. NOQ-1)
        const bool early_d = ou >= 2;  // the flags of plane p - ou are then already complete in phase C of step p
        const int dq = tid - (NT - NOQ);
        const int dq_b = dq / (TW / 4), dq_c = (dq % (TW / 4)) * 4;
        unsigned dIn = 0u;
        if (dq >= 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (v0 + dq_b < G.on[1] && w0 + dq_c + e < G.on[2]) dIn |= 0x80u << (8 * e);
        }
        const long long d_rowoff = (v0 + dq_b) * G.ost[1] + (w0 + dq_c) * G.ost[2];
        auto compact_outputs = [&](int p, int rf) {  // rf == p % CF
            const int uo = p - su + 1;
            if (dq < 0 || dIn == 0u || uo < 0 || u0 + uo >= G.on[0]) return;
            const unsigned char *cf = cflag + (rf + 1 == CF ? 0 : rf + 1) * (TV * TW);  // plane uo + ou == p - ou; CF == ou + 1
            const unsigned E = reinterpret_cast<const unsigned *>(cf)[dq] & dIn;  // 0x80 per in-bounds boundary voxel
            unsigned long long *o4 = out + (u0 + uo) * G.ost[0] + d_rowoff;
            if (E == 0u && dIn == 0x80808080u && G.out_vec) {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                reinterpret_cast<uint4 *>(o4)[0] = z;
                reinterpret_cast<uint4 *>(o4)[1] = z;
            } else {
                const int n = __popc(E);
                int base = n ? atomicAdd(&H.n_edge, n) : 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (E & (0x80u << (8 * e))) elist[base++] = (unsigned short)(dq * 4 + e);
                    else if (dIn & (0x80u << (8 * e))) o4[e * G.ost[2]] = 0ull;
                }
            }
        };

        // prologue: plane 0 -> raw
        if (TMA) {
            if (tid == 0) {
                mbar_expect_tx(bar_sa, (unsigned)plane * 4u);
                tma_load_3d(raw_sa, &tmap, (int)w0, (int)v0, (int)u0, bar_sa);
            }
            mbar_wait(bar_sa, par0);
            par0 ^= 1u;
        } else {
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                if (q < nquad) {
                    uint4 v;
                    load_quad(k, u0, v);
                    reinterpret_cast<uint4 *>(raw)[q] = v;
                }
            }
        }
        __syncthreads();
        bool aborted = false;
        int rp = 0, rf = 0;  // p % CR, p % CF (ring positions, kept without divisions)
        for (int p = 0; p < NP; ++p, rp = (rp + 1 == CR ? 0 : rp + 1), rf = (rf + 1 == CF ? 0 : rf + 1)) {
            const long long gu = u0 + p;
            const int rp1 = rp ? rp - 1 : CR - 1, rp2 = rp1 ? rp1 - 1 : CR - 1, rpn = rp + 1 == CR ? 0 : rp + 1;
            // A. prefetch plane p+1 into registers; relabel pass 1 (find / insert the key of every voxel of plane p)
            // (16-byte aligned uint32 rows: cp.async straight into the other raw plane -- nothing waits for it until the
            // last barrier of this step; otherwise through registers, stored in phase B)
            const unsigned *rawc = raw + (p & 1) * rawpitch;
            unsigned *rawn = raw + ((p + 1) & 1) * rawpitch;
            uint4 pre[VEC4 ? 1 : MAXQ];
            int hidx[MAXQ][4];
            if (TMA && tid == 0 && p + 1 < NP) {  // everybody left raw plane p - 1 at the last barrier of the previous step
                const unsigned b = bar_sa + 8u * (unsigned)((p + 1) & 1);
                mbar_expect_tx(b, (unsigned)plane * 4u);
                tma_load_3d(raw_sa + (unsigned)(((p + 1) & 1) * rawpitch * 4), &tmap, (int)w0, (int)v0, (int)(u0 + p + 1), b);
            }
            if (TMA && p + 1 < NP) pend = (p + 1) & 1;
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                if (TMA) {
                } else if (VEC4) {
                    if (q < nquad && p + 1 < NP) {
                        const bool in = gu + 1 < G.n[0] && qok[k] != 0u;
                        const unsigned *src = reinterpret_cast<const unsigned *>(arr) + (in ? (gu + 1) * G.ist[0] + qoff[k] : 0);
                        cp_async16(rawn + 4 * q, src, in ? 4 * __popc(qok[k]) : 0);  // in-bounds voxels of a quad are a prefix
                    }
                } else {
                    pre[k] = make_uint4(0u, 0u, 0u, 0u);
                    if (q < nquad && p + 1 < NP) load_quad(k, gu + 1, pre[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                hidx[k][0] = hidx[k][1] = hidx[k][2] = hidx[k][3] = -1;
                if (q < nquad) {
                    const uint4 a = reinterpret_cast<const uint4 *>(rawc)[q];
                    int h0 = -1, h1 = -1, h2 = -1, h3 = -1;
                    if (a.x != 0u) h0 = hash_find_insert(H, a.x, p);
                    if (a.y == a.x) h1 = h0;
                    if (a.z == a.x) h2 = h0;
                    if (a.w == a.x) h3 = h0;
                    // the other distinct labels of the quad (rare: a boundary crosses it) share one call site
                    unsigned need = 0u;
                    if (a.y != 0u && a.y != a.x) need |= 2u;
                    if (a.z != 0u && a.z != a.x && a.z != a.y) need |= 4u;
                    if (a.w != 0u && a.w != a.x && a.w != a.y && a.w != a.z) need |= 8u;
                    while (need) {
                        const unsigned lab = (need & 2u) ? a.y : (need & 4u) ? a.z : a.w;
                        const int h = hash_find_insert(H, lab, p);
                        if (a.y == lab) h1 = h;
                        if (a.z == lab) h2 = h;
                        if (a.w == lab) h3 = h;
                        need &= need - 1u;
                    }
                    hidx[k][0] = h0;
                    hidx[k][1] = h1;
                    hidx[k][2] = h2;
                    hidx[k][3] = h3;
                }
            }
            __syncthreads();
            // B. slots are published: compact-index plane; next raw plane
            if (H.ovf) { aborted = true; break; }
            if (tid == 0) H.n_edge = 0;  // everybody is past the previous plane's boundary-voxel loop
            const unsigned long long used = ~H.freemask & ALL_SLOTS;  // stable until phase C
            if (H.newflag && tid < KMAX && ((used >> tid) & 1ull)) {  // ranks by id (arg-max tie-break: smallest id wins)
                const unsigned me = H.ids[tid];
                int r = 0;
                for (unsigned long long m = used; m; m &= m - 1ull) r += H.ids[__ffsll((long long)m) - 1] < me;
                const int n = tid & 7;
                H.tb[(tid & ~7) | ((n & 1) << 2) | (n & 2) | (n >> 2)] = (unsigned short)(((255 - r) << 8) | tid);
            }
            const int NG = used ? ((64 - __clzll((long long)used) + 7) >> 3) : 0;
#ifdef SYK_NG_HIST
            if (tid == 0) atomicAdd(&g_ids_hist[__popcll(used)], 1ull);  // development: live ids per plane
#endif
            unsigned char *cp = comp + rp * plane;
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                if (q < nquad) {
                    const int h0 = hidx[k][0], h1 = hidx[k][1], h2 = hidx[k][2], h3 = hidx[k][3];
                    const unsigned j0 = h0 < 0 ? 0u : H.sl[h0] + 1u;
                    unsigned w4 = j0 * 0x01010101u;  // uniform quad (the common case)
                    if (!(h1 == h0 && h2 == h0 && h3 == h0)) {
                        const unsigned j1 = h1 < 0 ? 0u : H.sl[h1] + 1u;
                        const unsigned j2 = h2 < 0 ? 0u : H.sl[h2] + 1u;
                        const unsigned j3 = h3 < 0 ? 0u : H.sl[h3] + 1u;
                        w4 = j0 | (j1 << 8) | (j2 << 16) | (j3 << 24);
                    }
                    reinterpret_cast<unsigned *>(cp)[q] = w4;
                    if (!VEC4) reinterpret_cast<uint4 *>(rawn)[q] = pre[VEC4 ? 0 : k];
                }
            }
            __syncthreads();
            if (tid == 0) H.newflag = 0;
            // recycle the slots of ids that left the su-plane window (after this step their sum fields are zero again)
            if (tid < KMAX && ((used >> tid) & 1ull)) {
                const int h = H.owner[tid];
                if (H.lastp[h] + su <= p) {
                    H.sl[h] = SL_DEAD;
                    atomicOr(&H.freemask, 1ull << tid);
                }
            }
            // C1. boundary flags of plane p-1 (needs planes p-2, p-1, p), four voxels per thread
            if (tid < NOQ && p >= 2 && p - 1 >= ou && p - 1 <= LU - 1 + ou) {
                const int pc = p - 1;
                const unsigned char *c0 = comp + rp2 * plane, *c1 = comp + rp1 * plane, *c2 = comp + rp * plane;
                const long long cu = u0 + pc;
                const unsigned mUlo = cu > 0 ? 0xFFFFFFFFu : 0u, mUhi = cu + 1 < G.n[0] ? 0xFFFFFFFFu : 0u;
                const int rowo = (oq_b + ov) * WP, col = oq_c + ow;
                const unsigned C = ld4(c1 + rowo, col);
                unsigned D = ((C ^ ld4(c0 + rowo, col)) & mUlo) | ((C ^ ld4(c2 + rowo, col)) & mUhi);
                D |= ((C ^ ld4(c1 + rowo - WP, col)) & mVlo) | ((C ^ ld4(c1 + rowo + WP, col)) & mVhi);
                D |= ((C ^ ld4(c1 + rowo, col - 1)) & mL) | ((C ^ ld4(c1 + rowo, col + 1)) & mR);
                unsigned flags = nz_bytes(D) & nz_bytes(C);
                if (EDGES) {  // the caller's mask decides (a flagged background centre is legal: its id is 0)
                    flags = 0u;
                    const long long ev = v0 + oq_b + ov;
                    if (ev < G.n[1]) {
                        const long long eb = cu * G.est[0] + ev * G.est[1];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const long long ew = w0 + oq_c + e + ow;
                            if (ew < G.n[2]) {
                                const long long a = eb + ew * G.est[2];
                                const bool on = G.edge_bytes == 4 ? (__ldg((const unsigned *)G.edges + a) != 0u)
                                                                  : (__ldg((const unsigned char *)G.edges + a) != 0);
                                flags |= on ? (0x80u << (8 * e)) : 0u;
                            }
                        }
                    }
                }
                reinterpret_cast<unsigned *>(cflag + (rf ? rf - 1 : CF - 1) * (TV * TW))[tid] = C | flags;
            }
            // C2. v-sums (4-bit fields) of the entering plane p and of the leaving plane p - su; their difference
            //     advances the running sum over the last su planes (8-bit fields)
            {
                const unsigned char *cn = comp + rp * plane;
                const unsigned char *co = comp + rpn * plane;  // plane p - su, since CR == su + 1
                const bool has_old = p >= su;
#pragma unroll 1
                for (int k = 0; k < MAXIT; ++k) {
                    if (tid + k * NT >= NG * (TV / 8) * WP) break;
                    const int g = itg[k];
                    const unsigned *lut = H.lut[g];
                    unsigned an[8], ao[8];
                    const unsigned gsh = 4u + 32u * (unsigned)g;
                    vsum8(cn + itoff[k], WP, sv, lut, gsh, an);
                    if (has_old) vsum8(co + itoff[k], WP, sv, lut, gsh, ao);
                    uint2 *sp = ssum + g * oplane + itoff[k];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const unsigned a = an[q], o = has_old ? ao[q] : 0u;
                        if (a != o) {
                            uint2 s2 = sp[q * WP];
                            s2.x += (a & 0x0F0F0F0Fu) - (o & 0x0F0F0F0Fu);
                            s2.y += ((a >> 4) & 0x0F0F0F0Fu) - ((o >> 4) & 0x0F0F0F0Fu);
                            sp[q * WP] = s2;
                        }
                    }
                }
            }
            if (!early_d) __syncthreads();  // the flags of plane p - ou were written in this very phase
            compact_outputs(p, rf);
            if (TMA) {  // raw plane p + 1 has landed (every thread observes the completion itself)
                if (p + 1 < NP) {
                    if ((p + 1) & 1) { mbar_wait(bar_sa + 8u, par1); par1 ^= 1u; }
                    else { mbar_wait(bar_sa, par0); par0 ^= 1u; }
                    pend = -1;
                }
            } else if (VEC4) cp_async_wait_all();  // this thread's part of raw plane p + 1 has landed; the barrier publishes all parts
            __syncthreads();
            // D. boundary voxels of plane uo = p - su + 1: final sum along w and arg-max
            const int uo = p - su + 1;
            if (uo >= 0 && u0 + uo < G.on[0]) {
                const unsigned char *cf = cflag + (rf + 1 == CF ? 0 : rf + 1) * (TV * TW);
                unsigned long long *orow = out + (u0 + uo) * G.ost[0];
                const int ne = H.n_edge;
#ifdef SYK_D_FORWARD
                for (int e = tid; e < ne; e += NT) {
#else
                for (int e = NT - 1 - tid; e < ne; e += NT) {  // from the last thread down: threads 0.. carry the second relabel quad
#endif
                    const int i = elist[e];
                    const int b = i / TW, c = i - b * TW;
                    const int jc = cf[i] & 0x7F;
                    const int gc = (jc - 1) >> 3, nc = (jc - 1) & 7;
                    unsigned best = 0u;
                    for (int g = 0; g < NG; ++g) {
                        const uint2 *sp = ssum + g * oplane + b * WP + c;
                        unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
                        int r = 0;
                        if (pair_ok) {  // add two ring sums in 8-bit fields first (2*su*sv <= 255), then widen
                            for (; r + 1 < sw; r += 2) {
                                const uint2 a = sp[r], d = sp[r + 1];
                                const unsigned l = a.x + d.x, h = a.y + d.y;
                                c0 += l & 0x00FF00FFu;
                                c1 += (l >> 8) & 0x00FF00FFu;
                                c2 += h & 0x00FF00FFu;
                                c3 += (h >> 8) & 0x00FF00FFu;
                            }
                        }
                        for (; r < sw; ++r) {
                            const uint2 a = sp[r];
                            c0 += a.x & 0x00FF00FFu;         // slots 0, 4
                            c1 += (a.x >> 8) & 0x00FF00FFu;  // slots 2, 6
                            c2 += a.y & 0x00FF00FFu;         // slots 1, 5
                            c3 += (a.y >> 8) & 0x00FF00FFu;  // slots 3, 7
                        }
                        if ((c0 | c1 | c2 | c3) == 0u) continue;
                        if (g == gc) {  // the centre id does not compete: clear its 16-bit field
                            const unsigned keep = (nc & 4) ? 0x0000FFFFu : 0xFFFF0000u;
                            const int r4 = nc & 3;
                            if (r4 == 0) c0 &= keep;
                            else if (r4 == 1) c2 &= keep;
                            else if (r4 == 2) c1 &= keep;
                            else c3 &= keep;
                        }
                        // keys (count << 16) | tie-break; slots that are free or absent have count 0 -> key < 65536
                        const uint4 t = reinterpret_cast<const uint4 *>(H.tb)[g];
                        const unsigned k0 = max(__byte_perm(t.x, c0, 0x5410), __byte_perm(t.x, c0, 0x7632));
                        const unsigned k1 = max(__byte_perm(t.y, c1, 0x5410), __byte_perm(t.y, c1, 0x7632));
                        const unsigned k2 = max(__byte_perm(t.z, c2, 0x5410), __byte_perm(t.z, c2, 0x7632));
                        const unsigned k3 = max(__byte_perm(t.w, c3, 0x5410), __byte_perm(t.w, c3, 0x7632));
                        best = max(best, max(max(k0, k1), max(k2, k3)));
                    }
                    unsigned long long res = 0ull;
                    if (best >> 16) {
                        const unsigned center = (!EDGES || jc) ? H.ids[jc - 1] : 0u, key = H.ids[best & 0xFFu];
                        res = center > key ? (((unsigned long long)key << 32) + center) : (((unsigned long long)center << 32) + key);
                    }
                    orow[(v0 + b) * G.ost[1] + (w0 + c) * G.ost[2]] = res;
                }
            }
            // no barrier here: the next iteration's pass 1 only touches the hash; its barrier orders everything else
        }
        if (aborted) {
            // the copy of the next plane must not land in the next segment's buffers
            if (TMA) {
                if (pend == 1) { mbar_wait(bar_sa + 8u, par1); par1 ^= 1u; }
                else if (pend == 0) { mbar_wait(bar_sa, par0); par0 ^= 1u; }
                pend = -1;
            } else if (VEC4) cp_async_wait_all();
            if (tid == 0) hard_list[atomicAdd(hard_count, 1u)] = (unsigned)seg;
        }
        wi = s_next;  // written before the segment's first barrier; the next write comes after the next segment's top barrier
    }
}

}  // namespace csfast
