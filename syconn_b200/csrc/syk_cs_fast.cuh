// syk_cs_fast.cuh -- fast path of the fused contact-site kernel (detect_seg_boundaries + process_block_nonzero).
//
// Exact reformulation of the reference's per-voxel window histogram (block_processing_C.pyx:21-49) as separable
// box sums of id-indicator volumes, evaluated only for the ids that actually occur near the tile:
//   * a CTA owns a TV x TW cross-section of the output and marches along the slowest axis (u) over a segment of LU
//     output planes; every input plane is read once per CTA (halo re-reads hit L2);
//   * the ids met on the way get compact slots (shared-memory hash, <= KMAX per segment); 8 ids share one 32-bit
//     word as 4-bit indicator fields, so one integer add advances 8 box sums at once;
//   * per plane: sliding sum along v (window sv <= 15 -> 4-bit fields), running sum over the last su planes
//     (<= 255 -> 8-bit fields, ring buffer of su planes in shared memory), and -- only for boundary voxels, which are
//     compacted first -- the final sum along w in 16-bit fields, arg-max with the reference tie-break (smallest id);
//   * segments that meet more than KMAX ids are appended to a device list and redone by the generic kernel.
// All arithmetic is integer; results are bit-identical to the generic path.
#pragma once
#include "syk_common.cuh"

namespace csfast {

constexpr int TV = 16;
constexpr int TW = 32;
constexpr int LU = 32;
constexpr int NT = 256;
constexpr int GMAX = 3;
constexpr int KMAX = GMAX * 8;
constexpr int HASH = 128;
constexpr int MAXQ = 2;   // relabel quads (4 voxels along w) per thread: ceil(VP*WP/4 / NT)
constexpr int MAXIT = 3;  // v-pass items per thread: ceil(GMAX*(TV/4)*WP / NT)

struct FastGeom {
    long long n[3];    // input extents (internal axes u, v, w)
    long long ist[3];  // input strides (elements)
    long long ost[3];  // output strides (elements)
    long long on[3];   // output extents
    int sten[3], off[3];
    int VP, WP;        // haloed plane dims: TV + sv - 1, TW + sw - 1 rounded up to a multiple of 4
    int CF;            // centre-flag ring depth
    long long segs[3];
    long long nsegs;
    int elem_bytes;
    int vec4;          // 16-byte aligned uint32 rows: quads are loaded with one LDG.128
};

struct FastSmem {  // offsets in bytes into dynamic shared memory
    int raw, comp, ind, ring, slo, shi, cflag, elist, total;
};

inline FastSmem fast_layout(const FastGeom &G) {
    FastSmem L;
    const int plane = G.VP * G.WP;
    const int oplane = TV * G.WP;
    int o = 0;
    auto take = [&](int bytes) { int at = o; o += (bytes + 15) & ~15; return at; };  // every region 16-byte aligned
    L.raw = take(plane * 4);
    L.comp = take(3 * plane);
    L.ind = take(GMAX * plane * 4);
    L.ring = take(GMAX * G.sten[0] * oplane * 4);   // ring, slo, shi stay contiguous (zeroed together)
    L.slo = take(GMAX * oplane * 4);
    L.shi = take(GMAX * oplane * 4);
    L.cflag = take(G.CF * TV * TW);
    L.elist = take(TV * TW * 2);
    L.total = o;
    return L;
}

struct Hash {
    unsigned keys[HASH];
    unsigned char slot[HASH];
    unsigned ids[KMAX];
    unsigned char rnk[KMAX];
    int nslots, newflag, n_edge;
};

// find the hash index of `lab` (!= 0), inserting it (and allocating a compact slot) when it is new
__device__ __forceinline__ int hash_find_insert(Hash &H, unsigned lab) {
    unsigned h = (lab * 2654435761u) >> 25;  // 7 bits
    for (int probes = 0; probes < HASH; ++probes) {
        const unsigned cur = H.keys[h];
        if (cur == lab) return (int)h;
        if (cur == 0u) {
            const unsigned prev = atomicCAS(&H.keys[h], 0u, lab);
            if (prev == 0u) {  // winner allocates the slot; others read it after the next barrier
                const int s = atomicAdd(&H.nslots, 1);
                if (s < KMAX) {
                    H.ids[s] = lab;
                    H.slot[h] = (unsigned char)s;
                }
                H.newflag = 1;
                return (int)h;
            }
            if (prev == lab) return (int)h;
        }
        h = (h + 1) & (HASH - 1);
    }
    atomicExch(&H.nslots, KMAX + 1);  // hash full => overflow
    return 0;
}

template <bool VEC4>
__global__ void __launch_bounds__(NT, 2)
k_cs_fast(const void *__restrict__ arr, unsigned long long *__restrict__ out, FastGeom G, FastSmem L,
          unsigned *__restrict__ hard_list, unsigned *__restrict__ hard_count) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned *raw = reinterpret_cast<unsigned *>(sm + L.raw);
    unsigned char *comp = sm + L.comp;
    unsigned *ind = reinterpret_cast<unsigned *>(sm + L.ind);
    unsigned *ring = reinterpret_cast<unsigned *>(sm + L.ring);
    unsigned *slo = reinterpret_cast<unsigned *>(sm + L.slo);
    unsigned *shi = reinterpret_cast<unsigned *>(sm + L.shi);
    unsigned char *cflag = sm + L.cflag;
    unsigned short *elist = reinterpret_cast<unsigned short *>(sm + L.elist);
    __shared__ Hash H;

    const int tid = threadIdx.x, lane = tid & 31;
    const int su = G.sten[0], sv = G.sten[1], sw = G.sten[2];
    const int ou = G.off[0], ov = G.off[1], ow = G.off[2];
    const int VP = G.VP, WP = G.WP;
    const int plane = VP * WP, oplane = TV * WP;
    const int nquad = plane >> 2, qpr = WP >> 2;  // quads per plane / per row
    const int NP = LU + su - 1;                   // input planes per segment

    for (long long seg = blockIdx.x; seg < G.nsegs; seg += gridDim.x) {
        const long long tw = seg % G.segs[2];
        const long long r0 = seg / G.segs[2];
        const long long tv = r0 % G.segs[1];
        const long long tu = r0 / G.segs[1];
        const long long u0 = tu * LU, v0 = tv * TV, w0 = tw * TW;  // output origin == input origin of the haloed block
        __syncthreads();
        // ---- segment init: zero rings / sums / hash ----
        {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 *p4 = reinterpret_cast<uint4 *>(sm + L.ring);
            const int n4 = (L.cflag - L.ring) / 16;  // ring, slo, shi are contiguous
            for (int i = tid; i < n4; i += NT) p4[i] = z;
            for (int i = tid; i < HASH; i += NT) {
                H.keys[i] = 0u;
                H.slot[i] = 0xFF;
            }
            if (tid == 0) {
                H.nslots = 0;
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
        int itoff[MAXIT], itg[MAXIT];  // v-pass items: offset inside a group plane (row vc*4, column c), group
#pragma unroll
        for (int k = 0; k < MAXIT; ++k) {
            const int it = tid + k * NT;
            const int c = it % WP;
            const int r = it / WP;
            itoff[k] = (r % (TV / 4)) * 4 * WP + c;
            itg[k] = r / (TV / 4);
        }
        // boundary-test invariants of my two output voxels (i = tid, tid + NT): which in-plane neighbours exist
        unsigned nbmask[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = tid + k * NT;
            const int b = i / TW, c = i - b * TW;
            const long long cv = v0 + b + ov, cw = w0 + c + ow;
            nbmask[k] = (cv > 0 ? 1u : 0u) | (cv + 1 < G.n[1] ? 2u : 0u) | (cw > 0 ? 4u : 0u) | (cw + 1 < G.n[2] ? 8u : 0u);
        }

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

        // prologue: plane 0 -> raw
#pragma unroll
        for (int k = 0; k < MAXQ; ++k) {
            const int q = tid + k * NT;
            if (q < nquad) {
                uint4 v;
                load_quad(k, u0, v);
                reinterpret_cast<uint4 *>(raw)[q] = v;
            }
        }
        __syncthreads();
        bool aborted = false;
        for (int p = 0; p < NP; ++p) {
            const long long gu = u0 + p;
            // A. prefetch plane p+1 into registers; relabel pass 1 (find / insert the key of every voxel of plane p)
            uint4 pre[MAXQ];
            int hidx[MAXQ][4];
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                pre[k] = make_uint4(0u, 0u, 0u, 0u);
                if (q < nquad && p + 1 < NP) load_quad(k, gu + 1, pre[k]);
            }
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                hidx[k][0] = hidx[k][1] = hidx[k][2] = hidx[k][3] = -1;
                if (q < nquad) {
                    const uint4 a = reinterpret_cast<const uint4 *>(raw)[q];
                    if (a.x != 0u) hidx[k][0] = hash_find_insert(H, a.x);
                    if (a.y != 0u) hidx[k][1] = (a.y == a.x) ? hidx[k][0] : hash_find_insert(H, a.y);
                    if (a.z != 0u) hidx[k][2] = (a.z == a.y) ? hidx[k][1] : hash_find_insert(H, a.z);
                    if (a.w != 0u) hidx[k][3] = (a.w == a.z) ? hidx[k][2] : hash_find_insert(H, a.w);
                }
            }
            __syncthreads();
            // B. slots are published: compact-index plane + indicator planes; next raw plane
            const int K = H.nslots;
            if (K > KMAX) { aborted = true; break; }
            if (H.newflag && tid < K) {  // ranks by id (tie-break of the arg-max: smallest id wins)
                const unsigned me = H.ids[tid];
                int r = 0;
                for (int s = 0; s < K; ++s) r += H.ids[s] < me;
                H.rnk[tid] = (unsigned char)r;
            }
            const int NG = (K + 7) >> 3;
            unsigned char *cp = comp + (p % 3) * plane;
#pragma unroll
            for (int k = 0; k < MAXQ; ++k) {
                const int q = tid + k * NT;
                if (q < nquad) {
                    int j[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) j[e] = hidx[k][e] < 0 ? 0 : (int)H.slot[hidx[k][e]] + 1;
                    reinterpret_cast<unsigned *>(cp)[q] = (unsigned)j[0] | ((unsigned)j[1] << 8) | ((unsigned)j[2] << 16) | ((unsigned)j[3] << 24);
                    for (int g = 0; g < NG; ++g) {
                        unsigned w4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int s = j[e] - 1;
                            w4[e] = (j[e] != 0 && (s >> 3) == g) ? (1u << ((s & 7) * 4)) : 0u;
                        }
                        reinterpret_cast<uint4 *>(ind + g * plane)[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                    reinterpret_cast<uint4 *>(raw)[q] = pre[k];
                }
            }
            __syncthreads();
            if (tid == 0) {
                H.newflag = 0;
                H.n_edge = 0;
            }
            // C1. boundary flags of plane p-1 (needs planes p-2, p-1, p)
            if (p >= 2 && p - 1 >= ou && p - 1 <= LU - 1 + ou) {
                const int pc = p - 1;
                const unsigned char *c0 = comp + ((p - 2) % 3) * plane, *c1 = comp + ((p - 1) % 3) * plane, *c2 = comp + (p % 3) * plane;
                unsigned char *cf = cflag + (pc % G.CF) * (TV * TW);
                const long long cu = u0 + pc;
                const bool has_lo = cu > 0, has_hi = cu + 1 < G.n[0];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int i = tid + k * NT;
                    const int b = i / TW, c = i - b * TW;
                    const int q = (b + ov) * WP + (c + ow);
                    const unsigned j = c1[q];
                    bool e = false;
                    if (j != 0u) {
                        e = (has_lo && c0[q] != j) || (has_hi && c2[q] != j) || ((nbmask[k] & 1u) && c1[q - WP] != j) ||
                            ((nbmask[k] & 2u) && c1[q + WP] != j) || ((nbmask[k] & 4u) && c1[q - 1] != j) ||
                            ((nbmask[k] & 8u) && c1[q + 1] != j);
                    }
                    cf[i] = (unsigned char)(j | (e ? 0x80u : 0u));
                }
            }
            // C2. v-pass (4-bit fields) + running sum over the last su planes (8-bit fields)
            {
                const int slot = p % su;
#pragma unroll
                for (int k = 0; k < MAXIT; ++k) {
                    if (tid + k * NT >= NG * (TV / 4) * WP) break;
                    const int g = itg[k];
                    const unsigned *ip = ind + g * plane + itoff[k];
                    unsigned acc = 0u;
                    for (int r = 0; r < sv; ++r) acc += ip[r * WP];
                    unsigned *rp = ring + (g * su + slot) * oplane + itoff[k];
                    unsigned *lo = slo + g * oplane + itoff[k];
                    unsigned *hi = shi + g * oplane + itoff[k];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q) acc += ip[(q + sv - 1) * WP] - ip[(q - 1) * WP];
                        const unsigned old = rp[q * WP];
                        rp[q * WP] = acc;
                        lo[q * WP] += (acc & 0x0F0F0F0Fu) - (old & 0x0F0F0F0Fu);
                        hi[q * WP] += ((acc >> 4) & 0x0F0F0F0Fu) - ((old >> 4) & 0x0F0F0F0Fu);
                    }
                }
            }
            __syncthreads();
            // D. outputs of plane uo = p - su + 1: zeros for non-boundary voxels, compacted list of boundary voxels
            const int uo = p - su + 1;
            if (uo >= 0 && u0 + uo < G.on[0]) {
                const unsigned char *cf = cflag + ((uo + ou) % G.CF) * (TV * TW);
                unsigned long long *orow = out + (u0 + uo) * G.ost[0];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int i = tid + k * NT;
                    const int b = i / TW, c = i - b * TW;
                    const long long gv = v0 + b, gw = w0 + c;
                    const bool inside = gv < G.on[1] && gw < G.on[2];
                    const bool e = inside && (cf[i] & 0x80u);
                    const unsigned m = __ballot_sync(0xffffffffu, e);
                    int base = 0;
                    if (lane == 0 && m) base = atomicAdd(&H.n_edge, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (e) elist[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
                    else if (inside) orow[gv * G.ost[1] + gw * G.ost[2]] = 0ull;
                }
                __syncthreads();
                const int ne = H.n_edge;
                for (int e = tid; e < ne; e += NT) {
                    const int i = elist[e];
                    const int b = i / TW, c = i - b * TW;
                    const int jc = cf[i] & 0x7F;
                    unsigned best = 0u;
                    for (int g = 0; g < NG; ++g) {
                        const unsigned *lo = slo + g * oplane + b * WP + c;
                        const unsigned *hi = shi + g * oplane + b * WP + c;
                        unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
                        for (int r = 0; r < sw; ++r) {
                            const unsigned l = lo[r], h = hi[r];
                            c0 += l & 0x00FF00FFu;         // slots 0, 4
                            c1 += (l >> 8) & 0x00FF00FFu;  // slots 2, 6
                            c2 += h & 0x00FF00FFu;         // slots 1, 5
                            c3 += (h >> 8) & 0x00FF00FFu;  // slots 3, 7
                        }
                        const unsigned cnt[8] = {c0 & 0xFFFFu, c2 & 0xFFFFu, c1 & 0xFFFFu, c3 & 0xFFFFu,
                                                 c0 >> 16,     c2 >> 16,     c1 >> 16,     c3 >> 16};
#pragma unroll
                        for (int n = 0; n < 8; ++n) {
                            const int s = g * 8 + n;
                            if (s < K && s + 1 != jc && cnt[n] != 0u) {
                                const unsigned key = (cnt[n] << 16) | ((255u - H.rnk[s]) << 8) | (unsigned)s;
                                best = key > best ? key : best;
                            }
                        }
                    }
                    unsigned long long res = 0ull;
                    if (best) {
                        const unsigned center = H.ids[jc - 1], key = H.ids[best & 0xFFu];
                        res = center > key ? (((unsigned long long)key << 32) + center) : (((unsigned long long)center << 32) + key);
                    }
                    orow[(v0 + b) * G.ost[1] + (w0 + c) * G.ost[2]] = res;
                }
            }
            // no barrier here: the next iteration's pass 1 only touches the hash; its barrier orders everything else
        }
        if (aborted) {
            if (tid == 0) hard_list[atomicAdd(hard_count, 1u)] = (unsigned)seg;
        }
    }
}

}  // namespace csfast
