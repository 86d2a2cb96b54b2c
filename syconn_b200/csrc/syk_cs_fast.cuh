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
constexpr int NT = 512;
constexpr int GMAX = 4;
constexpr int KMAX = GMAX * 8;
constexpr int HASH = 128;
constexpr int MAXPER = 4;  // relabel voxels per thread: ceil(VP*WP / NT)

struct FastGeom {
    long long n[3];    // input extents (internal axes u, v, w)
    long long ist[3];  // input strides (elements)
    long long ost[3];  // output strides (elements)
    long long on[3];   // output extents
    int sten[3], off[3];
    int VP, WP;        // haloed plane dims: TV + sv - 1, TW + sw - 1
    int CF;            // centre-flag ring depth
    long long segs[3];
    long long nsegs;
    int elem_bytes;
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

__device__ __forceinline__ unsigned ld_trunc(const void *base, int elem_bytes, long long idx) {
    return elem_bytes == 8 ? (unsigned)__ldg((const unsigned long long *)base + idx) : __ldg((const unsigned *)base + idx);
}

__global__ void __launch_bounds__(NT, 1)
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
    __shared__ unsigned hkeys[HASH];
    __shared__ unsigned char hslot[HASH];
    __shared__ unsigned ids[KMAX];
    __shared__ unsigned char rnk[KMAX];
    __shared__ int nslots, newflag, n_edge;

    const int tid = threadIdx.x, lane = tid & 31;
    const int su = G.sten[0], sv = G.sten[1], sw = G.sten[2];
    const int ou = G.off[0], ov = G.off[1], ow = G.off[2];
    const int VP = G.VP, WP = G.WP;
    const int plane = VP * WP, oplane = TV * WP;
    const int NP = LU + su - 1;  // input planes per segment

    for (long long seg = blockIdx.x; seg < G.nsegs; seg += gridDim.x) {
        const long long tw = seg % G.segs[2];
        const long long r0 = seg / G.segs[2];
        const long long tv = r0 % G.segs[1];
        const long long tu = r0 / G.segs[1];
        const long long u0 = tu * LU, v0 = tv * TV, w0 = tw * TW;  // output origin == input origin of the haloed block
        __syncthreads();
        // ---- segment init: zero rings / sums / hash ----
        {
            uint4 z = make_uint4(0u, 0u, 0u, 0u);
            uint4 *p4 = reinterpret_cast<uint4 *>(sm + L.ring);
            const int n4 = (L.cflag - L.ring) / 16;  // ring, slo, shi are contiguous
            for (int i = tid; i < n4; i += NT) p4[i] = z;
            for (int i = tid; i < HASH; i += NT) {
                hkeys[i] = 0u;
                hslot[i] = 0xFF;
            }
            if (tid == 0) {
                nslots = 0;
                newflag = 0;
            }
        }
        // prologue: plane 0 -> raw
        for (int i = tid; i < plane; i += NT) {
            const int b = i / WP, c = i - b * WP;
            const long long gv = v0 + b, gw = w0 + c;
            unsigned v = 0u;
            if (u0 < G.n[0] && gv < G.n[1] && gw < G.n[2]) v = ld_trunc(arr, G.elem_bytes, u0 * G.ist[0] + gv * G.ist[1] + gw * G.ist[2]);
            raw[i] = v;
        }
        __syncthreads();
        bool aborted = false;
        for (int p = 0; p < NP; ++p) {
            const long long gu = u0 + p;
            // 1. prefetch plane p+1 into registers
            unsigned pre[MAXPER];
#pragma unroll
            for (int k = 0; k < MAXPER; ++k) {
                const int i = tid + k * NT;
                pre[k] = 0u;
                if (i < plane && p + 1 < NP) {
                    const int b = i / WP, c = i - b * WP;
                    const long long gv = v0 + b, gw = w0 + c;
                    if (gu + 1 < G.n[0] && gv < G.n[1] && gw < G.n[2])
                        pre[k] = ld_trunc(arr, G.elem_bytes, (gu + 1) * G.ist[0] + gv * G.ist[1] + gw * G.ist[2]);
                }
            }
            // 2. relabel pass 1: find or insert the key of every voxel of plane p
            int hidx[MAXPER];
#pragma unroll
            for (int k = 0; k < MAXPER; ++k) {
                const int i = tid + k * NT;
                hidx[k] = -1;
                if (i < plane) {
                    const unsigned lab = raw[i];
                    if (lab != 0u) {
                        unsigned h = (lab * 2654435761u) >> 25;  // 7 bits
                        for (int probes = 0; probes < HASH; ++probes) {
                            const unsigned cur = hkeys[h];
                            if (cur == lab) { hidx[k] = (int)h; break; }
                            if (cur == 0u) {
                                const unsigned prev = atomicCAS(&hkeys[h], 0u, lab);
                                if (prev == 0u || prev == lab) { hidx[k] = (int)h; break; }
                            }
                            h = (h + 1) & (HASH - 1);
                        }
                        if (hidx[k] < 0) atomicExch(&nslots, KMAX + 1);  // hash full => overflow
                    }
                }
            }
            __syncthreads();
            // 3. slots for new keys
            if (tid < HASH && hkeys[tid] != 0u && hslot[tid] == 0xFF) {
                const int s = atomicAdd(&nslots, 1);
                if (s < KMAX) {
                    hslot[tid] = (unsigned char)s;
                    ids[s] = hkeys[tid];
                }
                newflag = 1;
            }
            __syncthreads();
            const int K = nslots;
            if (K > KMAX) { aborted = true; break; }
            if (newflag) {  // ranks by id (tie-break of the arg-max: smallest id wins)
                if (tid < K) {
                    const unsigned me = ids[tid];
                    int r = 0;
                    for (int s = 0; s < K; ++s) r += ids[s] < me;
                    rnk[tid] = (unsigned char)r;
                }
            }
            const int NG = (K + 7) >> 3;
            // 5. relabel pass 2: compact index plane + indicator planes
            unsigned char *cp = comp + (p % 3) * plane;
#pragma unroll
            for (int k = 0; k < MAXPER; ++k) {
                const int i = tid + k * NT;
                if (i < plane) {
                    const int j = hidx[k] < 0 ? 0 : (int)hslot[hidx[k]] + 1;
                    cp[i] = (unsigned char)j;
                    const int s = j - 1;
                    for (int g = 0; g < NG; ++g) ind[g * plane + i] = (j != 0 && (s >> 3) == g) ? (1u << ((s & 7) * 4)) : 0u;
                }
            }
            // 6. next raw plane
#pragma unroll
            for (int k = 0; k < MAXPER; ++k) {
                const int i = tid + k * NT;
                if (i < plane) raw[i] = pre[k];
            }
            __syncthreads();
            if (tid == 0) newflag = 0;
            // 8. boundary flags of plane p-1 (needs planes p-2, p-1, p)
            if (p >= 2) {
                const int pc = p - 1;
                if (pc >= ou && pc <= LU - 1 + ou) {
                    const unsigned char *c0 = comp + ((p - 2) % 3) * plane, *c1 = comp + ((p - 1) % 3) * plane, *c2 = comp + (p % 3) * plane;
                    unsigned char *cf = cflag + (pc % G.CF) * (TV * TW);
                    const long long cu = u0 + pc;
                    for (int i = tid; i < TV * TW; i += NT) {
                        const int b = i / TW, c = i - b * TW;
                        const int q = (b + ov) * WP + (c + ow);
                        const unsigned j = c1[q];
                        bool e = false;
                        if (j != 0u) {
                            const long long cv = v0 + b + ov, cw = w0 + c + ow;
                            if (cu > 0 && c0[q] != j) e = true;
                            if (cu + 1 < G.n[0] && c2[q] != j) e = true;
                            if (cv > 0 && c1[q - WP] != j) e = true;
                            if (cv + 1 < G.n[1] && c1[q + WP] != j) e = true;
                            if (cw > 0 && c1[q - 1] != j) e = true;
                            if (cw + 1 < G.n[2] && c1[q + 1] != j) e = true;
                        }
                        cf[i] = (unsigned char)(j | (e ? 0x80u : 0u));
                    }
                }
            }
            // 9. v-pass (4-bit fields) + running sum over the last su planes (8-bit fields)
            {
                const int slot = p % su;
                const int nitem = NG * (TV / 4) * WP;
                for (int it = tid; it < nitem; it += NT) {
                    const int c = it % WP;
                    const int r = it / WP;
                    const int vc = r % (TV / 4);
                    const int g = r / (TV / 4);
                    const unsigned *ip = ind + g * plane + (vc * 4) * WP + c;
                    unsigned acc = 0u;
                    for (int k = 0; k < sv; ++k) acc += ip[k * WP];
                    unsigned *rp = ring + (g * su + slot) * oplane + (vc * 4) * WP + c;
                    unsigned *lo = slo + g * oplane + (vc * 4) * WP + c;
                    unsigned *hi = shi + g * oplane + (vc * 4) * WP + c;
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
            if (tid == 0) n_edge = 0;
            __syncthreads();
            // 11. outputs of plane uo = p - su + 1
            const int uo = p - su + 1;
            if (uo >= 0 && u0 + uo < G.on[0]) {
                const unsigned char *cf = cflag + ((uo + ou) % G.CF) * (TV * TW);
                for (int i = tid; i < TV * TW; i += NT) {
                    const int b = i / TW, c = i - b * TW;
                    const long long gv = v0 + b, gw = w0 + c;
                    const bool inside = gv < G.on[1] && gw < G.on[2];
                    const bool e = inside && (cf[i] & 0x80u);
                    const unsigned m = __ballot_sync(0xffffffffu, e);
                    int base = 0;
                    if (lane == 0 && m) base = atomicAdd(&n_edge, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (e) elist[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
                    else if (inside) out[(u0 + uo) * G.ost[0] + gv * G.ost[1] + gw * G.ost[2]] = 0ull;
                }
                __syncthreads();
                const int ne = n_edge;
                for (int e = tid; e < ne; e += NT) {
                    const int i = elist[e];
                    const int b = i / TW, c = i - b * TW;
                    const int jc = cf[i] & 0x7F;
                    unsigned best = 0u;
                    for (int g = 0; g < NG; ++g) {
                        const unsigned *lo = slo + g * oplane + b * WP + c;
                        const unsigned *hi = shi + g * oplane + b * WP + c;
                        unsigned c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
                        for (int k = 0; k < sw; ++k) {
                            const unsigned l = lo[k], h = hi[k];
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
                                const unsigned key = (cnt[n] << 16) | ((255u - rnk[s]) << 8) | (unsigned)s;
                                best = key > best ? key : best;
                            }
                        }
                    }
                    unsigned long long res = 0ull;
                    if (best) {
                        const unsigned center = ids[jc - 1], key = ids[best & 0xFFu];
                        res = center > key ? (((unsigned long long)key << 32) + center) : (((unsigned long long)center << 32) + key);
                    }
                    out[(u0 + uo) * G.ost[0] + (v0 + b) * G.ost[1] + (w0 + c) * G.ost[2]] = res;
                }
            }
            __syncthreads();
        }
        if (aborted) {
            if (tid == 0) hard_list[atomicAdd(hard_count, 1u)] = (unsigned)seg;
        }
    }
}

}  // namespace csfast
