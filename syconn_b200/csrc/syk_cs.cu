// syk_cs.cu -- contact-site detection (sm_100a).
//
//   syk_detect_seg_boundaries  <- syconn/extraction/find_object_properties.py:424-455
//   syk_process_block_nonzero  <- syconn/extraction/block_processing_C.pyx:53-75 (+ kernel :21-49)
//   syk_detect_cs              <- syconn/extraction/find_object_properties.py:458-472 (fused, no edge volume in HBM)
//
// Generic path (this file): haloed 3-D tile of uint32 ids staged in shared memory; boundary voxels of the tile are
// compacted into a list; one warp per boundary voxel sweeps the stencil window 32 neighbours at a time,
// de-duplicates them with __match_any_sync and keeps a warp-distributed histogram (lane i owns the i-th distinct
// id).  Windows with more than 32 distinct ids fall back to a block-cooperative shared-memory hash table.
// The arg-max follows the reference exactly: ids 0 and centre excluded, ties -> smallest id.
#include <stdlib.h>
#include <string.h>

#include "syk_common.cuh"
#include "syk_cs_fast.cuh"
#include "syk_cs_march.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CS_THREADS = 128;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int OU = 4, OV = 8, OW = 32;  // output tile (internal axes u, v, w); w = lane axis
constexpr int OT = OU * OV * OW;

struct CsGeom {
    long long n[3];    // input extents along internal axes
    long long ist[3];  // input strides (elements)
    long long est[3];  // edge-volume strides (elements), only with explicit edges
    long long ost[3];  // output strides (elements)
    long long on[3];   // output extents = n - sten + 1
    int sten[3];       // stencil along internal axes
    int off[3];        // sten / 2
    int hlo[3];        // smem halo below the output tile: max(off, 1)
    int hd[3];         // smem tile dims
    int total;         // sten[0]*sten[1]*sten[2]
    int hslots;        // fallback hash slots (power of two >= 2*total)
    long long tiles[3];
    long long ntiles;
    int elem_bytes;
    int edge_bytes;
    // list mode (fallback of the fast path): only the segments named in seg_list are processed
    const unsigned *seg_list;
    const unsigned *seg_count;
    long long segs[3];  // segment grid of the fast path
    int seg_tiles[3];   // tiles per segment along u, v, w
    // first-seen mode (detect_contact_partners, find_object_properties.py:371-421): ties of the arg-max go to the id met
    // first when the window is scanned in LOGICAL x, y, z order, and the output is (centre << 32) | partner, unordered
    int first_seen;
    int la[3];          // logical axis of internal axis u, v, w
    int lsten[3];       // stencil along the logical axes
};

__device__ __forceinline__ unsigned ld_id(const void *base, int elem_bytes, long long idx) {
    return elem_bytes == 8 ? (unsigned)__ldg((const unsigned long long *)base + idx) : __ldg((const unsigned *)base + idx);
}

__device__ __forceinline__ unsigned long long pack_result(unsigned center, unsigned key, unsigned best) {
    if (best == 0u) return 0ull;
    return center > key ? (((unsigned long long)key << 32) + center) : (((unsigned long long)center << 32) + key);
}

__global__ void __launch_bounds__(CS_THREADS) k_detect_cs(const void *__restrict__ arr, const void *__restrict__ edges,
                                                          unsigned long long *__restrict__ out, CsGeom G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *tile = reinterpret_cast<unsigned *>(smem_raw);                       // hd[0]*hd[1]*hd[2]
    const int tile_n = G.hd[0] * G.hd[1] * G.hd[2];
    unsigned short *offs = reinterpret_cast<unsigned short *>(tile + tile_n);       // total (padded to x2)
    unsigned short *lrank = offs + ((G.total + 1) & ~1);                            // total (padded), first-seen mode only
    unsigned short *elist = lrank + (G.first_seen ? ((G.total + 1) & ~1) : 0);      // OT
    unsigned short *olist = elist + OT;                                             // OT (overflow voxels)
    unsigned *hkeys = reinterpret_cast<unsigned *>(olist + OT);                     // hslots
    unsigned *hcnt = hkeys + G.hslots;                                              // hslots
    unsigned *hfirst = hcnt + G.hslots;                                             // hslots, first-seen mode only
    __shared__ int n_edge, n_over;
    __shared__ unsigned long long best_sh;

    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    // window offsets inside the smem tile, relative to the window origin
    for (int i = tid; i < G.total; i += CS_THREADS) {
        const int k = i % G.sten[2];
        const int r = i / G.sten[2];
        const int j = r % G.sten[1];
        const int ii = r / G.sten[1];
        offs[i] = (unsigned short)((ii * G.hd[1] + j) * G.hd[2] + k);
        if (G.first_seen) {  // rank of the window voxel in the reference's x, y, z loop nest
            int l[3];
            l[G.la[0]] = ii;
            l[G.la[1]] = j;
            l[G.la[2]] = k;
            lrank[i] = (unsigned short)((l[0] * G.lsten[1] + l[1]) * G.lsten[2] + l[2]);
        }
    }
    for (int i = tid; i < G.hslots; i += CS_THREADS) {
        hkeys[i] = 0u;
        hcnt[i] = 0u;
        if (G.first_seen) hfirst[i] = 0xFFFFFFFFu;
    }

    const int tps = G.seg_tiles[0] * G.seg_tiles[1] * G.seg_tiles[2];
    const long long ntiles = G.seg_list ? (long long)(*G.seg_count) * tps : G.ntiles;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        long long tu, tv, tw;
        if (G.seg_list) {  // tile t = sub-tile (t % tps) of listed segment t / tps
            const long long seg = G.seg_list[t / tps];
            int sub = (int)(t % tps);
            const int sw_ = sub % G.seg_tiles[2];
            sub /= G.seg_tiles[2];
            const int sv_ = sub % G.seg_tiles[1];
            const int su_ = sub / G.seg_tiles[1];
            const long long gw = seg % G.segs[2];
            const long long r1 = seg / G.segs[2];
            tw = gw * G.seg_tiles[2] + sw_;
            tv = (r1 % G.segs[1]) * G.seg_tiles[1] + sv_;
            tu = (r1 / G.segs[1]) * G.seg_tiles[0] + su_;
            if (tu * OU >= G.on[0] || tv * OV >= G.on[1] || tw * OW >= G.on[2]) continue;
        } else {
            tw = t % G.tiles[2];
            const long long r0 = t / G.tiles[2];
            tv = r0 % G.tiles[1];
            tu = r0 / G.tiles[1];
        }
        const long long o0[3] = {tu * OU, tv * OV, tw * OW};  // output-tile origin (output coords)
        // input coords of smem tile origin: output coord + off - hlo
        const long long i0[3] = {o0[0] + G.off[0] - G.hlo[0], o0[1] + G.off[1] - G.hlo[1], o0[2] + G.off[2] - G.hlo[2]};
        __syncthreads();
        if (tid == 0) {
            n_edge = 0;
            n_over = 0;
        }
        // ---- stage the haloed tile (zero fill outside the volume; never read by valid windows) ----
        for (int i = tid; i < tile_n; i += CS_THREADS) {
            const int k = i % G.hd[2];
            const int r = i / G.hd[2];
            const int j = r % G.hd[1];
            const int ii = r / G.hd[1];
            const long long a = i0[0] + ii, b = i0[1] + j, c = i0[2] + k;
            unsigned v = 0u;
            if (a >= 0 && a < G.n[0] && b >= 0 && b < G.n[1] && c >= 0 && c < G.n[2])
                v = ld_id(arr, G.elem_bytes, a * G.ist[0] + b * G.ist[1] + c * G.ist[2]);
            tile[i] = v;
        }
        __syncthreads();
        // ---- boundary mask (detect_seg_boundaries) + zero fill of non-boundary outputs ----
        for (int i = tid; i < OT; i += CS_THREADS) {
            const int lw = i % OW;
            const int r = i / OW;
            const int lv = r % OV;
            const int lu = r / OV;
            const long long ou = o0[0] + lu, ov = o0[1] + lv, ow = o0[2] + lw;
            if (ou >= G.on[0] || ov >= G.on[1] || ow >= G.on[2]) continue;
            const int ci = ((lu + G.hlo[0]) * G.hd[1] + (lv + G.hlo[1])) * G.hd[2] + (lw + G.hlo[2]);
            const unsigned c = tile[ci];
            bool edge;
            if (edges != nullptr) {
                const long long ei = (ou + G.off[0]) * G.est[0] + (ov + G.off[1]) * G.est[1] + (ow + G.off[2]) * G.est[2];
                edge = G.edge_bytes == 1 ? (__ldg((const unsigned char *)edges + ei) != 0)
                                         : (__ldg((const unsigned *)edges + ei) != 0u);
            } else {
                edge = false;
                if (c != 0u) {
                    const long long cu = ou + G.off[0], cv = ov + G.off[1], cw = ow + G.off[2];  // input coords of the centre
                    const int su = G.hd[1] * G.hd[2], sv = G.hd[2];
                    if (cu > 0 && tile[ci - su] != c) edge = true;
                    if (cu + 1 < G.n[0] && tile[ci + su] != c) edge = true;
                    if (cv > 0 && tile[ci - sv] != c) edge = true;
                    if (cv + 1 < G.n[1] && tile[ci + sv] != c) edge = true;
                    if (cw > 0 && tile[ci - 1] != c) edge = true;
                    if (cw + 1 < G.n[2] && tile[ci + 1] != c) edge = true;
                }
            }
            if (edge) {
                elist[atomicAdd(&n_edge, 1)] = (unsigned short)i;
            } else {
                out[ou * G.ost[0] + ov * G.ost[1] + ow * G.ost[2]] = 0ull;
            }
        }
        __syncthreads();
        // ---- one warp per boundary voxel ----
        const int ne = n_edge;
        for (int e = wib; e < ne; e += CS_WARPS) {
            const int i = elist[e];
            const int lw = i % OW;
            const int r = i / OW;
            const int lv = r % OV;
            const int lu = r / OV;
            // window origin inside the smem tile
            const int wbase = ((lu + G.hlo[0] - G.off[0]) * G.hd[1] + (lv + G.hlo[1] - G.off[1])) * G.hd[2] + (lw + G.hlo[2] - G.off[2]);
            const unsigned center = tile[((lu + G.hlo[0]) * G.hd[1] + (lv + G.hlo[1])) * G.hd[2] + (lw + G.hlo[2])];
            unsigned my_key = 0u, my_cnt = 0u, my_first = 0xFFFFu;
            int nheld = 0;
            bool overflow = false;
            for (int base = 0; base < G.total && !overflow; base += 32) {
                const int idx = base + lane;
                unsigned id = 0u;
                if (idx < G.total) id = tile[wbase + offs[idx]];
                if (id == center) id = 0u;
                if (!__any_sync(FULL, id != 0u)) continue;
                const unsigned peers = __match_any_sync(FULL, id);
                const bool leader = (id != 0u) && (lane == __ffs(peers) - 1);
                const unsigned n = (unsigned)__popc(peers);
                unsigned rk = 0xFFFFu;  // first-seen mode: smallest logical rank of this id in the block of 32
                if (G.first_seen) rk = __reduce_min_sync(peers, idx < G.total ? (unsigned)lrank[idx] : 0xFFFFu);
                unsigned todo = __ballot_sync(FULL, leader);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const unsigned kv = __shfl_sync(FULL, id, src);
                    const unsigned kn = __shfl_sync(FULL, n, src);
                    const unsigned kr = __shfl_sync(FULL, rk, src);
                    const unsigned hit = __ballot_sync(FULL, lane < nheld && my_key == kv);
                    if (hit) {
                        if (lane == __ffs(hit) - 1) {
                            my_cnt += kn;
                            my_first = min(my_first, kr);
                        }
                    } else if (nheld < 32) {
                        if (lane == nheld) {
                            my_key = kv;
                            my_cnt = kn;
                            my_first = kr;
                        }
                        ++nheld;
                    } else {
                        overflow = true;
                        break;
                    }
                }
            }
            if (overflow) {
                if (lane == 0) olist[atomicAdd(&n_over, 1)] = (unsigned short)i;
                continue;
            }
            // arg-max: larger count wins, ties -> smaller id (first-seen mode: -> smaller logical rank; window <= 16384
            // voxels, so count takes 15 bits and the rank 14)
            unsigned long long best = 0ull;
            if (lane < nheld)
                best = G.first_seen ? (((unsigned long long)my_cnt << 46) | ((unsigned long long)(0x3FFFu - my_first) << 32) | my_key)
                                    : (((unsigned long long)my_cnt << 32) | (unsigned long long)(~my_key));
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(FULL, best, o);
                best = other > best ? other : best;
            }
            if (lane == 0) {
                unsigned long long res;
                if (G.first_seen) {
                    res = best ? (((unsigned long long)center << 32) | (best & 0xffffffffull)) : 0ull;
                } else {
                    const unsigned bc = (unsigned)(best >> 32);
                    const unsigned bk = ~(unsigned)(best & 0xffffffffull);
                    res = pack_result(center, bk, bc);
                }
                out[(o0[0] + lu) * G.ost[0] + (o0[1] + lv) * G.ost[1] + (o0[2] + lw) * G.ost[2]] = res;
            }
        }
        __syncthreads();
        // ---- windows with > 32 distinct ids: block-cooperative shared-memory hash ----
        const int no = n_over;
        for (int e = 0; e < no; ++e) {
            const int i = olist[e];
            const int lw = i % OW;
            const int r = i / OW;
            const int lv = r % OV;
            const int lu = r / OV;
            const int wbase = ((lu + G.hlo[0] - G.off[0]) * G.hd[1] + (lv + G.hlo[1] - G.off[1])) * G.hd[2] + (lw + G.hlo[2] - G.off[2]);
            const unsigned center = tile[((lu + G.hlo[0]) * G.hd[1] + (lv + G.hlo[1])) * G.hd[2] + (lw + G.hlo[2])];
            if (tid == 0) best_sh = 0ull;
            __syncthreads();
            const unsigned hm = (unsigned)G.hslots - 1u;
            for (int idx = tid; idx < G.total; idx += CS_THREADS) {
                const unsigned id = tile[wbase + offs[idx]];
                if (id == 0u || id == center) continue;
                unsigned h = syk_mix32(id) & hm;
                for (;;) {
                    const unsigned prev = atomicCAS(&hkeys[h], 0u, id);
                    if (prev == 0u || prev == id) {
                        atomicAdd(&hcnt[h], 1u);
                        if (G.first_seen) atomicMin(&hfirst[h], (unsigned)lrank[idx]);
                        break;
                    }
                    h = (h + 1u) & hm;
                }
            }
            __syncthreads();
            unsigned long long best = 0ull;
            for (int h = tid; h < G.hslots; h += CS_THREADS) {
                const unsigned k = hkeys[h];
                if (k != 0u) {
                    const unsigned long long cand =
                        G.first_seen ? (((unsigned long long)hcnt[h] << 46) | ((unsigned long long)(0x3FFFu - hfirst[h]) << 32) | k)
                                     : (((unsigned long long)hcnt[h] << 32) | (unsigned long long)(~k));
                    best = cand > best ? cand : best;
                    hkeys[h] = 0u;
                    hcnt[h] = 0u;
                    if (G.first_seen) hfirst[h] = 0xFFFFFFFFu;
                }
            }
            if (best) atomicMax(&best_sh, best);
            __syncthreads();
            if (tid == 0) {
                const unsigned long long b = best_sh;
                unsigned long long res;
                if (G.first_seen) {
                    res = b ? (((unsigned long long)center << 32) | (b & 0xffffffffull)) : 0ull;
                } else {
                    const unsigned bc = (unsigned)(b >> 32);
                    const unsigned bk = ~(unsigned)(b & 0xffffffffull);
                    res = pack_result(center, bk, bc);
                }
                out[(o0[0] + lu) * G.ost[0] + (o0[1] + lv) * G.ost[1] + (o0[2] + lw) * G.ost[2]] = res;
            }
            __syncthreads();
        }
    }
}

__global__ void k_seg_boundaries(const void *__restrict__ arr, int elem_bytes, long long nx, long long ny, long long nz,
                                 long long sx, long long sy, long long sz, unsigned char *__restrict__ out) {
    const long long total = nx * ny * nz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long z = i % nz;
        const long long r = i / nz;
        const long long y = r % ny;
        const long long x = r / ny;
        const long long a = x * sx + y * sy + z * sz;
        unsigned char b = 0;
        if (elem_bytes == 8) {
            const unsigned long long *p = (const unsigned long long *)arr;
            const unsigned long long c = __ldg(p + a);
            if (c != 0ull) {
                if (x > 0 && __ldg(p + a - sx) != c) b = 1;
                if (x + 1 < nx && __ldg(p + a + sx) != c) b = 1;
                if (y > 0 && __ldg(p + a - sy) != c) b = 1;
                if (y + 1 < ny && __ldg(p + a + sy) != c) b = 1;
                if (z > 0 && __ldg(p + a - sz) != c) b = 1;
                if (z + 1 < nz && __ldg(p + a + sz) != c) b = 1;
            }
        } else {
            const unsigned *p = (const unsigned *)arr;
            const unsigned c = __ldg(p + a);
            if (c != 0u) {
                if (x > 0 && __ldg(p + a - sx) != c) b = 1;
                if (x + 1 < nx && __ldg(p + a + sx) != c) b = 1;
                if (y > 0 && __ldg(p + a - sy) != c) b = 1;
                if (y + 1 < ny && __ldg(p + a + sy) != c) b = 1;
                if (z > 0 && __ldg(p + a - sz) != c) b = 1;
                if (z + 1 < nz && __ldg(p + a + sz) != c) b = 1;
            }
        }
        out[i] = b;
    }
}

// internal axes: w = axis with the smallest |input stride| (lanes / coalescing), u = largest
static void cs_plan(const int64_t shape[3], const int64_t strides[3], const int32_t stencil[3], const int64_t *edge_strides,
                    const int64_t out_strides[3], int elem_bytes, int edge_bytes, CsGeom &G) {
    int ax[3] = {0, 1, 2};
    auto key = [&](int a) { return strides[a] < 0 ? -strides[a] : strides[a]; };
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (key(ax[j]) > key(ax[i])) {
                int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    // The box-sum kernels keep the sums along v in 4-bit fields (window <= 15) and those along u, v in 8-bit fields.  Only w
    // is tied to the memory layout: when the stencil along the natural v axis is too wide but the one along u is not
    // (e.g. 17 x 17 x 9 on x-fastest data: u = z (9), v = y (17)), march along the other axis instead.
    if (stencil[ax[1]] > 15 && stencil[ax[0]] <= 15 && stencil[ax[0]] >= 3) {
        const int t = ax[0];
        ax[0] = ax[1];
        ax[1] = t;
    }
    G.total = 1;
    for (int a = 0; a < 3; ++a) {
        const int l = ax[a];
        G.n[a] = shape[l];
        G.ist[a] = strides[l];
        G.est[a] = edge_strides ? edge_strides[l] : 0;
        G.ost[a] = out_strides[l];
        G.sten[a] = stencil[l];
        G.off[a] = stencil[l] / 2;
        G.on[a] = shape[l] - stencil[l] + 1;
        G.hlo[a] = G.off[a] > 1 ? G.off[a] : 1;
        G.total *= stencil[l];
        G.la[a] = l;
        G.lsten[a] = stencil[a];
    }
    G.first_seen = 0;
    const int od[3] = {OU, OV, OW};
    for (int a = 0; a < 3; ++a) {
        G.hd[a] = od[a] + 2 * G.hlo[a];
        G.tiles[a] = (G.on[a] + od[a] - 1) / od[a];
    }
    G.ntiles = G.tiles[0] * G.tiles[1] * G.tiles[2];
    int hs = 64;
    while (hs < 2 * G.total) hs <<= 1;
    G.hslots = hs;
    G.elem_bytes = elem_bytes;
    G.edge_bytes = edge_bytes;
    G.seg_list = nullptr;
    G.seg_count = nullptr;
    G.segs[0] = G.segs[1] = G.segs[2] = 1;
    G.seg_tiles[0] = G.seg_tiles[1] = G.seg_tiles[2] = 1;
}

// The fast path needs: every stencil dim >= 3 (the boundary test of a centre plane uses its two neighbour planes),
// 4-bit fields for the v-sum (sv <= 15), 8-bit fields for the plane ring sum (su*sv <= 255), 16-bit totals, and the
// haloed plane / rings must fit the CTA's shared memory.
static bool fast_plan(const CsGeom &C, csfast::FastGeom &F) {
    using namespace csfast;
    for (int a = 0; a < 3; ++a)
        if (C.sten[a] < 3) return false;
    if (C.sten[1] > 15 || C.sten[0] * C.sten[1] > 255 || C.total > 65535) return false;
    for (int a = 0; a < 3; ++a) {
        F.n[a] = C.n[a];
        F.ist[a] = C.ist[a];
        F.ost[a] = C.ost[a];
        F.on[a] = C.on[a];
        F.sten[a] = C.sten[a];
        F.off[a] = C.off[a];
    }
    F.VP = TV + C.sten[1] - 1;
    F.WP = (TW + C.sten[2] - 1 + 3) & ~3;
    F.CF = C.off[0] + 1;
    F.CR = C.sten[0] + 1;
    F.segs[0] = (C.on[0] + LU - 1) / LU;
    F.segs[1] = (C.on[1] + TV - 1) / TV;
    F.segs[2] = (C.on[2] + TW - 1) / TW;
    F.nsegs = F.segs[0] * F.segs[1] * F.segs[2];
    F.elem_bytes = C.elem_bytes;
    F.vec4 = 0;
    F.tma = 0;
    F.edges = nullptr;
    F.edge_bytes = 0;
    for (int a = 0; a < 3; ++a) F.est[a] = 0;
    F.pair_ok = (2 * C.sten[0] * C.sten[1] <= 255) ? 1 : 0;
    F.out_vec = 0;
    if (F.VP * F.WP / 4 > MAX_NQUAD || F.WP > MAX_WP || F.nsegs >= (1ll << 30)) return false;
    const csfast::FastSmem L = fast_layout(F, GMAX_T2);
    return L.total <= 200 * 1024;
}

static int cs_launch_impl(const void *edges, int edge_bytes, const int64_t *edge_strides, const void *arr, int elem_bytes,
                          const int64_t strides[3], const int64_t shape[3], const int32_t stencil[3], uint64_t *out,
                          const int64_t out_strides[3], cudaStream_t s, int first_seen) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(shape && strides && stencil && out_strides, "NULL geometry argument");
    for (int a = 0; a < 3; ++a) {
        SYK_CHECK_ARG(stencil[a] >= 1 && (stencil[a] % 2) == 1, "stencil must be odd along every axis (block_processing_C.pyx:57)");
        SYK_CHECK_ARG(shape[a] >= 0, "negative shape");
    }
    SYK_CHECK_ARG((long long)stencil[0] * stencil[1] * stencil[2] <= 16384, "stencil window larger than 16384 voxels");
    for (int a = 0; a < 3; ++a)
        if (shape[a] - stencil[a] + 1 <= 0) return SYK_OK;  // empty output
    SYK_CHECK_ARG(arr != nullptr && out != nullptr, "arr/out is NULL");
    CsGeom G;
    cs_plan(shape, strides, stencil, edge_strides, out_strides, elem_bytes, edge_bytes, G);
    const size_t tile_n = (size_t)G.hd[0] * G.hd[1] * G.hd[2];
    SYK_CHECK_ARG(tile_n < 65536, "stencil too large for the shared-memory tile");
    G.first_seen = first_seen;
    const size_t smem = tile_n * 4 + (size_t)((G.total + 1) & ~1) * 2 * (first_seen ? 2 : 1) + (size_t)OT * 2 * 2 +
                        (size_t)G.hslots * (first_seen ? 12 : 8);
    SYK_CHECK_ARG(smem <= 220 * 1024, "stencil too large for shared memory");
    if ((rc = syk_ensure_dyn_smem((const void *)k_detect_cs, (int)smem))) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int bps = (int)((220 * 1024) / (smem + 1024));
    if (bps < 1) bps = 1;
    if (bps > 8) bps = 8;
    long long grid = (long long)sms * bps;
    if (grid > G.ntiles) grid = G.ntiles;
    csfast::FastGeom F;
    if (!first_seen && !getenv("SYK_CS_GENERIC") && fast_plan(G, F)) {
        using namespace csfast;
        if (edges != nullptr) {  // explicit edge mask: same kernels, the mask replaces the fused boundary test
            F.edges = edges;
            F.edge_bytes = edge_bytes;
            for (int a = 0; a < 3; ++a) F.est[a] = G.est[a];
        }
        // tier 1 (<= 24 ids near the plane, many CTAs/SM) -> tier 2 (<= 64 ids) on the listed segments -> generic kernel
        unsigned *hard = nullptr;  // [count1, count2, queue1, queue2 | list1[nsegs], list2[nsegs]]; counts and queue heads per tier
        SYK_CUDA(cudaMallocAsync((void **)&hard, sizeof(unsigned) * (size_t)(2 * F.nsegs + 8), s));
        SYK_CUDA(cudaMemsetAsync(hard, 0, 8 * sizeof(unsigned), s));
        unsigned *cnt1 = hard, *cnt2 = hard + 4, *list1 = hard + 8, *list2 = hard + 8 + F.nsegs;
        // rows of uint32 that are 16-byte aligned everywhere => LDG.128 quads
        F.vec4 = (elem_bytes == 4 && F.ist[2] == 1 && (F.ist[0] % 4) == 0 && (F.ist[1] % 4) == 0 && ((uintptr_t)arr % 16) == 0) ? 1 : 0;
        F.out_vec = (F.ost[2] == 1 && (F.ost[0] % 2) == 0 && (F.ost[1] % 2) == 0 && ((uintptr_t)out % 16) == 0) ? 1 : 0;
        const FastSmem L1 = fast_layout(F, GMAX_T1), L2 = fast_layout(F, GMAX_T2);
#ifndef SYK_MINB2
#define SYK_MINB2 2
#endif
        constexpr int MINB1 = SYK_MINB1, MINB2 = SYK_MINB2;
        int ctas1 = (int)((227 * 1024) / (L1.total + (int)sizeof(HashT<GMAX_T1>) + 1280));
        if (ctas1 > MINB1) ctas1 = MINB1;
        if (ctas1 < 1) ctas1 = 1;
        int ctas2 = (int)((227 * 1024) / (L2.total + (int)sizeof(HashT<GMAX_T2>) + 1280));
        if (ctas2 > MINB2) ctas2 = MINB2;
        if (ctas2 < 1) ctas2 = 1;
        long long g1 = (long long)sms * ctas1;
        if (g1 > F.nsegs) g1 = F.nsegs;
        const long long g2 = (long long)sms * ctas2;
        unsigned long long *o = (unsigned long long *)out;
        // tier 1 is specialised for the production stencil [13, 13, 7] in both memory orders (x fastest: internal
        // (u, v, w) = (7, 13, 13); C order: (13, 13, 7)); any other stencil runs the same kernel with run-time geometry
        void (*k1)(const void *, unsigned long long *, FastGeom, FastSmem, const unsigned *, const unsigned *, unsigned *, unsigned *,
                   const CUtensorMap);
        void (*k2)(const void *, unsigned long long *, FastGeom, FastSmem, const unsigned *, const unsigned *, unsigned *, unsigned *,
                   const CUtensorMap);
        // input planes by TMA: one box of WP x VP x 1 uint32 per plane (the haloed cross-section of a segment)
        CUtensorMap tmap_fast;
        memset(&tmap_fast, 0, sizeof(tmap_fast));
        F.tma = (F.vec4 && !getenv("SYK_CS_NO_TMA") && syk_make_tmap3(&tmap_fast, arr, 4, F.n, F.ist, F.WP, F.VP)) ? 1 : 0;
        const bool s7 = F.sten[0] == 7 && F.sten[1] == 13 && F.sten[2] == 13 && !getenv("SYK_CS_NOSPEC");
        const bool s13 = F.sten[0] == 13 && F.sten[1] == 13 && F.sten[2] == 7 && !getenv("SYK_CS_NOSPEC");
        if (edges != nullptr) {  // explicit edge mask: run-time-geometry kernels with the EDGES variant of the boundary phase
            F.tma = 0;
            if (F.vec4) {
                k1 = k_cs_fast<true, GMAX_T1, NT_T1, MINB1, 0, 0, 0, true>;
                k2 = k_cs_fast<true, GMAX_T2, NT_T2, MINB2, 0, 0, 0, true>;
            } else {
                k1 = k_cs_fast<false, GMAX_T1, NT_T1, MINB1, 0, 0, 0, true>;
                k2 = k_cs_fast<false, GMAX_T2, NT_T2, MINB2, 0, 0, 0, true>;
            }
        } else if (F.tma) {
            k1 = s7 ? k_cs_fast<true, GMAX_T1, NT_T1, MINB1, 7, 13, 13, false, true>
                    : s13 ? k_cs_fast<true, GMAX_T1, NT_T1, MINB1, 13, 13, 7, false, true>
                          : k_cs_fast<true, GMAX_T1, NT_T1, MINB1, 0, 0, 0, false, true>;
            // tier 2 with the production stencil folded in as well: small supervoxels (pitch 16 x 16 x 8) run entirely in it
            k2 = s7 ? k_cs_fast<true, GMAX_T2, NT_T2, MINB2, 7, 13, 13, false, true>
                    : s13 ? k_cs_fast<true, GMAX_T2, NT_T2, MINB2, 13, 13, 7, false, true>
                          : k_cs_fast<true, GMAX_T2, NT_T2, MINB2, 0, 0, 0, false, true>;
        } else if (F.vec4) {
            k1 = s7 ? k_cs_fast<true, GMAX_T1, NT_T1, MINB1, 7, 13, 13>
                    : s13 ? k_cs_fast<true, GMAX_T1, NT_T1, MINB1, 13, 13, 7> : k_cs_fast<true, GMAX_T1, NT_T1, MINB1>;
            k2 = k_cs_fast<true, GMAX_T2, NT_T2, MINB2>;
        } else {
            k1 = k_cs_fast<false, GMAX_T1, NT_T1, MINB1>;
            k2 = k_cs_fast<false, GMAX_T2, NT_T2, MINB2>;
        }
        if ((rc = syk_ensure_dyn_smem((const void *)k2, L2.total))) return rc;
        // Experimental tier 1 (SYK_CS_MARCH=1): the u -> v -> w marching kernel of syk_cs_march.cuh (TMA planes, two barriers
        // per plane).  Bit-exact, but measured SLOWER than k_cs_fast on the production chunk (3.98 vs 2.74 ms,
        // profiles/r2_ncu_cs_march_summary.txt: its relabel / u-sum update run with few active lanes), so it is opt-in.
        CUtensorMap tmap;
        bool marched = false;
        if (F.vec4 && (s7 || s13) && edges == nullptr && getenv("SYK_CS_MARCH")) {
            constexpr int GM = 3, MB = 3;
            using C7 = csm::Cfg<7, 13, 13, GM>;
            using C13 = csm::Cfg<13, 13, 7, GM>;
            if (syk_make_tmap3(&tmap, arr, 4, F.n, F.ist, s7 ? C7::WP : C13::WP, s7 ? C7::VP : C13::VP)) {
                auto km7 = csm::k_cs_march<7, 13, 13, GM, MB>;
                auto km13 = csm::k_cs_march<13, 13, 7, GM, MB>;
                const int dyn = s7 ? C7::DYN_BYTES : C13::DYN_BYTES, nt = s7 ? C7::NT : C13::NT;
                if ((rc = syk_ensure_dyn_smem(s7 ? (const void *)km7 : (const void *)km13, dyn))) return rc;
                int per_sm = 1;
                SYK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s7 ? (const void *)km7 : (const void *)km13, nt, dyn));
                if (per_sm < 1) per_sm = 1;
                long long gm = (long long)sms * per_sm;
                if (gm > F.nsegs) gm = F.nsegs;
                ctas1 = per_sm;
                if (s7) km7<<<(unsigned)gm, nt, dyn, s>>>(tmap, o, F, list1, cnt1);
                else km13<<<(unsigned)gm, nt, dyn, s>>>(tmap, o, F, list1, cnt1);
                marched = true;
            }
        }
        if (!marched) {
            if ((rc = syk_ensure_dyn_smem((const void *)k1, L1.total))) return rc;
            k1<<<(unsigned)g1, NT_T1, L1.total, s>>>(arr, o, F, L1, nullptr, nullptr, list1, cnt1, tmap_fast);
        }
        k2<<<(unsigned)g2, NT_T2, L2.total, s>>>(arr, o, F, L2, list1, cnt1, list2, cnt2, tmap_fast);
        SYK_CUDA(cudaGetLastError());
        G.seg_list = list2;
        G.seg_count = cnt2;
        for (int a = 0; a < 3; ++a) G.segs[a] = F.segs[a];
        G.seg_tiles[0] = LU / OU;
        G.seg_tiles[1] = TV / OV;
        G.seg_tiles[2] = TW / OW;
        k_detect_cs<<<(unsigned)((long long)sms * bps), CS_THREADS, smem, s>>>(arr, edges, o, G);  // listed segments only
        SYK_CUDA(cudaGetLastError());
#ifdef SYK_NG_HIST
        if (getenv("SYK_CS_DEBUG")) {
            unsigned long long h[65];
            cudaStreamSynchronize(s);
            cudaMemcpyFromSymbol(h, csfast::g_ids_hist, sizeof(h));
            fprintf(stderr, "[syk] live ids per plane:");
            for (int i = 0; i < 65; ++i)
                if (h[i]) fprintf(stderr, " %d:%llu", i, h[i]);
            fprintf(stderr, "\n");
        }
#endif
        if (getenv("SYK_CS_DEBUG")) {
            unsigned nh8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            cudaMemcpyAsync(nh8, hard, sizeof(nh8), cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            const unsigned nh[2] = {nh8[0], nh8[4]};
            fprintf(stderr, "[syk] detect_cs fast path: %lld segments, %u redone by tier 2, %u by the generic kernel; smem %d / %d B, "
                            "CTAs/SM %d / %d, vec4=%d out_vec=%d tma=%d\n", F.nsegs, nh[0], nh[1], L1.total, L2.total, ctas1, ctas2, F.vec4,
                    F.out_vec, F.tma);
        }
        SYK_CUDA(cudaFreeAsync(hard, s));
        return SYK_OK;
    }
    k_detect_cs<<<(unsigned)grid, CS_THREADS, smem, s>>>(arr, edges, (unsigned long long *)out, G);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

// detect_cs on 64-bit ids: the reference takes the boundary mask on the FULL 64-bit values and only then narrows the ids to
// uint32 for the window histogram (find_object_properties.py:466-468) -- neighbours that are equal modulo 2^32 are
// still boundaries.  The mask is therefore computed first (64-bit compares) and handed to the kernels as explicit edges.
static int cs_launch(const void *edges, int edge_bytes, const int64_t *edge_strides, const void *arr, int elem_bytes,
                     const int64_t strides[3], const int64_t shape[3], const int32_t stencil[3], uint64_t *out,
                     const int64_t out_strides[3], cudaStream_t s, int first_seen = 0) {
    if (edges != nullptr || elem_bytes != 8 || first_seen || !shape || !strides || !arr)
        return cs_launch_impl(edges, edge_bytes, edge_strides, arr, elem_bytes, strides, shape, stencil, out, out_strides, s, first_seen);
    const long long total = shape[0] * shape[1] * shape[2];
    if (total <= 0) return cs_launch_impl(nullptr, 0, nullptr, arr, elem_bytes, strides, shape, stencil, out, out_strides, s, 0);
    int rc = syk_require_device();
    if (rc) return rc;
    unsigned char *mask = nullptr;
    SYK_CUDA(cudaMallocAsync((void **)&mask, (size_t)total, s));
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_seg_boundaries<<<(unsigned)blocks, 256, 0, s>>>(arr, 8, shape[0], shape[1], shape[2], strides[0], strides[1], strides[2], mask);
    const int64_t est[3] = {shape[1] * shape[2], shape[2], 1};
    rc = cs_launch_impl(mask, 1, est, arr, 8, strides, shape, stencil, out, out_strides, s, 0);
    cudaFreeAsync(mask, s);
    return rc;
}

}  // namespace

// extract_cs_syntype: stream compaction of the synaptic voxels (cs != 0 && syn != 0); lanes along the fastest cs axis
__global__ void __launch_bounds__(256, 6) k_syn_voxels(const void *__restrict__ cs, int elem_bytes, long long n0, long long n1, long long n2,  // internal u,v,w
                             long long c0, long long c1, long long c2, const unsigned char *__restrict__ syn, long long s0,
                             long long s1, long long s2, const unsigned char *__restrict__ asym, long long a0, long long a1,
                             long long a2, const unsigned char *__restrict__ sym, long long y0, long long y1, long long y2,
                             long long l0, long long l1, long long l2,  // linear-index coefficients of the internal axes
                             syk_synvox_t *__restrict__ out, unsigned long long max_out, unsigned long long *counter,
                             TableView syn_t, int p0, int p1, int p2,  // logical axis of the internal axes u, v, w
                             long long o0, long long o1, long long o2, unsigned chunk_seq) {
    const unsigned lane = threadIdx.x & 31;
    const long long nrows = n0 * n1;
    const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;  // warps in the grid
    constexpr int ILP = 4;  // segments in flight per warp: one 256-byte request per warp and round trip starves the memory system
    // a warp owns whole rows (u, v): the four base addresses are computed once per row, a segment costs a few additions --
    // with the index arithmetic per 32-voxel segment the pass was instruction-bound (~150 instructions per segment)
    for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += wstride) {
      const long long u = row / n1, v = row - u * n1;
      const long long cs_row = u * c0 + v * c1, syn_row = u * s0 + v * s1, as_row = u * a0 + v * a1, sy_row = u * y0 + v * y1;
      const long long lin_row = u * l0 + v * l1;
      for (long long w0 = 0; w0 < n2; w0 += ILP * 32) {
        unsigned long long keys[ILP];
        long long ws[ILP];
        bool hits[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            ws[k] = w0 + k * 32 + lane;
            keys[k] = 0ull;
            if (ws[k] < n2) {
                const long long ci = cs_row + ws[k] * c2;
                keys[k] = elem_bytes == 8 ? __ldg((const unsigned long long *)cs + ci) : (unsigned long long)__ldg((const unsigned *)cs + ci);
            }
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            hits[k] = false;
            if (keys[k] != 0ull) hits[k] = __ldg(syn + syn_row + ws[k] * s2) != 0;
        }
        // ONE position reservation per warp and round for all its segments: the tuple counter is a single address, and an
        // atomic per 32-voxel segment with a hit (hundreds of thousands per chunk) serialises there
        unsigned ms[ILP];
        unsigned nhit = 0u;
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            ms[k] = __ballot_sync(FULL, hits[k]);
            nhit += (unsigned)__popc(ms[k]);
        }
        if (nhit == 0u) continue;
        unsigned long long pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(counter, (unsigned long long)nhit);
        pos0 = __shfl_sync(FULL, pos0, 0);
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const unsigned long long key = keys[k];
            const long long w = ws[k];
            const bool hit = hits[k];
            const unsigned m = ms[k];
            if (k > 0) pos0 += (unsigned long long)__popc(ms[k - 1]);
            if (!m) continue;
            if (hit) {
                const unsigned long long pos = pos0 + __popc(m & ((1u << lane) - 1u));
                if (pos < max_out) {
                    syk_synvox_t t;
                    t.id = key;
                    t.lin = (unsigned long long)(lin_row + w * l2);
                    t.flags = (__ldg(asym + as_row + w * a2) == 1 ? 1ull : 0ull) | (__ldg(sym + sy_row + w * y2) == 1 ? 2ull : 0ull);
                    t._pad = 0;
                    out[pos] = t;
                }
                if (syn_t.slots) {  // props of the synaptic part of the contact (block_processing_C.pyx:117-137), one update per id and warp
                    const unsigned peers = __match_any_sync(m, key);
                    const long long lin = lin_row + w * l2;
                    long long c[3];  // logical coordinates
                    c[p0] = u, c[p1] = v, c[p2] = w;
                    // voxels per call < 2^31 (checked on the host): 32-bit warp reductions
                    const int mn0 = __reduce_min_sync(peers, (int)c[0]), mn1 = __reduce_min_sync(peers, (int)c[1]), mn2 = __reduce_min_sync(peers, (int)c[2]);
                    const int mx0 = __reduce_max_sync(peers, (int)c[0]), mx1 = __reduce_max_sync(peers, (int)c[1]), mx2 = __reduce_max_sync(peers, (int)c[2]);
                    const unsigned first = __reduce_min_sync(peers, (unsigned)lin);
                    if ((unsigned)(__ffs(peers) - 1) == lane) {
                        const unsigned long long rep_key = ((unsigned long long)chunk_seq << 40) | (SYK_REP_MASK - (unsigned long long)first);
                        syk_table_update(syn_t, key, (unsigned long long)__popc(peers), rep_key, (int)(mn0 + o0), (int)(mn1 + o1), (int)(mn2 + o2),
                                         (int)(mx0 + 1 + o0), (int)(mx1 + 1 + o1), (int)(mx2 + 1 + o2));
                    }
                }
            }
        }
      }
    }
}

SYK_API int syk_extract_cs_syntype_props(syk_table_t *cs_t, syk_table_t *syn_t, const void *cs_dev, int elem_bytes, const int64_t shape[3],
                                         const int64_t cs_strides[3], const uint8_t *syn_dev, const int64_t syn_strides[3],
                                         const uint8_t *asym_dev, const int64_t asym_strides[3], const uint8_t *sym_dev,
                                         const int64_t sym_strides[3], const int64_t origin[3], uint32_t chunk_seq, syk_synvox_t *vox_dev,
                                         uint64_t max_vox, uint64_t *counter_dev, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(shape && cs_strides && syn_strides && asym_strides && sym_strides, "NULL geometry argument");
    SYK_CHECK_ARG(counter_dev != nullptr && (vox_dev != nullptr || max_vox == 0), "NULL output");
    SYK_CHECK_ARG(!syn_t || origin, "the synaptic props need the block origin");
    if (cs_t) {
        rc = syk_find_object_properties(cs_t, cs_dev, elem_bytes, shape, cs_strides, origin, chunk_seq, stream);
        if (rc) return rc;
    }
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0) return SYK_OK;
    SYK_CHECK_ARG(cs_dev && syn_dev && asym_dev && sym_dev, "NULL buffer");
    SYK_CHECK_ARG(!syn_t || total < (1ll << 31), "more than 2^31 voxels per call");
    int ax[3] = {0, 1, 2};  // internal order: largest |cs stride| first, lanes along the smallest
    auto key = [&](int a) { return cs_strides[a] < 0 ? -cs_strides[a] : cs_strides[a]; };
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (key(ax[j]) > key(ax[i])) {
                int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    const long long lin[3] = {shape[1] * shape[2], shape[2], 1};
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_syn_voxels<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        cs_dev, elem_bytes, shape[ax[0]], shape[ax[1]], shape[ax[2]], cs_strides[ax[0]], cs_strides[ax[1]], cs_strides[ax[2]], syn_dev,
        syn_strides[ax[0]], syn_strides[ax[1]], syn_strides[ax[2]], asym_dev, asym_strides[ax[0]], asym_strides[ax[1]],
        asym_strides[ax[2]], sym_dev, sym_strides[ax[0]], sym_strides[ax[1]], sym_strides[ax[2]], lin[ax[0]], lin[ax[1]], lin[ax[2]],
        vox_dev, max_vox, (unsigned long long *)counter_dev, view_of(syn_t), ax[0], ax[1], ax[2], origin ? origin[0] : 0,
        origin ? origin[1] : 0, origin ? origin[2] : 0, chunk_seq);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

SYK_API int syk_extract_cs_syntype(syk_table_t *cs_t, const void *cs_dev, int elem_bytes, const int64_t shape[3],
                                   const int64_t cs_strides[3], const uint8_t *syn_dev, const int64_t syn_strides[3],
                                   const uint8_t *asym_dev, const int64_t asym_strides[3], const uint8_t *sym_dev,
                                   const int64_t sym_strides[3], const int64_t origin[3], uint32_t chunk_seq, syk_synvox_t *vox_dev,
                                   uint64_t max_vox, uint64_t *counter_dev, void *stream) {
    return syk_extract_cs_syntype_props(cs_t, nullptr, cs_dev, elem_bytes, shape, cs_strides, syn_dev, syn_strides, asym_dev, asym_strides,
                                        sym_dev, sym_strides, origin, chunk_seq, vox_dev, max_vox, counter_dev, stream);
}

SYK_API int syk_detect_seg_boundaries(const void *arr_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                      uint8_t *out_dev, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(shape && strides, "NULL geometry argument");
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0) return SYK_OK;
    SYK_CHECK_ARG(arr_dev && out_dev, "NULL buffer");
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_seg_boundaries<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(arr_dev, elem_bytes, shape[0], shape[1], shape[2],
                                                                        strides[0], strides[1], strides[2], out_dev);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

SYK_API int syk_process_block_nonzero(const void *edges_dev, int edge_bytes, const int64_t edge_strides[3], const void *arr_dev,
                                      int elem_bytes, const int64_t arr_strides[3], const int64_t shape[3],
                                      const int32_t stencil[3], uint64_t *out_dev, const int64_t out_strides[3], void *stream) {
    SYK_CHECK_ARG(edges_dev != nullptr && edge_strides != nullptr, "edges is NULL");
    SYK_CHECK_ARG(edge_bytes == 1 || edge_bytes == 4, "edge_bytes must be 1 or 4");
    return cs_launch(edges_dev, edge_bytes, edge_strides, arr_dev, elem_bytes, arr_strides, shape, stencil, out_dev, out_strides,
                     (cudaStream_t)stream);
}

SYK_API int syk_detect_cs(const void *arr_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                          const int32_t stencil[3], uint64_t *out_dev, const int64_t out_strides[3], void *stream) {
    return cs_launch(nullptr, 0, nullptr, arr_dev, elem_bytes, strides, shape, stencil, out_dev, out_strides, (cudaStream_t)stream);
}

// detect_contact_partners (find_object_properties.py:371-421), the numba 64-bit twin of process_block_nonzero: same
// window histogram, but ties go to the id met first in the x, y, z window scan and the result is the unordered
// pair (centre << 32) | partner of uint32 labels (0 = no partner).  edges_dev == NULL: boundary mask computed on the fly.
SYK_API int syk_detect_contact_partners(const void *edges_dev, int edge_bytes, const int64_t edge_strides[3], const void *arr_dev,
                                        const int64_t arr_strides[3], const int64_t shape[3], const int32_t stencil[3],
                                        uint64_t *out_dev, const int64_t out_strides[3], void *stream) {
    if (edges_dev) SYK_CHECK_ARG(edge_bytes == 1 || edge_bytes == 4, "edge_bytes must be 1 or 4");
    return cs_launch(edges_dev, edge_bytes, edge_strides, arr_dev, 4, arr_strides, shape, stencil, out_dev, out_strides,
                     (cudaStream_t)stream, 1);
}

// (centre << 32) | partner of dense labels -> [min, max] of the original ids (ids_dev[label - 1]; NULL = identity)
__global__ void k_cs64_unpack(const unsigned long long *__restrict__ packed, unsigned long long n,
                              const unsigned long long *__restrict__ ids, unsigned long long *__restrict__ out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = packed[i];
        unsigned long long a = 0ull, b = 0ull;
        if (p != 0ull) {
            const unsigned lc = (unsigned)(p >> 32), lp = (unsigned)p;
            a = lc ? (ids ? ids[lc - 1u] : lc) : 0ull;
            b = lp ? (ids ? ids[lp - 1u] : lp) : 0ull;
            if (a > b) {
                const unsigned long long t = a;
                a = b;
                b = t;
            }
        }
        out[2 * i] = a;
        out[2 * i + 1] = b;
    }
}
SYK_API int syk_cs64_unpack(const uint64_t *packed_dev, uint64_t n, const uint64_t *ids_dev, uint64_t *out_dev, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    if (n == 0) return SYK_OK;
    SYK_CHECK_ARG(packed_dev && out_dev, "NULL buffer");
    unsigned long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_cs64_unpack<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned long long *)packed_dev, n,
                                                                     (const unsigned long long *)ids_dev, (unsigned long long *)out_dev);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}
