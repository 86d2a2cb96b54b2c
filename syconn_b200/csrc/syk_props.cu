// syk_props.cu -- per-object property extraction and organelle x cell overlap mapping (sm_100a).
//
//   syk_find_object_properties      <- syconn/extraction/find_object_properties_C.pyx:24-49
//   syk_map_subcell_extract_props   <- syconn/extraction/find_object_properties_C.pyx:112-192 (and map_subcell_C :72-109)
//
// Design (HBM-bound integer scan, every voxel is read exactly once):
//   * lanes of a warp lie along the memory-contiguous axis (w): one coalesced load instruction per row
//     (256 B of uint64 labels); a lane keeps R consecutive rows (axis v) in registers, so label runs along v
//     are compressed per lane without any communication;
//   * one __match_any_sync per pass groups the lanes that hold the same id; the group leader derives count,
//     bounding box and first-voxel index from the peer mask (popc / ffs / clz) -- no per-voxel atomics;
//   * the leader updates a WARP-PRIVATE open-addressing table in shared memory (plain LDS/STS, no atomics
//     except the slot claim);
//   * per tile (8 x 32 x 32 voxels) the private table is flushed into the global HBM table with one 64-bit
//     atomicCAS claim + fire-and-forget atomicAdd/atomicMax updates per (tile, id).
// Algorithmic traffic: elem_bytes per voxel and channel (8 B/voxel for uint64 labels).
#include "syk_common.cuh"

namespace {

constexpr int TU = 8;    // tile extent along the slowest internal axis
constexpr int TV = 32;   // tile extent along the row axis
constexpr int TW = 32;   // tile extent along the lane axis (one voxel per lane)
constexpr unsigned FULL = 0xffffffffu;

struct ScanGeom {
    long long n[3];       // extents along the internal axes (u, v, w)
    long long st[3];      // element strides of the cell volume along (u, v, w)
    long long sst[3];     // element strides of the organelle volumes along (u, v, w)
    int la[3];            // logical axis (0=x,1=y,2=z) of internal axis u, v, w
    long long S[3];       // logical shape
    long long origin[3];  // logical origin (added to every coordinate)
    unsigned chunk_seq;
    long long tiles[3];   // number of tiles along u, v, w
    long long ntiles;
};

// ---- warp-private table of per-id partial records ------------------------------------------------------------
// rec[slot][0] = {count, rep_local, min_u, min_v}, rec[slot][1] = {min_w, max_u, max_v, max_w}; tile-local, inclusive
template <int WS>
struct WarpTab {
    unsigned long long keys[WS];
    uint4 rec[WS][2];
};
template <int PS>
struct WarpPairTab {
    unsigned long long sub[PS];
    unsigned long long cell[PS];
    unsigned int cnt[PS];
};

struct TileCtx {
    unsigned cu, cv, cw;   // rep_local = lu*cu + lv*cv + lw*cw  (logical scan order inside the tile)
    int td[3];             // logical tile dims, td[la[a]] = tile dim of internal axis a
    long long t0[3];       // tile origin along internal axes
};

template <int WS>
__device__ __forceinline__ void wtab_clear(WarpTab<WS> &tab, int lane) {
    for (int i = lane; i < WS; i += 32) tab.keys[i] = 0ull;
    __syncwarp();
}

// leader-only: add one aggregated group to the private table; returns true when a new slot was claimed
template <int WS>
__device__ __forceinline__ bool wtab_add(WarpTab<WS> &tab, unsigned long long key, unsigned cnt, unsigned rep, unsigned mnu,
                                         unsigned mnv, unsigned mnw, unsigned mxu, unsigned mxv, unsigned mxw) {
    unsigned slot = syk_hash_id32(key) & (WS - 1);
    bool inserted = false;
    for (;;) {
        unsigned long long k = tab.keys[slot];
        if (k == key) break;
        if (k == 0ull) {
            unsigned long long prev = atomicCAS(&tab.keys[slot], 0ull, key);
            if (prev == 0ull) {
                inserted = true;
                break;
            }
            if (prev == key) break;
        }
        slot = (slot + 1) & (WS - 1);
    }
    if (inserted) {
        tab.rec[slot][0] = make_uint4(cnt, rep, mnu, mnv);
        tab.rec[slot][1] = make_uint4(mnw, mxu, mxv, mxw);
    } else {
        uint4 a = tab.rec[slot][0], b = tab.rec[slot][1];
        a.x += cnt;
        a.y = min(a.y, rep);
        a.z = min(a.z, mnu);
        a.w = min(a.w, mnv);
        b.x = min(b.x, mnw);
        b.y = max(b.y, mxu);
        b.z = max(b.z, mxv);
        b.w = max(b.w, mxw);
        tab.rec[slot][0] = a;
        tab.rec[slot][1] = b;
    }
    return inserted;
}

// flush the private table into the global table (one claim + 8 fire-and-forget atomics per (tile, id))
template <int WS>
__device__ __forceinline__ void wtab_flush(WarpTab<WS> &tab, const TableView &g, const ScanGeom &G, const TileCtx &T, int lane) {
    __syncwarp();
    for (int i = lane; i < WS; i += 32) {
        const unsigned long long key = tab.keys[i];
        if (key == 0ull) continue;
        const uint4 a = tab.rec[i][0], b = tab.rec[i][1];
        tab.keys[i] = 0ull;
        // internal (u,v,w) tile-local -> logical global
        long long mn[3], mx[3];
        mn[G.la[0]] = T.t0[0] + a.z;
        mn[G.la[1]] = T.t0[1] + a.w;
        mn[G.la[2]] = T.t0[2] + b.x;
        mx[G.la[0]] = T.t0[0] + b.y + 1;
        mx[G.la[1]] = T.t0[1] + b.z + 1;
        mx[G.la[2]] = T.t0[2] + b.w + 1;
        long long o[3];
        o[G.la[0]] = T.t0[0];
        o[G.la[1]] = T.t0[1];
        o[G.la[2]] = T.t0[2];
        const unsigned rep = a.y;
        const unsigned lz = rep % (unsigned)T.td[2];
        const unsigned q = rep / (unsigned)T.td[2];
        const unsigned ly = q % (unsigned)T.td[1];
        const unsigned lx = q / (unsigned)T.td[1];
        const unsigned long long lin =
            ((unsigned long long)(o[0] + lx) * (unsigned long long)G.S[1] + (unsigned long long)(o[1] + ly)) *
                (unsigned long long)G.S[2] +
            (unsigned long long)(o[2] + lz);
        const unsigned long long rep_key = ((unsigned long long)G.chunk_seq << 40) | (SYK_REP_MASK - lin);
        syk_table_update(g, key, a.x, rep_key, (int)(mn[0] + G.origin[0]), (int)(mn[1] + G.origin[1]),
                         (int)(mn[2] + G.origin[2]), (int)(mx[0] + G.origin[0]), (int)(mx[1] + G.origin[1]),
                         (int)(mx[2] + G.origin[2]));
    }
    __syncwarp();
}

template <int PS>
__device__ __forceinline__ void ptab_flush(WarpPairTab<PS> &tab, const PairView &g, int lane) {
    __syncwarp();
    for (int i = lane; i < PS; i += 32) {
        const unsigned long long s = tab.sub[i];
        if (s == 0ull) continue;
        syk_pairs_update(g, s, tab.cell[i], (unsigned long long)tab.cnt[i]);
        tab.sub[i] = 0ull;
        tab.cell[i] = 0ull;
    }
    __syncwarp();
}

template <int R>
__device__ __forceinline__ unsigned long long pick(const unsigned long long (&v)[R], int s) {
    unsigned long long k = v[0];
#pragma unroll
    for (int j = 1; j < R; ++j) k = (s == j) ? v[j] : k;
    return k;
}

// Accumulate one batch of R rows (rows lv0 .. lv0+R-1 of the tile at tile-local u = lu) of one label channel.
template <int R, int WS>
__device__ __forceinline__ void acc_batch(const unsigned long long (&v)[R], WarpTab<WS> &tab, int &n_used,
                                          const TableView &g, const ScanGeom &G, const TileCtx &T, unsigned lu, unsigned lv0,
                                          int lane) {
    // per-lane run starts along v
    unsigned bnd = 1u;
#pragma unroll
    for (int j = 1; j < R; ++j) bnd |= (v[j] != v[j - 1]) ? (1u << j) : 0u;
    // drop background runs right away
#pragma unroll
    for (int j = 0; j < R; ++j) {
        // a run of zeros still needs its start bit to terminate the previous run; it is skipped below via key == 0
    }
    const unsigned rowrep = lu * T.cu + lv0 * T.cv + (unsigned)lane * T.cw;
    while (__any_sync(FULL, bnd != 0u)) {
        if (n_used > WS - 32) {  // keep room for up to 32 inserts per pass
            wtab_flush<WS>(tab, g, G, T, lane);
            n_used = 0;
        }
        int s = R, e = R;
        unsigned long long key = 0ull;
        if (bnd) {
            s = __ffs(bnd) - 1;
            const unsigned rest = bnd & (bnd - 1u);
            e = rest ? (__ffs(rest) - 1) : R;
            key = pick<R>(v, s);
            bnd = rest;
        }
        if (!__any_sync(FULL, key != 0ull)) continue;
        const unsigned peers = __match_any_sync(FULL, key);
        const bool partial = (key != 0ull) && !(s == 0 && e == R);
        const unsigned partial_mask = __ballot_sync(FULL, partial);
        unsigned cnt, vs, ve, rep;
        const int first = __ffs(peers) - 1;
        const int last = 31 - __clz(peers);
        if ((peers & partial_mask) == 0u) {
            cnt = (unsigned)R * (unsigned)__popc(peers);
            vs = 0u;
            ve = (unsigned)R;
            rep = rowrep - (unsigned)(lane - first) * T.cw;
        } else {
            // some lane of this group holds a partial run: reduce over the peer group
            cnt = __reduce_add_sync(peers, (unsigned)(e - s));
            vs = __reduce_min_sync(peers, (unsigned)s);
            ve = __reduce_max_sync(peers, (unsigned)e);
            rep = __reduce_min_sync(peers, rowrep + (unsigned)s * T.cv);
        }
        bool inserted = false;
        if (lane == first && key != 0ull) {
            inserted = wtab_add<WS>(tab, key, cnt, rep, lu, lv0 + vs, (unsigned)first, lu, lv0 + ve - 1u, (unsigned)last);
        }
        n_used += __popc(__ballot_sync(FULL, inserted));
        __syncwarp();
    }
}

// Overlap pairs of one organelle channel against the cell channel for one batch of rows.
template <int R, int PS>
__device__ __forceinline__ void acc_pairs(const unsigned long long (&sv)[R], const unsigned long long (&cv)[R],
                                          WarpPairTab<PS> &tab, int &n_used, const PairView &g, int lane) {
    unsigned bnd = 1u;
#pragma unroll
    for (int j = 1; j < R; ++j) bnd |= (sv[j] != sv[j - 1] || cv[j] != cv[j - 1]) ? (1u << j) : 0u;
    while (__any_sync(FULL, bnd != 0u)) {
        if (n_used > PS - 32) {
            ptab_flush<PS>(tab, g, lane);
            n_used = 0;
        }
        int s = R, e = R;
        unsigned long long ks = 0ull, kc = 0ull;
        if (bnd) {
            s = __ffs(bnd) - 1;
            const unsigned rest = bnd & (bnd - 1u);
            e = rest ? (__ffs(rest) - 1) : R;
            ks = pick<R>(sv, s);
            kc = pick<R>(cv, s);
            bnd = rest;
        }
        const bool active = (ks != 0ull) && (kc != 0ull);
        if (!active) ks = kc = 0ull;
        if (!__any_sync(FULL, active)) continue;
        const unsigned peers = __match_any_sync(FULL, ks) & __match_any_sync(FULL, kc);
        const bool partial = active && !(s == 0 && e == R);
        const unsigned partial_mask = __ballot_sync(FULL, partial);
        unsigned cnt;
        const int first = __ffs(peers) - 1;
        if ((peers & partial_mask) == 0u) cnt = (unsigned)R * (unsigned)__popc(peers);
        else cnt = __reduce_add_sync(peers, (unsigned)(e - s));
        bool inserted = false;
        if (lane == first && active) {
            unsigned slot = syk_hash_id32(ks * 0x9E3779B97F4A7C15ULL + kc) & (PS - 1);
            for (;;) {
                const unsigned long long cs = tab.sub[slot];
                if (cs == ks && tab.cell[slot] == kc) break;
                if (cs == 0ull) {
                    const unsigned long long prev = atomicCAS(&tab.sub[slot], 0ull, ks);
                    if (prev == 0ull) {
                        tab.cell[slot] = kc;
                        tab.cnt[slot] = 0u;
                        inserted = true;
                        break;
                    }
                }
                slot = (slot + 1) & (PS - 1);
            }
            tab.cnt[slot] += cnt;
        }
        n_used += __popc(__ballot_sync(FULL, inserted));
        __syncwarp();
    }
}

template <typename T>
__device__ __forceinline__ unsigned long long ld_stream(const T *p) {
    return (unsigned long long)__ldcs(p);  // streaming: every voxel is read once
}

__device__ __forceinline__ void tile_setup(const ScanGeom &G, long long tile, TileCtx &T) {
    const long long tw = tile % G.tiles[2];
    const long long r = tile / G.tiles[2];
    const long long tv = r % G.tiles[1];
    const long long tu = r / G.tiles[1];
    T.t0[0] = tu * TU;
    T.t0[1] = tv * TV;
    T.t0[2] = tw * TW;
    const int dims[3] = {TU, TV, TW};
    T.td[G.la[0]] = dims[0];
    T.td[G.la[1]] = dims[1];
    T.td[G.la[2]] = dims[2];
    unsigned c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int l = G.la[a];
        c[a] = (l == 2) ? 1u : (l == 1 ? (unsigned)T.td[2] : (unsigned)(T.td[1] * T.td[2]));
    }
    T.cu = c[0];
    T.cv = c[1];
    T.cw = c[2];
}

// ---- find_object_properties ------------------------------------------------------------------------------------
constexpr int PROPS_WARPS = 8;
constexpr int PROPS_WS = 128;

template <typename T, int R>
__global__ void __launch_bounds__(PROPS_WARPS * 32) k_props(const T *__restrict__ base, ScanGeom G, TableView g) {
    __shared__ WarpTab<PROPS_WS> tabs[PROPS_WARPS];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    WarpTab<PROPS_WS> &tab = tabs[wib];
    wtab_clear<PROPS_WS>(tab, lane);
    int n_used = 0;
    const long long nwarps = (long long)gridDim.x * PROPS_WARPS;
    for (long long tile = (long long)blockIdx.x * PROPS_WARPS + wib; tile < G.ntiles; tile += nwarps) {
        TileCtx Tc;
        tile_setup(G, tile, Tc);
        const long long w = Tc.t0[2] + lane;
        const bool wok = w < G.n[2];
        const T *colp = base + w * G.st[2];
        constexpr int NB = TU * (TV / R);  // batches per tile
        unsigned long long cur[R], nxt[R];
        auto load_batch = [&](int b, unsigned long long (&dst)[R]) {
            const int lu = b / (TV / R);
            const int lv0 = (b % (TV / R)) * R;
            const long long u = Tc.t0[0] + lu;
            const bool uok = wok && (u < G.n[0]);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const long long v = Tc.t0[1] + lv0 + j;
                dst[j] = (uok && v < G.n[1]) ? ld_stream(colp + u * G.st[0] + v * G.st[1]) : 0ull;
            }
        };
        load_batch(0, cur);
        for (int b = 0; b < NB; ++b) {
            if (b + 1 < NB) load_batch(b + 1, nxt);
            const unsigned lu = (unsigned)(b / (TV / R));
            const unsigned lv0 = (unsigned)((b % (TV / R)) * R);
            acc_batch<R, PROPS_WS>(cur, tab, n_used, g, G, Tc, lu, lv0, lane);
#pragma unroll
            for (int j = 0; j < R; ++j) cur[j] = nxt[j];
        }
        wtab_flush<PROPS_WS>(tab, g, G, Tc, lane);
        n_used = 0;
    }
}

// ---- map_subcell_extract_props -----------------------------------------------------------------------------------
constexpr int MAP_WARPS = 4;
constexpr int MAP_WS = 64;
constexpr int MAP_PS = 64;
constexpr int MAX_SUB = 4;

struct MapArgs {
    const void *sub[MAX_SUB];
    TableView sub_t[MAX_SUB];
    PairView pair_t[MAX_SUB];
    int n_sub;
    int do_cell_props;
    int do_sub_props;
};

template <int R>
struct MapSmemWarp {
    WarpTab<MAP_WS> cell;
};

template <typename T, int R>
__global__ void __launch_bounds__(MAP_WARPS * 32) k_map(const T *__restrict__ cell, ScanGeom G, TableView cell_t, MapArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout per warp: WarpTab cell | n_sub x WarpTab sub | n_sub x WarpPairTab
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const size_t per_warp = sizeof(WarpTab<MAP_WS>) * (1 + A.n_sub) + sizeof(WarpPairTab<MAP_PS>) * A.n_sub;
    unsigned char *mine = smem_raw + per_warp * wib;
    WarpTab<MAP_WS> *ctab = reinterpret_cast<WarpTab<MAP_WS> *>(mine);
    WarpTab<MAP_WS> *stab = ctab + 1;
    WarpPairTab<MAP_PS> *ptab = reinterpret_cast<WarpPairTab<MAP_PS> *>(stab + A.n_sub);
    wtab_clear<MAP_WS>(*ctab, lane);
    for (int c = 0; c < A.n_sub; ++c) {
        wtab_clear<MAP_WS>(stab[c], lane);
        for (int i = lane; i < MAP_PS; i += 32) {
            ptab[c].sub[i] = 0ull;
            ptab[c].cell[i] = 0ull;
            ptab[c].cnt[i] = 0u;
        }
    }
    __syncwarp();
    int used_c = 0;
    int used_s[MAX_SUB], used_p[MAX_SUB];
#pragma unroll
    for (int c = 0; c < MAX_SUB; ++c) used_s[c] = used_p[c] = 0;

    const long long nwarps = (long long)gridDim.x * MAP_WARPS;
    for (long long tile = (long long)blockIdx.x * MAP_WARPS + wib; tile < G.ntiles; tile += nwarps) {
        TileCtx Tc;
        tile_setup(G, tile, Tc);
        const long long w = Tc.t0[2] + lane;
        const bool wok = w < G.n[2];
        constexpr int NB = TU * (TV / R);
        for (int b = 0; b < NB; ++b) {
            const unsigned lu = (unsigned)(b / (TV / R));
            const unsigned lv0 = (unsigned)((b % (TV / R)) * R);
            const long long u = Tc.t0[0] + lu;
            const bool uok = wok && (u < G.n[0]);
            unsigned long long cv[R];
            unsigned long long sv[MAX_SUB][R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const long long v = Tc.t0[1] + lv0 + j;
                cv[j] = (uok && v < G.n[1]) ? ld_stream(cell + w * G.st[2] + u * G.st[0] + v * G.st[1]) : 0ull;
            }
#pragma unroll
            for (int c = 0; c < MAX_SUB; ++c) {
                if (c < A.n_sub) {
                    const T *sp = reinterpret_cast<const T *>(A.sub[c]) + w * G.sst[2] + u * G.sst[0];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const long long v = Tc.t0[1] + lv0 + j;
                        sv[c][j] = (uok && v < G.n[1]) ? ld_stream(sp + v * G.sst[1]) : 0ull;
                    }
                }
            }
            if (A.do_cell_props) acc_batch<R, MAP_WS>(cv, *ctab, used_c, cell_t, G, Tc, lu, lv0, lane);
#pragma unroll
            for (int c = 0; c < MAX_SUB; ++c) {
                if (c < A.n_sub) {
                    bool nz = false;
#pragma unroll
                    for (int j = 0; j < R; ++j) nz |= (sv[c][j] != 0ull);
                    if (!__any_sync(FULL, nz)) continue;
                    if (A.do_sub_props) acc_batch<R, MAP_WS>(sv[c], stab[c], used_s[c], A.sub_t[c], G, Tc, lu, lv0, lane);
                    acc_pairs<R, MAP_PS>(sv[c], cv, ptab[c], used_p[c], A.pair_t[c], lane);
                }
            }
        }
        if (A.do_cell_props) {
            wtab_flush<MAP_WS>(*ctab, cell_t, G, Tc, lane);
            used_c = 0;
        }
#pragma unroll
        for (int c = 0; c < MAX_SUB; ++c) {
            if (c < A.n_sub) {
                if (A.do_sub_props) {
                    wtab_flush<MAP_WS>(stab[c], A.sub_t[c], G, Tc, lane);
                    used_s[c] = 0;
                }
            }
        }
    }
    for (int c = 0; c < A.n_sub; ++c) ptab_flush<MAP_PS>(ptab[c], A.pair_t[c], lane);
}

// choose the lane axis = smallest stride, row axis = next, slow axis = largest
static void plan_axes(const int64_t shape[3], const int64_t strides[3], const int64_t origin[3], uint32_t chunk_seq,
                      ScanGeom &G) {
    int ax[3] = {0, 1, 2};
    auto key = [&](int a) { return strides[a] < 0 ? -strides[a] : strides[a]; };
    // sort axes by |stride| descending -> u, v, w ; degenerate axes (extent 1) go first
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j) {
            const bool swap = (shape[ax[j]] == 1 && shape[ax[i]] != 1) ? true
                              : (shape[ax[i]] == 1 && shape[ax[j]] != 1) ? false
                                                                         : key(ax[j]) > key(ax[i]);
            if (swap) {
                int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
        }
    for (int a = 0; a < 3; ++a) {
        G.la[a] = ax[a];
        G.n[a] = shape[ax[a]];
        G.st[a] = strides[ax[a]];
        G.sst[a] = strides[ax[a]];
    }
    for (int a = 0; a < 3; ++a) {
        G.S[a] = shape[a];
        G.origin[a] = origin ? origin[a] : 0;
    }
    G.chunk_seq = chunk_seq;
    G.tiles[0] = (G.n[0] + TU - 1) / TU;
    G.tiles[1] = (G.n[1] + TV - 1) / TV;
    G.tiles[2] = (G.n[2] + TW - 1) / TW;
    G.ntiles = G.tiles[0] * G.tiles[1] * G.tiles[2];
}

static int check_geom(const int64_t shape[3], const int64_t strides[3], const int64_t origin[3], int elem_bytes) {
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8 (uint32 / uint64 labels)");
    SYK_CHECK_ARG(shape && strides, "shape/strides are NULL");
    double nvox = 1.0;
    for (int a = 0; a < 3; ++a) {
        SYK_CHECK_ARG(shape[a] >= 0, "negative shape");
        nvox *= (double)shape[a];
        if (origin) SYK_CHECK_ARG(origin[a] > -(1ll << 29) && origin[a] + shape[a] < (1ll << 29), "coordinates exceed 2^29");
    }
    SYK_CHECK_ARG(nvox < (double)((1ull << 40) - 2), "more than 2^40 voxels per call");
    return SYK_OK;
}

static int grid_for(long long ntiles, int warps_per_block, int blocks_per_sm) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (ntiles + warps_per_block - 1) / warps_per_block;
    long long cap = (long long)sms * blocks_per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

SYK_API int syk_find_object_properties(syk_table_t *t, const void *labels_dev, int elem_bytes, const int64_t shape[3],
                                       const int64_t strides[3], const int64_t origin[3], uint32_t chunk_seq, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(t != nullptr, "table is NULL");
    rc = check_geom(shape, strides, origin, elem_bytes);
    if (rc) return rc;
    SYK_CHECK_ARG(chunk_seq < (1u << 24), "chunk_seq must be < 2^24");
    if (shape[0] == 0 || shape[1] == 0 || shape[2] == 0) return SYK_OK;
    SYK_CHECK_ARG(labels_dev != nullptr, "labels_dev is NULL");
    ScanGeom G;
    plan_axes(shape, strides, origin, chunk_seq, G);
    const int grid = grid_for(G.ntiles, PROPS_WARPS, 4);
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_bytes == 8)
        k_props<unsigned long long, 4><<<grid, PROPS_WARPS * 32, 0, s>>>((const unsigned long long *)labels_dev, G, view_of(t));
    else
        k_props<unsigned int, 4><<<grid, PROPS_WARPS * 32, 0, s>>>((const unsigned int *)labels_dev, G, view_of(t));
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

SYK_API int syk_map_subcell_extract_props(syk_table_t *cell_t, syk_table_t *const *sub_t, syk_pairs_t *const *pair_t,
                                          const void *cell_dev, const int64_t cell_strides[3], const void *const *subcell_dev,
                                          const int64_t sub_strides[3], int n_sub, int elem_bytes, const int64_t shape[3],
                                          const int64_t origin[3], uint32_t chunk_seq, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    rc = check_geom(shape, cell_strides, origin, elem_bytes);
    if (rc) return rc;
    SYK_CHECK_ARG(n_sub >= 0 && n_sub <= MAX_SUB, "n_sub must be in [0, 4]");
    SYK_CHECK_ARG(chunk_seq < (1u << 24), "chunk_seq must be < 2^24");
    SYK_CHECK_ARG(n_sub == 0 || (subcell_dev && sub_strides && pair_t), "subcell arguments are NULL");
    if (shape[0] == 0 || shape[1] == 0 || shape[2] == 0) return SYK_OK;
    SYK_CHECK_ARG(cell_dev != nullptr, "cell_dev is NULL");
    ScanGeom G;
    plan_axes(shape, cell_strides, origin, chunk_seq, G);
    for (int a = 0; a < 3; ++a) G.sst[a] = n_sub ? sub_strides[G.la[a]] : 0;
    MapArgs A;
    memset(&A, 0, sizeof(A));
    A.n_sub = n_sub;
    A.do_cell_props = cell_t != nullptr;
    A.do_sub_props = sub_t != nullptr;
    for (int c = 0; c < n_sub; ++c) {
        SYK_CHECK_ARG(subcell_dev[c] != nullptr && pair_t[c] != nullptr, "subcell channel / pair table is NULL");
        A.sub[c] = subcell_dev[c];
        A.pair_t[c] = view_of(pair_t[c]);
        if (sub_t) {
            SYK_CHECK_ARG(sub_t[c] != nullptr, "sub table is NULL");
            A.sub_t[c] = view_of(sub_t[c]);
        }
    }
    const size_t per_warp = sizeof(WarpTab<MAP_WS>) * (1 + n_sub) + sizeof(WarpPairTab<MAP_PS>) * n_sub;
    const size_t smem = per_warp * MAP_WARPS;
    cudaStream_t s = (cudaStream_t)stream;
    int bps = (int)((200 * 1024) / (smem + 1024));
    if (bps < 1) bps = 1;
    if (bps > 6) bps = 6;
    const int grid = grid_for(G.ntiles, MAP_WARPS, bps);
    if (elem_bytes == 8) {
        SYK_CUDA(cudaFuncSetAttribute(k_map<unsigned long long, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_map<unsigned long long, 4><<<grid, MAP_WARPS * 32, smem, s>>>((const unsigned long long *)cell_dev, G, view_of(cell_t), A);
    } else {
        SYK_CUDA(cudaFuncSetAttribute(k_map<unsigned int, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_map<unsigned int, 4><<<grid, MAP_WARPS * 32, smem, s>>>((const unsigned int *)cell_dev, G, view_of(cell_t), A);
    }
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}
