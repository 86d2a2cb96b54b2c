// syk_props.cu -- per-object property extraction and organelle x cell overlap mapping (sm_100a).
//
//   syk_find_object_properties      <- syconn/extraction/find_object_properties_C.pyx:24-49
//   syk_map_subcell_extract_props   <- syconn/extraction/find_object_properties_C.pyx:112-192 (and map_subcell_C :72-109)
//
// Design (HBM-bound integer scan, every voxel is read exactly once):
//   * a warp owns a tile of TU x TV rows x 32 lanes; lanes lie along the memory-contiguous axis (w) so every row is one
//     coalesced 256 B (uint64) request.  Rows are streamed global -> shared with cp.async (LDGSTS), R rows per batch,
//     double buffered: no registers are tied up by loads in flight and the next batch is always on its way;
//   * per batch a lane run-length-compresses its R-row column (labels are blobs: 1-3 runs per column); pass k handles
//     the k-th run of every lane: one __match_any_sync groups the lanes holding the same id, the group leader derives
//     count / bounding box / first voxel from the peer mask (popc, ffs, clz; warp REDUX only for ragged groups);
//   * the leader updates a WARP-PRIVATE open-addressing table in shared memory (plain LDS/STS, one CAS to claim);
//   * per tile the private table is flushed into the global HBM table: one 64-bit atomicCAS claim plus
//     fire-and-forget atomicAdd/atomicMax per (tile, id).  A full private table degrades to direct global updates.
// Algorithmic traffic: elem_bytes per voxel and channel (8 B/voxel for uint64 labels).
#include <cuda.h>
#include <stdlib.h>

#include "syk_common.cuh"

namespace {

constexpr int TW = 32;   // tile extent along the lane axis (one voxel per lane)
constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_SUB = 4;

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

// ---- warp-private tables -----------------------------------------------------------------------------------------
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

struct Group {             // one aggregated (id, tile-local box) contribution
    unsigned cnt, rep, mnu, mnv, mnw, mxu, mxv, mxw;
};

__device__ __forceinline__ void global_add(const TableView &g, const ScanGeom &G, const TileCtx &T, unsigned long long key,
                                           const Group &r) {
    long long mn[3], mx[3], o[3];
    mn[G.la[0]] = T.t0[0] + r.mnu;
    mn[G.la[1]] = T.t0[1] + r.mnv;
    mn[G.la[2]] = T.t0[2] + r.mnw;
    mx[G.la[0]] = T.t0[0] + r.mxu + 1;
    mx[G.la[1]] = T.t0[1] + r.mxv + 1;
    mx[G.la[2]] = T.t0[2] + r.mxw + 1;
    o[G.la[0]] = T.t0[0];
    o[G.la[1]] = T.t0[1];
    o[G.la[2]] = T.t0[2];
    const unsigned lz = r.rep % (unsigned)T.td[2];
    const unsigned q = r.rep / (unsigned)T.td[2];
    const unsigned ly = q % (unsigned)T.td[1];
    const unsigned lx = q / (unsigned)T.td[1];
    const unsigned long long lin =
        ((unsigned long long)(o[0] + lx) * (unsigned long long)G.S[1] + (unsigned long long)(o[1] + ly)) * (unsigned long long)G.S[2] +
        (unsigned long long)(o[2] + lz);
    const unsigned long long rep_key = ((unsigned long long)G.chunk_seq << 40) | (SYK_REP_MASK - lin);
    syk_table_update(g, key, r.cnt, rep_key, (int)(mn[0] + G.origin[0]), (int)(mn[1] + G.origin[1]), (int)(mn[2] + G.origin[2]),
                     (int)(mx[0] + G.origin[0]), (int)(mx[1] + G.origin[1]), (int)(mx[2] + G.origin[2]));
}

// leader-only.  Returns false when the private table is full (caller then updates the global table directly).
template <int WS>
__device__ __forceinline__ bool wtab_add(WarpTab<WS> &tab, unsigned long long key, const Group &r) {
    unsigned slot = ((unsigned)key * 0x9E3779B1u ^ (unsigned)(key >> 32) * 0x85EBCA6Bu) >> (32 - __builtin_ctz(WS));
    bool inserted = false, found = false;
    for (int probes = 0; probes < WS; ++probes) {
        const unsigned long long k = tab.keys[slot];
        if (k == key) { found = true; break; }
        if (k == 0ull) {
            const unsigned long long prev = atomicCAS(&tab.keys[slot], 0ull, key);
            if (prev == 0ull) { inserted = true; break; }
            if (prev == key) { found = true; break; }
        }
        slot = (slot + 1) & (WS - 1);
    }
    if (inserted) {
        tab.rec[slot][0] = make_uint4(r.cnt, r.rep, r.mnu, r.mnv);
        tab.rec[slot][1] = make_uint4(r.mnw, r.mxu, r.mxv, r.mxw);
        return true;
    }
    if (!found) return false;
    uint4 a = tab.rec[slot][0], b = tab.rec[slot][1];
    a.x += r.cnt;
    a.y = min(a.y, r.rep);
    a.z = min(a.z, r.mnu);
    a.w = min(a.w, r.mnv);
    b.x = min(b.x, r.mnw);
    b.y = max(b.y, r.mxu);
    b.z = max(b.z, r.mxv);
    b.w = max(b.w, r.mxw);
    tab.rec[slot][0] = a;
    tab.rec[slot][1] = b;
    return true;
}

template <int WS>
__device__ __forceinline__ void wtab_flush(WarpTab<WS> &tab, const TableView &g, const ScanGeom &G, const TileCtx &T, int lane) {
    __syncwarp();
    for (int i = lane; i < WS; i += 32) {
        const unsigned long long key = tab.keys[i];
        if (key == 0ull) continue;
        const uint4 a = tab.rec[i][0], b = tab.rec[i][1];
        tab.keys[i] = 0ull;
        Group r;
        r.cnt = a.x; r.rep = a.y; r.mnu = a.z; r.mnv = a.w; r.mnw = b.x; r.mxu = b.y; r.mxv = b.z; r.mxw = b.w;
        global_add(g, G, T, key, r);
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

// ---- async row staging ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? (int)sizeof(T) : 0;  // src-size 0 => zero fill
    if (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA (cp.async.bulk.tensor) staging: one instruction per channel and batch, completion on an mbarrier ----------
struct alignas(64) TmapSet {
    CUtensorMap m[1 + MAX_SUB];
};
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SYK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SYK_DONE;\n"
        "bra SYK_WAIT;\n"
        "SYK_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<unsigned long long>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tile_setup(const ScanGeom &G, long long tile, int TU, int TV, TileCtx &T) {
    // ntiles < 2^31 (checked on the host): 32-bit divisions
    const unsigned t32 = (unsigned)tile, n2 = (unsigned)G.tiles[2], n1 = (unsigned)G.tiles[1];
    const unsigned r = t32 / n2;
    const long long tw = t32 - r * n2;
    const long long tu = r / n1;
    const long long tv = r - (unsigned)tu * n1;
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

// Run-length compress the lane's column buf[0..R)[lane]: bit j of the result (`bnd`) is set when a run starts at row j,
// bit j of `nzs` when that run is foreground (label != 0).  Background runs never cost a pass.
template <typename T, int R>
__device__ __forceinline__ unsigned run_starts(const T *buf, int lane, unsigned &nzs) {
    unsigned bnd = 1u;
    T prev = buf[lane];
    nzs = prev != 0 ? 1u : 0u;
#pragma unroll
    for (int j = 1; j < R; ++j) {
        const T v = buf[j * 32 + lane];
        const unsigned ch = (v != prev) ? (1u << j) : 0u;
        bnd |= ch;
        nzs |= (v != 0) ? ch : 0u;
        prev = v;
    }
    return bnd;
}

// Accumulate one staged batch (R rows starting at tile-local row lv0, tile-local u = lu) of one label channel.
template <typename T, int R, int WS>
__device__ __forceinline__ void acc_batch(const T *buf, unsigned bnd, unsigned nzs, WarpTab<WS> &tab, const TableView &g,
                                          const ScanGeom &G, const TileCtx &Tc, unsigned lu, unsigned lv0, int lane) {
    const unsigned rowrep = lu * Tc.cu + lv0 * Tc.cv + (unsigned)lane * Tc.cw;
    while (__any_sync(FULL, nzs != 0u)) {  // pass k: the k-th FOREGROUND run of every lane
        int s = R, e = R;
        unsigned long long key = 0ull;
        if (nzs) {
            s = __ffs(nzs) - 1;
            const unsigned above = bnd & (0xFFFFFFFEu << s);  // run boundaries after row s
            e = above ? (__ffs(above) - 1) : R;
            key = (unsigned long long)buf[s * 32 + lane];
            nzs &= nzs - 1u;
        }
        const unsigned peers = __match_any_sync(FULL, key);
        const bool partial = (key != 0ull) && !(s == 0 && e == R);
        const unsigned partial_mask = __ballot_sync(FULL, partial);
        const int first = __ffs(peers) - 1;
        const int last = 31 - __clz(peers);
        Group r;
        unsigned smin = (unsigned)s;
        if ((peers & partial_mask) == 0u) {  // every lane of the group holds one run spanning the whole batch
            r.cnt = (unsigned)R * (unsigned)__popc(peers);
            r.mnv = lv0;
            r.mxv = lv0 + R - 1;
        } else {  // ragged group: reduce over the peers (the group's lanes all take this branch)
            static_assert(R <= 16, "run start / end are OR-reduced as two 16-bit one-hot fields");
            r.cnt = __reduce_add_sync(peers, (unsigned)(e - s));
            const unsigned se = __reduce_or_sync(peers, (1u << s) | (0x8000u << e));  // bit s | bit 16 + (e - 1)
            smin = (unsigned)__ffs(se & 0xFFFFu) - 1u;
            r.mnv = lv0 + smin;
            r.mxv = lv0 + (unsigned)(31 - __clz(se)) - 16u;
        }
        // first voxel of the group in logical scan order (all lanes converged again): rep_local is a mixed-radix number,
        // so the minimum over the group is lexicographic -- lane first when the lane axis is the slower logical axis,
        // row first otherwise
        const unsigned at_min = __ballot_sync(FULL, (unsigned)s == smin);
        const int src = (Tc.cv > Tc.cw) ? __ffs(peers & at_min) - 1 : first;
        r.rep = __shfl_sync(FULL, rowrep + (unsigned)s * Tc.cv, src);
        if (lane == first && key != 0ull) {
            r.mnu = r.mxu = lu;
            r.mnw = (unsigned)first;
            r.mxw = (unsigned)last;
            if (!wtab_add<WS>(tab, key, r)) global_add(g, G, Tc, key, r);
        }
        __syncwarp();
    }
}

// Overlap pairs of one organelle channel (sbuf) against the cell channel (cbuf) for one staged batch.
template <typename T, int R, int PS>
__device__ __forceinline__ void acc_pairs(const T *sbuf, const T *cbuf, WarpPairTab<PS> &tab, const PairView &g, int lane) {
    unsigned bnd = 1u, act;  // run starts of the (organelle, cell) pair column / those with both ids non-zero
    {
        T ps = sbuf[lane], pc = cbuf[lane];
        act = (ps != 0 && pc != 0) ? 1u : 0u;
#pragma unroll
        for (int j = 1; j < R; ++j) {
            const T s = sbuf[j * 32 + lane], c = cbuf[j * 32 + lane];
            const unsigned ch = (s != ps || c != pc) ? (1u << j) : 0u;
            bnd |= ch;
            act |= (s != 0 && c != 0) ? ch : 0u;
            ps = s;
            pc = c;
        }
    }
    while (__any_sync(FULL, act != 0u)) {
        int s = R, e = R;
        unsigned long long ks = 0ull, kc = 0ull;
        if (act) {
            s = __ffs(act) - 1;
            const unsigned above = bnd & (0xFFFFFFFEu << s);
            e = above ? (__ffs(above) - 1) : R;
            ks = (unsigned long long)sbuf[s * 32 + lane];
            kc = (unsigned long long)cbuf[s * 32 + lane];
            act &= act - 1u;
        }
        const bool active = ks != 0ull;
        const unsigned peers = __match_any_sync(FULL, ks) & __match_any_sync(FULL, kc);
        const bool partial = active && !(s == 0 && e == R);
        const unsigned partial_mask = __ballot_sync(FULL, partial);
        unsigned cnt;
        const int first = __ffs(peers) - 1;
        if ((peers & partial_mask) == 0u) cnt = (unsigned)R * (unsigned)__popc(peers);
        else cnt = __reduce_add_sync(peers, (unsigned)(e - s));
        if (lane == first && active) {
            unsigned slot = ((unsigned)ks * 0x9E3779B1u ^ (unsigned)(ks >> 32) * 0x85EBCA6Bu ^ (unsigned)kc * 0xC2B2AE35u) >>
                            (32 - __builtin_ctz(PS));
            bool done = false;
            for (int probes = 0; probes < PS; ++probes) {
                const unsigned long long cs = tab.sub[slot];
                if (cs == ks && tab.cell[slot] == kc) { tab.cnt[slot] += cnt; done = true; break; }
                if (cs == 0ull) {
                    const unsigned long long prev = atomicCAS(&tab.sub[slot], 0ull, ks);
                    if (prev == 0ull) {
                        tab.cell[slot] = kc;
                        tab.cnt[slot] = cnt;
                        done = true;
                        break;
                    }
                }
                slot = (slot + 1) & (PS - 1);
            }
            if (!done) syk_pairs_update(g, ks, kc, (unsigned long long)cnt);  // private table full
        }
        __syncwarp();
    }
}

// ---- kernels -----------------------------------------------------------------------------------------------------
struct MapArgs {
    const void *sub[MAX_SUB];
    TableView sub_t[MAX_SUB];
    PairView pair_t[MAX_SUB];
    int n_sub;
    int do_cell_props;
    int do_sub_props;
    int org_mode;  // organelle-first scan: the staged channel is ONE organelle volume (its props go to cell_t), sub[0] is the
                   // cell volume, read on demand only where the organelle is non-zero (organelles are sparse)
    // organelle-first scan of SEVERAL channels in one launch: tiles [c * ntiles, (c + 1) * ntiles) belong to channel c, whose
    // volume is org[c] (tensor map m[c]), props table org_t[c] and pair table pair_t[c]; one dynamic tile queue for all
    int n_org;
    const void *org[MAX_SUB];
    TableView org_t[MAX_SUB];
    unsigned long long *tile_ctr;  // device counter (zeroed per launch) for dynamic tile hand-out after the first wave; tiles differ
                                   // a lot in cost (empty / organelle content), a static round robin leaves a long tail
};

// One kernel for both entry points: NCH = 1 + n_sub staged channels (n_sub == 0 => find_object_properties).
//   R   rows per batch;  TU x TV rows per tile (TV multiple of R)
//   MODE 0: labels only (find_object_properties); 1: organelle-first scan of one channel; 2: fused, run-time channel count.
//   With MODE 0/1 the per-warp shared-memory layout is a compile-time constant (measured: -19% for MODE 0; the organelle scan
//   is launched with MODE 2, the specialised instantiation was 7% slower).
#ifndef SYK_ORG_MINB
#define SYK_ORG_MINB 5
#endif
template <typename T, int R, int TU, int TV, int WARPS, int WS, int WSS, int PS, int NBUF, bool TMA, int MODE>
__global__ void __launch_bounds__(WARPS * 32, (WARPS == 8 && R == 16) ? 4 : (MODE == 2 ? SYK_ORG_MINB : 3)) k_scan(const T *__restrict__ cell0, ScanGeom G, TableView cell_t, MapArgs A,
                                                     const __grid_constant__ TmapSet tm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar[WARPS][2];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int n_sub = MODE == 0 ? 0 : (MODE == 1 ? 1 : A.n_sub);
    const bool org_mode = MODE == 1 ? true : (MODE == 0 ? false : A.org_mode != 0);
    const int nch = 1 + n_sub;
    const int nstage = org_mode ? 1 : nch;  // channels staged by TMA / cp.async (org mode fills channel 1 on demand)
    // per-warp layout: stage[NBUF][nch][R*32] T (org mode: stage[NBUF][R*32] of the staged channel + ONE on-demand plane)
    //                  | WarpTab<WS> cell | n_stab x WarpTab<WSS> | n_sub x WarpPairTab<PS>
    const int buf_elems = (org_mode ? 1 : nch) * R * 32;  // elements between two stage buffers
    const size_t stage_bytes = (size_t)(org_mode ? NBUF + 1 : NBUF * nch) * R * 32 * sizeof(T);
    const int n_stab = org_mode ? 0 : n_sub;  // org mode keeps no props table for the on-demand (cell) channel
    const size_t per_warp = (stage_bytes + sizeof(WarpTab<WS>) + (size_t)n_stab * sizeof(WarpTab<WSS>) + (size_t)n_sub * sizeof(WarpPairTab<PS>) +
                             127) & ~(size_t)127;
    unsigned char *mine = smem_raw + per_warp * wib;
    T *stage = reinterpret_cast<T *>(mine);
    WarpTab<WS> *ctab = reinterpret_cast<WarpTab<WS> *>(mine + stage_bytes);
    WarpTab<WSS> *stab = reinterpret_cast<WarpTab<WSS> *>(ctab + 1);
    WarpPairTab<PS> *ptab = reinterpret_cast<WarpPairTab<PS> *>(stab + n_stab);
    for (int i = lane; i < WS; i += 32) ctab->keys[i] = 0ull;
    for (int c = 0; c < n_sub; ++c) {
        if (c < n_stab)
            for (int i = lane; i < WSS; i += 32) stab[c].keys[i] = 0ull;
        for (int i = lane; i < PS; i += 32) {
            ptab[c].sub[i] = 0ull;
            ptab[c].cell[i] = 0ull;
            ptab[c].cnt[i] = 0u;
        }
    }
    unsigned bar_addr[2] = {(unsigned)__cvta_generic_to_shared(&mbar[wib][0]), (unsigned)__cvta_generic_to_shared(&mbar[wib][1])};
    unsigned bar_phase[2] = {0u, 0u};
    if (TMA && lane == 0) {
        mbar_init(bar_addr[0], 1);
        mbar_init(bar_addr[1], 1);
    }
    __syncwarp();

    constexpr int NB = TU * (TV / R);  // batches per tile
    const long long nwarps = (long long)gridDim.x * WARPS;
    const long long my_first = (long long)blockIdx.x * WARPS + wib;
    const long long total_tiles = org_mode ? (long long)A.n_org * G.ntiles : G.ntiles;
    if (my_first >= total_tiles) return;

    // issue the async loads of batch b of the tile at tile coordinates (tu, tv, tw) into stage buffer `sb`
    struct TileId { long long tu, tv, tw, local; int ch; };
    auto tile_id = [&](long long tile) {
        TileId t;
        t.ch = org_mode ? (int)(tile / G.ntiles) : 0;
        t.local = tile - (long long)t.ch * G.ntiles;
        const unsigned t32 = (unsigned)t.local, n2 = (unsigned)G.tiles[2], n1 = (unsigned)G.tiles[1];
        const unsigned rr = t32 / n2;
        t.tw = t32 - rr * n2;
        t.tu = rr / n1;
        t.tv = rr - (unsigned)t.tu * n1;
        return t;
    };
    auto issue = [&](const TileId &t, int b, int sb) {
        if (TMA) {
            if (lane == 0) {
                const int c0 = (int)(t.tw * TW), c1 = (int)(t.tv * TV + (b % (TV / R)) * R), c2 = (int)(t.tu * TU + b / (TV / R));
                mbar_expect_tx(bar_addr[sb], (unsigned)(nstage * R * 32 * sizeof(T)));
                const unsigned dst0 = (unsigned)__cvta_generic_to_shared(stage + (size_t)sb * buf_elems);
                if (org_mode) tma_load_3d(dst0, &tm.m[t.ch], c0, c1, c2, bar_addr[sb]);
                else
                    for (int c = 0; c < nstage; ++c) tma_load_3d(dst0 + c * R * 32 * (unsigned)sizeof(T), &tm.m[c], c0, c1, c2, bar_addr[sb]);
            }
            return;
        }
        const long long w = t.tw * TW + lane;
        const long long u = t.tu * TU + b / (TV / R);
        const long long v0 = t.tv * TV + (b % (TV / R)) * R;
        const bool ok = (w < G.n[2]) && (u < G.n[0]);
        T *dst = stage + (size_t)sb * buf_elems + lane;
        const T *cell = org_mode ? reinterpret_cast<const T *>(A.org[t.ch]) : cell0;
        const T *src = ok ? cell + w * G.st[2] + u * G.st[0] + v0 * G.st[1] : cell;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const bool okj = ok && (v0 + j < G.n[1]);
            cp_async_elem<T>(dst + j * 32, okj ? src + j * G.st[1] : cell, okj);
        }
        for (int c = 0; c < nstage - 1; ++c) {
            const T *sc = reinterpret_cast<const T *>(A.sub[c]);
            const T *ss = ok ? sc + w * G.sst[2] + u * G.sst[0] + v0 * G.sst[1] : sc;
            T *dd = dst + (size_t)(1 + c) * R * 32;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const bool okj = ok && (v0 + j < G.n[1]);
                cp_async_elem<T>(dd + j * 32, okj ? ss + j * G.sst[1] : sc, okj);
            }
        }
        cp_async_commit();
    };

    int sb = 0;
    TileId cur = tile_id(my_first), nxt = cur;
    issue(cur, 0, 0);
    long long tile_next = my_first;
    for (long long tile = my_first; tile < total_tiles; tile = tile_next) {
        TileCtx Tc;
        tile_setup(G, cur.local, TU, TV, Tc);
        if (A.tile_ctr != nullptr) {  // first wave: one tile per warp; then whoever is free takes the next one
            unsigned long long t = 0ull;
            if (lane == 0) t = atomicAdd(A.tile_ctr, 1ull);
            tile_next = nwarps + (long long)__shfl_sync(FULL, t, 0);
        } else {
            tile_next = tile + nwarps;
        }
        const bool has_next = tile_next < total_tiles;
        if (has_next) nxt = tile_id(tile_next);
        for (int b = 0; b < NB; ++b) {
            const bool more = (b + 1 < NB) || has_next;
            // NBUF == 2: prefetch the next batch (possibly of the next tile) before consuming this one
            if (NBUF == 2 && more) {
                if (b + 1 < NB) issue(cur, b + 1, sb ^ 1);
                else issue(nxt, 0, sb ^ 1);
            }
            if (TMA) {
                mbar_wait(bar_addr[sb], bar_phase[sb]);
                bar_phase[sb] ^= 1u;
            } else {
                if (NBUF == 2 && more) cp_async_wait<1>();
                else cp_async_wait<0>();
            }
            __syncwarp();
            const unsigned lu = (unsigned)(b / (TV / R));
            const unsigned lv0 = (unsigned)((b % (TV / R)) * R);
            const T *cb = stage + (size_t)sb * buf_elems;
            unsigned nzs = 0u;
            bool empty = false;
            if (org_mode) {  // organelles are sparse: most batches hold no voxel at all -- 16-byte loads, OR-reduced, no run logic
                const uint4 *q4 = reinterpret_cast<const uint4 *>(cb);
                uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int i = 0; i < (int)(R * 32 * sizeof(T) / 16 / 32); ++i) {
                    const uint4 x = q4[i * 32 + lane];
                    acc.x |= x.x; acc.y |= x.y; acc.z |= x.z; acc.w |= x.w;
                }
                empty = !__any_sync(FULL, (acc.x | acc.y | acc.z | acc.w) != 0u);
            }
            if (empty) {
            } else if (org_mode) {
                // staged channel 0 = organelle (its props are accumulated as the "cell" channel); the real cell volume is
                // fetched only where the organelle is non-zero -- asynchronously (cp.async, zero fill elsewhere), so that
                // the DRAM round trip overlaps the props work of the same batch instead of stalling the warp
                const unsigned bnd = run_starts<T, R>(cb, lane, nzs);
                const bool any = __any_sync(FULL, nzs != 0u);
                T *cbuf = stage + (size_t)NBUF * R * 32;  // shared by the stage buffers
                if (any) {
                    const long long w = cur.tw * TW + lane, u = cur.tu * TU + lu, v0 = cur.tv * TV + lv0;
                    const T *cbase = reinterpret_cast<const T *>(A.sub[0]);
                    const T *cp = cbase + w * G.sst[2] + u * G.sst[0] + v0 * G.sst[1];
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        const bool need = cb[j * 32 + lane] != 0;
                        cp_async_elem<T>(cbuf + j * 32 + lane, need ? cp + j * G.sst[1] : cbase, need);
                    }
                    cp_async_commit();
                }
                if (A.do_cell_props) acc_batch<T, R, WS>(cb, bnd, nzs, *ctab, A.org_t[cur.ch], G, Tc, lu, lv0, lane);
                if (any) {
                    cp_async_wait<0>();
                    __syncwarp();
                    acc_pairs<T, R, PS>(cb, cbuf, ptab[0], A.pair_t[cur.ch], lane);
                }
            } else if (A.do_cell_props) {
                const unsigned bnd = run_starts<T, R>(cb, lane, nzs);
                acc_batch<T, R, WS>(cb, bnd, nzs, *ctab, cell_t, G, Tc, lu, lv0, lane);
            }
            if (!org_mode) {
                for (int c = 0; c < n_sub; ++c) {
                    const T *sbuf = cb + (size_t)(1 + c) * R * 32;
                    unsigned snz;
                    const unsigned bnd = run_starts<T, R>(sbuf, lane, snz);
                    if (!__any_sync(FULL, snz != 0u)) continue;  // organelles are sparse: most batches stop here
                    if (A.do_sub_props) acc_batch<T, R, WSS>(sbuf, bnd, snz, stab[c], A.sub_t[c], G, Tc, lu, lv0, lane);
                    acc_pairs<T, R, PS>(sbuf, cb, ptab[c], A.pair_t[c], lane);
                }
            }
            __syncwarp();
            if (NBUF == 1) {  // single buffer: refill it now; the SM's other warps cover the latency
                if (more) {
                    if (b + 1 < NB) issue(cur, b + 1, 0);
                    else issue(nxt, 0, 0);
                }
            } else {
                sb ^= 1;
            }
        }
        if (A.do_cell_props) wtab_flush<WS>(*ctab, org_mode ? A.org_t[cur.ch] : cell_t, G, Tc, lane);
        for (int c = 0; c < n_sub; ++c) {
            if (A.do_sub_props) wtab_flush<WSS>(stab[c], A.sub_t[c], G, Tc, lane);
            ptab_flush<PS>(ptab[c], A.pair_t[org_mode ? cur.ch : c], lane);  // per tile: a full private table would degrade every probe
        }
        cur = nxt;
    }
}

// choose the lane axis = smallest stride, row axis = next, slow axis = largest; degenerate axes (extent 1) go first
static void plan_axes(const int64_t shape[3], const int64_t strides[3], const int64_t origin[3], uint32_t chunk_seq, int TU,
                      int TV, ScanGeom &G) {
    int ax[3] = {0, 1, 2};
    auto key = [&](int a) { return strides[a] < 0 ? -strides[a] : strides[a]; };
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
    SYK_CHECK_ARG(nvox / 1024.0 < 2.0e9, "too many tiles per call");
    return SYK_OK;
}

// ---- TMA tensor maps (driver entry point resolved through the runtime: no link against libcuda) ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}

// rank-3 map over the internal axes (w innermost); box = 32 lanes x R rows x 1 plane.  false => use the LDGSTS path
static bool make_tmap(CUtensorMap *m, const void *base, int elem_bytes, const long long n[3], const long long st[3], int R) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || getenv("SYK_NO_TMA")) return false;
    if (st[2] != 1 || ((uintptr_t)base & 15)) return false;
    for (int a = 0; a < 2; ++a)
        if (st[a] <= 0 || ((st[a] * elem_bytes) & 15) || st[a] * elem_bytes >= (1ll << 40)) return false;
    for (int a = 0; a < 3; ++a)
        if (n[a] <= 0 || n[a] >= (1ll << 31)) return false;
    // the tensor is described with v as dim 1 and u as dim 2 whatever their stride order
    cuuint64_t dims[3] = {(cuuint64_t)n[2], (cuuint64_t)n[1], (cuuint64_t)n[0]};
    cuuint64_t strides[2] = {(cuuint64_t)(st[1] * elem_bytes), (cuuint64_t)(st[0] * elem_bytes)};
    cuuint32_t box[3] = {32u, (cuuint32_t)R, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    CUresult r = enc(m, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base),
                     dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <typename T, int R, int TU, int TV, int WARPS, int WS, int WSS, int PS, int NBUF, bool TMA, int MODE>
static int launch_cfg(const void *cell, const ScanGeom &G, const TableView &cell_t, const MapArgs &A, const TmapSet &tm, cudaStream_t s) {
    const int nch = 1 + A.n_sub;
    const int n_stab = (MODE == 1 || (MODE == 2 && A.org_mode)) ? 0 : A.n_sub;
    const bool org = MODE == 1 || (MODE == 2 && A.org_mode);
    const size_t per_warp = ((size_t)(org ? NBUF + 1 : NBUF * nch) * R * 32 * sizeof(T) + sizeof(WarpTab<WS>) + (size_t)n_stab * sizeof(WarpTab<WSS>) +
                             (size_t)A.n_sub * sizeof(WarpPairTab<PS>) + 127) & ~(size_t)127;
    const size_t smem = per_warp * WARPS;
    auto kern = k_scan<T, R, TU, TV, WARPS, WS, WSS, PS, NBUF, TMA, MODE>;
    { int rc_ = syk_ensure_dyn_smem((const void *)kern, (int)smem); if (rc_) return rc_; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int bps = 1;
    SYK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, WARPS * 32, smem));
    if (bps < 1) bps = 1;
    const long long all_tiles = (MODE != 0 && A.org_mode) ? G.ntiles * (A.n_org > 0 ? A.n_org : 1) : G.ntiles;
    long long want = (all_tiles + WARPS - 1) / WARPS;
    long long grid = (long long)sms * bps;
    if (grid > want) grid = want < 1 ? 1 : want;
    MapArgs A2 = A;
    unsigned long long *ctr = nullptr;
    if (!getenv("SYK_SCAN_STATIC")) {
        SYK_CUDA(cudaMallocAsync((void **)&ctr, sizeof(unsigned long long), s));
        SYK_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));
    }
    A2.tile_ctr = ctr;
    kern<<<(unsigned)grid, WARPS * 32, smem, s>>>((const T *)cell, G, cell_t, A2, tm);
    SYK_CUDA(cudaGetLastError());
    if (ctr) SYK_CUDA(cudaFreeAsync(ctr, s));
    return SYK_OK;
}

template <typename T, int MODE, int R, int TU, int TV, int WARPS, int WS, int WSS, int PS>
static int launch_scan(const void *cell, const ScanGeom &G, const TableView &cell_t, const MapArgs &A, cudaStream_t s) {
    TmapSet tm;
    memset(&tm, 0, sizeof(tm));
    bool tma = make_tmap(&tm.m[0], cell, (int)sizeof(T), G.n, G.st, R);
    for (int c = 0; c < A.n_sub && tma && !A.org_mode; ++c) tma = make_tmap(&tm.m[1 + c], A.sub[c], (int)sizeof(T), G.n, G.sst, R);
    for (int c = 1; c < A.n_org && tma && A.org_mode; ++c) tma = make_tmap(&tm.m[c], A.org[c], (int)sizeof(T), G.n, G.st, R);
    // TMA: one buffer per warp, many resident warps hide the latency; LDGSTS fallback: double buffered
    // (measured: one buffer + 32 resident warps/SM beats double buffering with 16-20 warps, and R=16 beats R=8)
#ifndef SYK_TMA_NBUF
#define SYK_TMA_NBUF 1
#endif
#ifndef SYK_ORG_NBUF
#define SYK_ORG_NBUF SYK_TMA_NBUF
#endif
    constexpr int NBUF_TMA = MODE == 2 ? SYK_ORG_NBUF : SYK_TMA_NBUF;
    if (tma) return launch_cfg<T, R, TU, TV, WARPS, WS, WSS, PS, NBUF_TMA, true, MODE>(cell, G, cell_t, A, tm, s);
    return launch_cfg<T, R, TU, TV, WARPS, WS, WSS, PS, 2, false, MODE>(cell, G, cell_t, A, tm, s);
}

}  // namespace

// props: R=16 rows per batch, tiles of 4 x 32 rows x 32 lanes; map: R=8, tiles of 8 x 16 rows (more channels staged)
#define PROPS_R 16
#ifndef PROPS_TU
#define PROPS_TU 4
#endif
#ifndef PROPS_TV
#define PROPS_TV 32
#endif
#define PROPS_WARPS 8
#define PROPS_CFG PROPS_R, PROPS_TU, PROPS_TV, PROPS_WARPS, 64, 32, 32
#define MAP_CFG 8, 8, 16, 4, 64, 32, 32
// organelle-first scan (measured: 4-warp CTAs with 16-row batches and 8-plane tiles, 0.47 vs 0.57 ms per 512^3 channel)
#ifndef ORG_R
#define ORG_R 16
#endif
#ifndef ORG_TU
#define ORG_TU 4  // with the dynamic tile queue smaller tiles balance better: 1.23 -> 1.15 ms for three 512^3 channels
#endif
#ifndef ORG_TV
#define ORG_TV 32
#endif
#ifndef ORG_WARPS
#define ORG_WARPS 4
#endif
// 32-slot private props table + no table for the on-demand channel: 10.1 KB per warp, 5 CTAs/SM (1.55 -> 1.49 ms for three channels)
#ifndef ORG_WS
#define ORG_WS 32
#endif
#ifndef ORG_PS
#define ORG_PS 32
#endif
#define ORG_CFG ORG_R, ORG_TU, ORG_TV, ORG_WARPS, ORG_WS, 32, ORG_PS

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
    plan_axes(shape, strides, origin, chunk_seq, PROPS_TU, PROPS_TV, G);
    MapArgs A;
    memset(&A, 0, sizeof(A));
    A.do_cell_props = 1;
    if (elem_bytes == 8) return launch_scan<unsigned long long, 0, PROPS_CFG>(labels_dev, G, view_of(t), A, (cudaStream_t)stream);
    return launch_scan<unsigned int, 0, PROPS_CFG>(labels_dev, G, view_of(t), A, (cudaStream_t)stream);
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
    for (int c = 0; c < n_sub; ++c) {
        SYK_CHECK_ARG(subcell_dev[c] != nullptr && pair_t[c] != nullptr, "subcell channel / pair table is NULL");
        if (sub_t) SYK_CHECK_ARG(sub_t[c] != nullptr, "sub table is NULL");
    }
    if (!getenv("SYK_MAP_FUSED")) {
        // Organelle-first decomposition (organelles are sparse): the cell props are one dense scan; every organelle channel
        // is streamed on its own (all of them by one launch), and the cell volume is touched again only where that organelle
        // is non-zero.
        if (cell_t) {
            rc = syk_find_object_properties(cell_t, cell_dev, elem_bytes, shape, cell_strides, origin, chunk_seq, stream);
            if (rc) return rc;
        }
        if (n_sub > 0) {  // all organelle channels in ONE launch: a shared dynamic tile queue, one ramp-up and one tail
            ScanGeom G;
            plan_axes(shape, sub_strides, origin, chunk_seq, ORG_TU, ORG_TV, G);
            for (int a = 0; a < 3; ++a) G.sst[a] = cell_strides[G.la[a]];
            SYK_CHECK_ARG(G.ntiles * n_sub < (1ll << 31), "too many tiles per call");
            MapArgs A;
            memset(&A, 0, sizeof(A));
            A.n_sub = 1;
            A.org_mode = 1;
            A.do_cell_props = sub_t != nullptr;  // props of the staged (organelle) channels
            A.sub[0] = cell_dev;
            A.n_org = n_sub;
            for (int c = 0; c < n_sub; ++c) {
                A.org[c] = subcell_dev[c];
                A.org_t[c] = view_of(sub_t ? sub_t[c] : nullptr);
                A.pair_t[c] = view_of(pair_t[c]);
            }
            rc = elem_bytes == 8 ? launch_scan<unsigned long long, 2, ORG_CFG>(subcell_dev[0], G, A.org_t[0], A, (cudaStream_t)stream)
                                 : launch_scan<unsigned int, 2, ORG_CFG>(subcell_dev[0], G, A.org_t[0], A, (cudaStream_t)stream);
            if (rc) return rc;
        }
        return SYK_OK;
    }
    ScanGeom G;
    plan_axes(shape, cell_strides, origin, chunk_seq, 8, 16, G);
    for (int a = 0; a < 3; ++a) G.sst[a] = n_sub ? sub_strides[G.la[a]] : 0;
    MapArgs A;
    memset(&A, 0, sizeof(A));
    A.n_sub = n_sub;
    A.do_cell_props = cell_t != nullptr;
    A.do_sub_props = sub_t != nullptr;
    for (int c = 0; c < n_sub; ++c) {
        A.sub[c] = subcell_dev[c];
        A.pair_t[c] = view_of(pair_t[c]);
        if (sub_t) A.sub_t[c] = view_of(sub_t[c]);
    }
    if (elem_bytes == 8) return launch_scan<unsigned long long, 2, MAP_CFG>(cell_dev, G, view_of(cell_t), A, (cudaStream_t)stream);
    return launch_scan<unsigned int, 2, MAP_CFG>(cell_dev, G, view_of(cell_t), A, (cudaStream_t)stream);
}
