// syk_ccl.cu -- connected-component labelling of a thresholded volume and label-overlap pairs (row f4, first slice).
//
//   syk_label_components     <- scipy.ndimage.label(tmp_data) as called by _object_segmentation_thread,
//                               syconn/extraction/object_extraction_steps.py:350-352 (default structure: 6-connectivity)
//                               after the threshold of :302-303 (tmp_data > thresholds[...])
//   syk_label_overlap_pairs  <- the co-located label pairs of _make_stitch_list_thread, :600-606
//
// Labelling = union-find over the voxel grid in MEMORY order (lanes along the contiguous axis, coalesced):
//   1. every foreground voxel starts as the child of the first voxel of its run along w inside its 32-voxel segment
//      (one ballot per warp, no memory traffic between neighbours);
//   2. run starts are united with the foreground neighbours one step back along w (across the segment edge), v and u
//      (atomicMin union, roots = smallest memory index, path halving);
//   3. every voxel is pointed at its root, and the root learns the smallest LOGICAL linear index (x, y, z order) of its
//      component (warp-aggregated atomicMin);
//   4. scipy numbers components in the order of their first voxel in logical scan order: the roots are sorted by that index
//      (cub radix sort) and voxels get 1 + rank.  The result is bit-identical to scipy.ndimage.label.
#include <cub/device/device_radix_sort.cuh>

#include "syk_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned BG = 0xFFFFFFFFu;

// exact division of a 32-bit number by a launch constant: one 64-bit multiply-high instead of the ~20-instruction udiv
struct FastDiv {
    unsigned long long m;  // floor(2^64 / d) + 1, 0 for d == 1
    unsigned d;
    __host__ FastDiv() : m(0), d(1) {}
    __host__ explicit FastDiv(unsigned dd) : m(dd > 1 ? ~0ull / dd + 1ull : 0ull), d(dd) {}
    __device__ __forceinline__ unsigned div(unsigned x) const { return m ? (unsigned)__umul64hi((unsigned long long)x, m) : x; }
};

struct CclGeom {
    FastDiv dn2, dn1;  // by n[2], n[1]
    long long n[3];    // internal axes u, v, w (w = smallest input stride)
    long long ist[3];  // input strides (elements)
    long long ost[3];  // label strides (elements)
    long long lc[3];   // logical linear-index coefficient of internal axis a
    long long total;
    unsigned long long thr;
    int elem_bytes;
};

__device__ __forceinline__ unsigned find_root(unsigned *parent, unsigned i) {
    unsigned p = parent[i];
    while (p != i) {
        const unsigned g = parent[p];
        if (g != p) parent[i] = g;  // path halving (benign race: parents only ever decrease towards the root)
        i = p;
        p = g;
    }
    return i;
}

// read-only variant for the flatten pass: there the only stores to parent[] are the final `parent[i] = root` of each voxel's
// own thread -- a path-halving store of another thread could overwrite such a final value with a stale, non-root ancestor
__device__ __forceinline__ unsigned find_root_ro(const unsigned *parent, unsigned i) {
    unsigned p = ((const volatile unsigned *)parent)[i];
    while (p != i) {
        i = p;
        p = ((const volatile unsigned *)parent)[i];
    }
    return i;
}

__device__ __forceinline__ void unite(unsigned *parent, unsigned a, unsigned b) {
    for (;;) {
        a = find_root(parent, a);
        b = find_root(parent, b);
        if (a == b) return;
        if (a > b) {
            const unsigned t = a;
            a = b;
            b = t;
        }
        const unsigned old = atomicMin(&parent[b], a);  // b was a root: hang it below the smaller root
        if (old == b) return;
        b = old;  // somebody re-parented b meanwhile: continue from there
    }
}

// rows of nw voxels are processed in segments of 32 lanes; index i = (u * nv + v) * nw + w.  A warp owns whole rows (the
// index arithmetic is paid once per row) and keeps four segments in flight (a one-byte load per lane is only 32 B per
// request).  Roots are always foreground, so only segments with foreground need the initial value of minlin.
template <typename T>
__global__ void k_ccl_init(const T *__restrict__ vol, CclGeom G, unsigned *__restrict__ parent, unsigned *__restrict__ minlin) {
    // total < 2^32 (checked on the host): 32-bit index arithmetic, 64-bit divisions would dominate the pass
    const unsigned n2 = (unsigned)G.n[2], n1 = (unsigned)G.n[1], nrows = (unsigned)(G.n[0] * G.n[1]);
    const unsigned lane = threadIdx.x & 31;
    const unsigned wstride = (gridDim.x * blockDim.x) >> 5;
    const T thr = (T)G.thr;  // the host clamps the threshold to the element type's range
    for (unsigned row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += wstride) {
        const unsigned u = G.dn1.div(row), v = row - u * n1;
        const T *src = vol + ((long long)u * G.ist[0] + (long long)v * G.ist[1]);
        const unsigned i0 = row * n2;
        for (unsigned w0 = 0; w0 < n2; w0 += 128u) {
            bool f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned w = w0 + 32u * k + lane;
                f[k] = w < n2 && src[(long long)w * G.ist[2]] > thr;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned w = w0 + 32u * k + lane;
                const unsigned m = __ballot_sync(FULL, f[k]);
                if (w < n2) {
                    unsigned p = BG;
                    if (f[k]) {
                        const unsigned below = ~m & ((1u << lane) - 1u);            // background lanes before this one
                        const unsigned start = below ? 32u - __clz((int)below) : 0u;  // first lane of this run inside the segment
                        p = i0 + w - lane + start;
                    }
                    parent[i0 + w] = p;
                    if (m) minlin[i0 + w] = BG;  // whole lines: scattered 4-byte stores would each cost a read-modify-write in DRAM
                }
            }
            if (w0 + 128u < w0) break;
        }
        if (row + wstride < row) break;  // the loop counter itself would wrap
    }
}

// Foreground is sparse and a union is a chain of dependent loads: executed where they are found, one or two lanes of a warp
// would chase pointers while thirty wait.  Every warp therefore queues its (voxel, neighbour) pairs in shared memory and
// unites them 32 at a time, all lanes busy.
__global__ void __launch_bounds__(256) k_ccl_union(CclGeom G, unsigned *__restrict__ parent) {
    __shared__ unsigned q_a[8][128], q_b[8][128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned *qa = q_a[wid], *qb = q_b[wid];
    unsigned qn = 0u;  // warp-uniform, < 32 between iterations; one iteration adds at most 96
    // total < 2^32 (checked on the host): 32-bit index arithmetic, a 64-bit division per voxel would dominate the pass
    const unsigned n2 = (unsigned)G.n[2], n1 = (unsigned)G.n[1], plane = n1 * n2;
    const unsigned lt = (1u << lane) - 1u;
    const long long wstride = (long long)gridDim.x * 8 * 32;
    long long base = ((long long)blockIdx.x * 8 + wid) * 32;
    unsigned p = base + lane < G.total ? parent[base + lane] : BG;
    for (; base < G.total; base += wstride) {
        const long long nb = base + wstride;
        const unsigned pn = nb + lane < G.total ? parent[nb + lane] : BG;  // next segment's load is in flight during this one
        if (__any_sync(FULL, p != BG)) {
            const unsigned i = (unsigned)base + (unsigned)lane;
            bool t0 = false, t1 = false, t2 = false;
            if (p != BG) {
                const unsigned row = G.dn2.div(i), w = i - row * n2, v = row - G.dn1.div(row) * n1;
                const bool left = w > 0 && parent[i - 1] != BG;
                const bool run_start = !left || (w & 31) == 0;
                t0 = left && (w & 31) == 0;  // runs were cut at the 32-voxel segment edge by the init pass
                // a voxel inside a run only needs its v / u neighbour when the neighbour's own left neighbour is background
                // (otherwise the run start or an earlier voxel of the run already made that connection)
                t1 = v > 0 && parent[i - n2] != BG && (run_start || parent[i - n2 - 1] == BG);
                t2 = i >= plane && parent[i - plane] != BG && (run_start || parent[i - plane - 1] == BG);
            }
            unsigned m = __ballot_sync(FULL, t0);
            if (t0) { const unsigned k = qn + __popc(m & lt); qa[k] = i; qb[k] = i - 1; }
            qn += __popc(m);
            m = __ballot_sync(FULL, t1);
            if (t1) { const unsigned k = qn + __popc(m & lt); qa[k] = i; qb[k] = i - n2; }
            qn += __popc(m);
            m = __ballot_sync(FULL, t2);
            if (t2) { const unsigned k = qn + __popc(m & lt); qa[k] = i; qb[k] = i - plane; }
            qn += __popc(m);
            __syncwarp();
            while (qn >= 32u) {
                qn -= 32u;
                unite(parent, qa[qn + lane], qb[qn + lane]);
                __syncwarp();
            }
        }
        p = pn;
    }
    if ((unsigned)lane < qn) unite(parent, qa[lane], qb[lane]);
}

// parent[i] = root; minlin[root] = min logical linear index over the component; the roots are counted and, while the list
// has room, collected on the way.  Foreground voxels are queued per warp like the unions, so that all lanes chase a chain.
__device__ __forceinline__ void flatten32(const CclGeom &G, unsigned *parent, unsigned *minlin, bool valid, unsigned i, unsigned p,
                                          unsigned lane) {
    unsigned root = BG, lin = BG;
    if (valid) {
        root = p == i ? i : find_root_ro(parent, p);
        if (root != p) parent[i] = root;  // this entry is only ever written here, by the one lane that owns voxel i
        const unsigned n2u = (unsigned)G.n[2], n1u = (unsigned)G.n[1];
        const unsigned q = G.dn2.div(i), w = i - q * n2u, u = G.dn1.div(q), v = q - u * n1u;
        lin = u * (unsigned)G.lc[0] + v * (unsigned)G.lc[1] + w * (unsigned)G.lc[2];
    }
    const unsigned peers = __match_any_sync(FULL, root);
    const unsigned mn = __reduce_min_sync(peers, lin);
    if (valid && (unsigned)(__ffs(peers) - 1) == lane) atomicMin(&minlin[root], mn);
}

__global__ void __launch_bounds__(256) k_ccl_flatten(CclGeom G, unsigned *__restrict__ parent, unsigned *__restrict__ minlin,
                                                     unsigned long long *n_roots, unsigned *__restrict__ root_list,
                                                     unsigned long long list_cap) {
    __shared__ unsigned q_i[8][64], q_p[8][64];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned *qi = q_i[wid], *qp = q_p[wid];
    unsigned qn = 0u;  // warp-uniform, < 32 between iterations
    const unsigned lt = (1u << lane) - 1u;
    const long long wstride = (long long)gridDim.x * 8 * 32;
    long long base = ((long long)blockIdx.x * 8 + wid) * 32;
    unsigned p = base + lane < G.total ? parent[base + lane] : BG;
    for (; base < G.total; base += wstride) {
        const long long nb = base + wstride;
        const unsigned pn = nb + lane < G.total ? parent[nb + lane] : BG;
        const bool fg = p != BG;
        const unsigned m = __ballot_sync(FULL, fg);
        if (m) {
            const unsigned i = (unsigned)base + lane;
            const bool is_root = fg && p == i;  // stable: only non-root entries are ever rewritten during this pass
            const unsigned roots_here = __ballot_sync(FULL, is_root);
            if (roots_here) {
                unsigned long long at = 0ull;
                if (lane == 0) at = atomicAdd(n_roots, (unsigned long long)__popc(roots_here));
                at = __shfl_sync(FULL, at, 0) + __popc(roots_here & lt);
                if (is_root && at < list_cap) root_list[at] = i;
            }
            if (fg) {
                const unsigned k = qn + __popc(m & lt);
                qi[k] = i;
                qp[k] = p;
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32u) {
                qn -= 32u;
                flatten32(G, parent, minlin, true, qi[qn + lane], qp[qn + lane], lane);
                __syncwarp();
            }
        }
        p = pn;
    }
    if (qn) flatten32(G, parent, minlin, lane < qn, lane < qn ? qi[lane] : 0u, lane < qn ? qp[lane] : 0u, lane);
}

__global__ void k_ccl_keys(const unsigned *__restrict__ roots, const unsigned *__restrict__ minlin, unsigned long long n,
                           unsigned *__restrict__ keys) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = minlin[roots[i]];
}

__global__ void k_ccl_collect_roots(const unsigned *__restrict__ parent, unsigned *__restrict__ roots, long long total,
                                    unsigned long long *counter, unsigned long long max_roots) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (parent[i] == (unsigned)i) {
            const unsigned long long pos = atomicAdd(counter, 1ull);
            if (pos < max_roots) roots[pos] = (unsigned)i;
        }
    }
}

__global__ void k_ccl_rank(const unsigned *__restrict__ roots_sorted, unsigned long long n, unsigned *__restrict__ label_of_root) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label_of_root[roots_sorted[i]] = (unsigned)(i + 1ull);
}

// DENSE: the label array has the layout of the internal index (the usual case: labels allocated like the input)
template <bool DENSE>
__global__ void k_ccl_write(CclGeom G, const unsigned *__restrict__ parent, const unsigned *__restrict__ label_of_root,
                            unsigned *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.total; i += (long long)gridDim.x * blockDim.x) {
        const unsigned p = parent[i];
        const unsigned lab = p == BG ? 0u : label_of_root[p];
        if (DENSE) {
            __stcs(out + i, lab);
        } else {
            const unsigned iu = (unsigned)i, n2u = (unsigned)G.n[2], n1u = (unsigned)G.n[1];
            const unsigned q = G.dn2.div(iu), w = iu - q * n2u, u = G.dn1.div(q), v = q - u * n1u;
            out[(long long)u * G.ost[0] + (long long)v * G.ost[1] + (long long)w * G.ost[2]] = lab;
        }
    }
}

__global__ void k_label_pairs(const unsigned *__restrict__ a, long long a0, long long a1, long long a2, const unsigned *__restrict__ b,
                              long long b0, long long b1, long long b2, long long n0, long long n1, long long n2,
                              unsigned long long a_off, unsigned long long b_off, PairView t) {
    const long long total = n0 * n1 * n2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (total + stride - 1) / stride;
    for (long long r = 0; r < rounds; ++r) {
        const long long i = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        unsigned la = 0u, lb = 0u;
        if (i < total) {
            const long long w = i % n2, q = i / n2, v = q % n1, u = q / n1;
            la = a[u * a0 + v * a1 + w * a2];
            lb = b[u * b0 + v * b1 + w * b2];
        }
        const bool hit = la != 0u && lb != 0u;
        const unsigned long long key = hit ? (((unsigned long long)la << 32) | lb) : 0ull;
        const unsigned peers = __match_any_sync(FULL, key);
        if (hit && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31))
            syk_pairs_update(t, (unsigned long long)la + a_off, (unsigned long long)lb + b_off, (unsigned long long)__popc(peers));
    }
}

__global__ void k_label_map(unsigned *__restrict__ lab, long long n0, long long n1, long long n2, long long s0, long long s1, long long s2,
                            const unsigned *__restrict__ lut, unsigned long long lut_len) {
    const long long total = n0 * n1 * n2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i % n2, q = i / n2, v = q % n1, u = q / n1;
        unsigned *p = lab + u * s0 + v * s1 + w * s2;
        const unsigned x = *p;
        if (x != 0u && x < lut_len) {
            const unsigned y = lut[x];
            if (y != x) *p = y;
        }
    }
}

}  // namespace

SYK_API int syk_label_map(uint32_t *labels_dev, const int64_t shape[3], const int64_t strides[3], const uint32_t *lut_dev,
                          uint64_t lut_len, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(shape && strides, "NULL geometry argument");
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0 || lut_len == 0) return SYK_OK;
    SYK_CHECK_ARG(labels_dev && lut_dev, "NULL buffer");
    int ax[3] = {0, 1, 2};  // iterate in memory order
    auto key = [&](int a) { return strides[a] < 0 ? -strides[a] : strides[a]; };
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (key(ax[j]) > key(ax[i])) {
                const int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_label_map<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(labels_dev, shape[ax[0]], shape[ax[1]], shape[ax[2]], strides[ax[0]],
                                                                    strides[ax[1]], strides[ax[2]], lut_dev, lut_len);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

SYK_API int syk_label_components(const void *vol_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                 uint64_t threshold, uint32_t *labels_dev, const int64_t label_strides[3], uint64_t *n_labels_host,
                                 void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 1, 2, 4 or 8");
    SYK_CHECK_ARG(shape && strides && label_strides && n_labels_host, "NULL geometry argument");
    cudaStream_t s = (cudaStream_t)stream;
    *n_labels_host = 0;
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0) return SYK_OK;
    SYK_CHECK_ARG(vol_dev && labels_dev, "NULL buffer");
    SYK_CHECK_ARG(total < 0xFFFFFFF0ll, "more than 2^32 - 16 voxels per call");
    int ax[3] = {0, 1, 2};  // internal order: largest |input stride| first
    auto key = [&](int a) { return strides[a] < 0 ? -strides[a] : strides[a]; };
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (key(ax[j]) > key(ax[i])) {
                const int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    const long long lcoef[3] = {shape[1] * shape[2], shape[2], 1};
    CclGeom G;
    for (int a = 0; a < 3; ++a) {
        G.n[a] = shape[ax[a]];
        G.ist[a] = strides[ax[a]];
        G.ost[a] = label_strides[ax[a]];
        G.lc[a] = lcoef[ax[a]];
    }
    G.dn2 = FastDiv((unsigned)G.n[2]);
    G.dn1 = FastDiv((unsigned)G.n[1]);
    G.total = total;
    G.thr = threshold;
    G.elem_bytes = elem_bytes;
    // scratch from the stream-ordered pool, returned on every exit path
    struct Scratch {
        cudaStream_t s;
        void *p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        ~Scratch() {
            for (void *x : p)
                if (x) cudaFreeAsync(x, s);
        }
    } sc{s};
    syk_pool_keep_warm();  // a call per chunk must not pay allocator round trips for its two 4-byte-per-voxel arrays
    SYK_CUDA(cudaMallocAsync(&sc.p[0], sizeof(unsigned) * (size_t)total, s));
    SYK_CUDA(cudaMallocAsync(&sc.p[1], sizeof(unsigned) * (size_t)total, s));
    SYK_CUDA(cudaMallocAsync(&sc.p[2], sizeof(unsigned long long), s));
    unsigned *parent = (unsigned *)sc.p[0], *minlin = (unsigned *)sc.p[1];
    unsigned long long *counter = (unsigned long long *)sc.p[2];
    SYK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    // a threshold at or above the element type's maximum selects nothing; clamping keeps the typed comparison exact
    const unsigned long long tmax = elem_bytes == 8 ? ~0ull : (1ull << (8 * elem_bytes)) - 1ull;
    if (G.thr > tmax) G.thr = tmax;
    if (elem_bytes == 1) k_ccl_init<<<(unsigned)blocks, 256, 0, s>>>((const unsigned char *)vol_dev, G, parent, minlin);
    else if (elem_bytes == 2) k_ccl_init<<<(unsigned)blocks, 256, 0, s>>>((const unsigned short *)vol_dev, G, parent, minlin);
    else if (elem_bytes == 4) k_ccl_init<<<(unsigned)blocks, 256, 0, s>>>((const unsigned *)vol_dev, G, parent, minlin);
    else k_ccl_init<<<(unsigned)blocks, 256, 0, s>>>((const unsigned long long *)vol_dev, G, parent, minlin);
    k_ccl_union<<<(unsigned)blocks, 256, 0, s>>>(G, parent);
    // roots found by the flatten pass go to a list sized for the common case; a volume with more components (up to total / 2
    // in a checkerboard) is swept once more with a list of the counted size
    unsigned long long list_cap = (unsigned long long)total / 64 + 1024;
    SYK_CUDA(cudaMallocAsync(&sc.p[3], sizeof(unsigned) * 4 * (size_t)list_cap, s));
    k_ccl_flatten<<<(unsigned)blocks, 256, 0, s>>>(G, parent, minlin, counter, (unsigned *)sc.p[3], list_cap);
    SYK_CUDA(cudaGetLastError());
    unsigned long long n_roots = 0;
    SYK_CUDA(cudaMemcpyAsync(&n_roots, counter, sizeof(n_roots), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaStreamSynchronize(s));
    *n_labels_host = n_roots;
    if (n_roots > 0) {
        if (n_roots > list_cap) {
            SYK_CUDA(cudaFreeAsync(sc.p[3], s));
            sc.p[3] = nullptr;
            SYK_CUDA(cudaMallocAsync(&sc.p[3], sizeof(unsigned) * 4 * (size_t)n_roots, s));
            SYK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
            k_ccl_collect_roots<<<(unsigned)blocks, 256, 0, s>>>(parent, (unsigned *)sc.p[3], total, counter, n_roots);
        }
        // [roots | keys | sorted keys | sorted roots]
        unsigned *roots = (unsigned *)sc.p[3], *keys = roots + n_roots, *keys2 = keys + n_roots, *roots2 = keys2 + n_roots;
        k_ccl_keys<<<(unsigned)((n_roots + 255) / 256), 256, 0, s>>>(roots, minlin, n_roots, keys);
        size_t tmp_bytes = 0;
        SYK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, roots, roots2, (int)n_roots, 0, 32, s));
        SYK_CUDA(cudaMallocAsync(&sc.p[4], tmp_bytes ? tmp_bytes : 16, s));
        SYK_CUDA(cub::DeviceRadixSort::SortPairs(sc.p[4], tmp_bytes, keys, keys2, roots, roots2, (int)n_roots, 0, 32, s));
        // minlin is no longer needed: reuse it as label_of_root
        k_ccl_rank<<<(unsigned)((n_roots + 255) / 256), 256, 0, s>>>(roots2, n_roots, minlin);
    }
    const bool dense_out = G.ost[2] == 1 && G.ost[1] == G.n[2] && G.ost[0] == G.n[1] * G.n[2];
    if (dense_out) k_ccl_write<true><<<(unsigned)blocks, 256, 0, s>>>(G, parent, minlin, labels_dev);
    else k_ccl_write<false><<<(unsigned)blocks, 256, 0, s>>>(G, parent, minlin, labels_dev);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

SYK_API int syk_label_overlap_pairs(syk_pairs_t *pairs, const uint32_t *a_dev, const int64_t a_strides[3], const uint32_t *b_dev,
                                    const int64_t b_strides[3], const int64_t shape[3], uint64_t a_offset, uint64_t b_offset,
                                    void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(pairs && a_strides && b_strides && shape, "NULL argument");
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0) return SYK_OK;
    SYK_CHECK_ARG(a_dev && b_dev, "NULL buffer");
    int ax[3] = {0, 1, 2};
    auto key = [&](int a) { return a_strides[a] < 0 ? -a_strides[a] : a_strides[a]; };
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (key(ax[j]) > key(ax[i])) {
                const int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_label_pairs<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a_dev, a_strides[ax[0]], a_strides[ax[1]], a_strides[ax[2]], b_dev,
                                                                     b_strides[ax[0]], b_strides[ax[1]], b_strides[ax[2]], shape[ax[0]],
                                                                     shape[ax[1]], shape[ax[2]], a_offset, b_offset, view_of(pairs));
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}
