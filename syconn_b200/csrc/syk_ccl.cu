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

struct CclGeom {
    long long n[3];    // internal axes u, v, w (w = smallest input stride)
    long long ist[3];  // input strides (elements)
    long long ost[3];  // label strides (elements)
    long long lc[3];   // logical linear-index coefficient of internal axis a
    long long total;
    unsigned long long thr;
    int elem_bytes;
};

__device__ __forceinline__ bool fg_at(const void *vol, const CclGeom &G, long long u, long long v, long long w) {
    const long long o = u * G.ist[0] + v * G.ist[1] + w * G.ist[2];
    unsigned long long x;
    if (G.elem_bytes == 1) x = ((const unsigned char *)vol)[o];
    else if (G.elem_bytes == 2) x = ((const unsigned short *)vol)[o];
    else if (G.elem_bytes == 4) x = ((const unsigned *)vol)[o];
    else x = ((const unsigned long long *)vol)[o];
    return x > G.thr;
}

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

// rows of nw voxels are processed in segments of 32 lanes; index i = (u * nv + v) * nw + w
__global__ void k_ccl_init(const void *__restrict__ vol, CclGeom G, unsigned *__restrict__ parent) {
    const long long segs_per_row = (G.n[2] + 31) / 32;
    const long long nseg = G.n[0] * G.n[1] * segs_per_row;
    const int lane = threadIdx.x & 31;
    for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nseg; s += ((long long)gridDim.x * blockDim.x) >> 5) {
        const long long row = s / segs_per_row, w = (s - row * segs_per_row) * 32 + lane;
        const long long u = row / G.n[1], v = row - u * G.n[1];
        const bool f = w < G.n[2] && fg_at(vol, G, u, v, w);
        const unsigned m = __ballot_sync(FULL, f);
        if (w < G.n[2]) {
            unsigned p = BG;
            if (f) {
                const unsigned below = ~m & ((1u << lane) - 1u);            // background lanes before this one
                const int start = below ? 32 - __clz((int)below) : 0;       // first lane of this run inside the segment
                p = (unsigned)(row * G.n[2] + (w - lane + start));
            }
            parent[row * G.n[2] + w] = p;
        }
    }
}

__global__ void k_ccl_union(CclGeom G, unsigned *__restrict__ parent) {
    const long long plane = G.n[1] * G.n[2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.total; i += (long long)gridDim.x * blockDim.x) {
        const unsigned p = parent[i];
        if (p == BG) continue;
        // total < 2^32 (checked on the host): 32-bit index arithmetic, a 64-bit division per voxel would dominate the pass
        const unsigned iu = (unsigned)i, n2u = (unsigned)G.n[2], n1u = (unsigned)G.n[1];
        const unsigned rowu = iu / n2u;
        const long long w = iu - rowu * n2u;
        const long long v = rowu % n1u;
        const bool run_start = (w == 0) || parent[i - 1] == BG || (w & 31) == 0;
        if ((w & 31) == 0 && w > 0 && parent[i - 1] != BG) unite(parent, (unsigned)i, (unsigned)(i - 1));
        // a voxel inside a run only needs its v / u neighbour when the neighbour's own left neighbour is background
        // (otherwise the run start or an earlier voxel of the run already made that connection)
        if (v > 0 && parent[i - G.n[2]] != BG && (run_start || parent[i - G.n[2] - 1] == BG)) unite(parent, (unsigned)i, (unsigned)(i - G.n[2]));
        if (i >= plane && parent[i - plane] != BG && (run_start || parent[i - plane - 1] == BG)) unite(parent, (unsigned)i, (unsigned)(i - plane));
    }
}

// parent[i] = root; minlin[root] = min logical linear index over the component
__global__ void k_ccl_flatten(CclGeom G, unsigned *__restrict__ parent, unsigned *__restrict__ minlin, unsigned long long *n_roots) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (G.total + stride - 1) / stride;
    for (long long r = 0; r < rounds; ++r) {
        const long long i = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        unsigned root = BG, lin = BG;
        if (i < G.total && parent[i] != BG) {
            root = find_root_ro(parent, (unsigned)i);
            parent[i] = root;
            const unsigned iu = (unsigned)i, n2u = (unsigned)G.n[2], n1u = (unsigned)G.n[1];
            const unsigned q = iu / n2u, w = iu - q * n2u, u = q / n1u, v = q - u * n1u;
            lin = u * (unsigned)G.lc[0] + v * (unsigned)G.lc[1] + w * (unsigned)G.lc[2];
        }
        const unsigned peers = __match_any_sync(FULL, root);
        const unsigned mn = __reduce_min_sync(peers, lin);
        if (root != BG && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicMin(&minlin[root], mn);
        const unsigned roots_here = __ballot_sync(FULL, root != BG && root == (unsigned)i);  // the components are counted on the way
        if ((threadIdx.x & 31) == 0 && roots_here) atomicAdd(n_roots, (unsigned long long)__popc(roots_here));
    }
}

__global__ void k_ccl_collect_roots(const unsigned *__restrict__ parent, const unsigned *__restrict__ minlin, long long total,
                                    unsigned *__restrict__ keys, unsigned *__restrict__ roots, unsigned long long *counter,
                                    unsigned long long max_roots) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (parent[i] == (unsigned)i) {
            const unsigned long long pos = atomicAdd(counter, 1ull);
            if (pos < max_roots) {
                keys[pos] = minlin[i];
                roots[pos] = (unsigned)i;
            }
        }
    }
}

__global__ void k_ccl_rank(const unsigned *__restrict__ roots_sorted, unsigned long long n, unsigned *__restrict__ label_of_root) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label_of_root[roots_sorted[i]] = (unsigned)(i + 1ull);
}

__global__ void k_ccl_write(CclGeom G, const unsigned *__restrict__ parent, const unsigned *__restrict__ label_of_root,
                            unsigned *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.total; i += (long long)gridDim.x * blockDim.x) {
        const unsigned iu = (unsigned)i, n2u = (unsigned)G.n[2], n1u = (unsigned)G.n[1];
        const unsigned q = iu / n2u, w = iu - q * n2u, u = q / n1u, v = q - u * n1u;
        const unsigned p = parent[i];
        out[(long long)u * G.ost[0] + (long long)v * G.ost[1] + (long long)w * G.ost[2]] = p == BG ? 0u : label_of_root[p];
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

}  // namespace

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
    G.total = total;
    G.thr = threshold;
    G.elem_bytes = elem_bytes;
    unsigned *parent = nullptr, *minlin = nullptr, *keys = nullptr, *roots = nullptr, *keys2 = nullptr, *roots2 = nullptr;
    unsigned long long *counter = nullptr;
    void *tmp = nullptr;
    // components are at most total / 2 + 1 under 6-connectivity only in pathological checkerboards; size the root list for the
    // worst case lazily: first count, then allocate
    SYK_CUDA(cudaMallocAsync((void **)&parent, sizeof(unsigned) * (size_t)total, s));
    SYK_CUDA(cudaMallocAsync((void **)&minlin, sizeof(unsigned) * (size_t)total, s));
    SYK_CUDA(cudaMallocAsync((void **)&counter, sizeof(unsigned long long), s));
    SYK_CUDA(cudaMemsetAsync(minlin, 0xFF, sizeof(unsigned) * (size_t)total, s));
    SYK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_ccl_init<<<(unsigned)blocks, 256, 0, s>>>(vol_dev, G, parent);
    k_ccl_union<<<(unsigned)blocks, 256, 0, s>>>(G, parent);
    k_ccl_flatten<<<(unsigned)blocks, 256, 0, s>>>(G, parent, minlin, counter);
    SYK_CUDA(cudaGetLastError());
    unsigned long long n_roots = 0;
    SYK_CUDA(cudaMemcpyAsync(&n_roots, counter, sizeof(n_roots), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaStreamSynchronize(s));
    *n_labels_host = n_roots;
    int ret = SYK_OK;
    if (n_roots > 0) {
        SYK_CUDA(cudaMallocAsync((void **)&keys, sizeof(unsigned) * 4 * (size_t)n_roots, s));
        roots = keys + n_roots;
        keys2 = roots + n_roots;
        roots2 = keys2 + n_roots;
        SYK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
        k_ccl_collect_roots<<<(unsigned)blocks, 256, 0, s>>>(parent, minlin, total, keys, roots, counter, n_roots);
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, roots, roots2, (int)n_roots, 0, 32, s);
        SYK_CUDA(cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 16, s));
        cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, roots, roots2, (int)n_roots, 0, 32, s);
        // minlin is no longer needed: reuse it as label_of_root
        k_ccl_rank<<<(unsigned)((n_roots + 255) / 256), 256, 0, s>>>(roots2, n_roots, minlin);
    }
    k_ccl_write<<<(unsigned)blocks, 256, 0, s>>>(G, parent, minlin, labels_dev);
    if (cudaGetLastError() != cudaSuccess) {
        syk_set_error("syk_label_components: kernel launch failed");
        ret = SYK_ECUDA;
    }
    if (tmp) cudaFreeAsync(tmp, s);
    if (keys) cudaFreeAsync(keys, s);
    cudaFreeAsync(counter, s);
    cudaFreeAsync(minlin, s);
    cudaFreeAsync(parent, s);
    return ret;
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
