// syk_morph_vol.cu -- binary morphology of a whole thresholded volume with an arbitrary small structuring element
// (row f4 of SURVEY.md section 8: the step between the threshold and the connected components).
//
// Replaces apply_morphological_operations(tmp_data, morph_ops, mop_kwargs=dict(structure=struct)) as called by
// _object_segmentation_thread (syconn/extraction/object_extraction_steps.py:312-358) on the thresholded uint8 volume, i.e.
// syconn/proc/image.py:485-507 -> _multi_mop_findobjects (:358-437) with a single object of id 1:
//   * every op works inside the bounding box B of the current foreground (scipy.ndimage.find_objects);
//   * binary_dilation / binary_closing: the box is padded by n_iters zeros on every side (also beyond the volume), the
//     op runs there with border_value 0 -- dilation steps are clipped at the padded box, erosion steps see zeros outside
//     of it -- and the result is cropped back to B: nothing ever grows out of B;
//   * binary_erosion / binary_opening run on B itself with border_value 0;
//   * runs of equal ops were merged into one call with iterations = run length by the caller (_count_subsequent_mops).
// Voxels are bit-packed along the memory-contiguous axis (32 per word) in a buffer that extends the volume by the largest
// padding; one step of a word is the OR / AND over the structuring element's offsets, shifts along the packed axis being
// funnel shifts of three neighbouring words.  Bits outside the op's domain box are forced to zero after every step,
// which is both the clipping and the border value.  All integer / bitwise: results are identical to scipy's.
#include "syk_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_ROWS = 64;  // distinct (du, dv) rows of a structuring element

struct VolGeom {
    int n[3];          // volume extents, internal axes u, v, w (w = contiguous)
    long long st[3];   // element strides
    int pad;           // zero margin of the bit buffer (voxels; along w rounded up to whole words)
    int eu, ev, wpr;   // bit buffer extents: rows along u, v and words per row
    int wbit0;         // bit index of volume voxel w = 0 inside a buffer row
};

struct Structure {
    int nrows;
    signed char du[MAX_ROWS], dv[MAX_ROWS];
    unsigned long long wmask[MAX_ROWS];  // bit (dw + 31): offset (du, dv, dw) belongs to the element, |dw| <= 31
};

struct Box {  // domain of a step in buffer coordinates: rows [lo, hi) along u, v; bits [lo, hi) along w
    int lo[3], hi[3];
};

// volume -> bits (value != 0); *bad is raised by values other than 0 / 1.  A warp owns whole rows (index arithmetic once per
// row); VEC4: contiguous uint8 rows whose starts are 4-byte aligned -- a lane then converts four voxels per load and a
// warp 128 voxels (four words) per step.
template <typename T, bool VEC4>
__global__ void k_vol_pack(const T *__restrict__ vol, VolGeom G, unsigned *__restrict__ bits, int *__restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int nrows = G.n[0] * G.n[1];
    const int k0 = G.wbit0 >> 5;  // wbit0 is a multiple of 32: a volume word is a buffer word
    const int wstride = (gridDim.x * blockDim.x) >> 5;
    bool seen_bad = false;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += wstride) {
        const int u = row / G.n[1], v = row - u * G.n[1];
        const T *src = vol + (u * G.st[0] + v * G.st[1]);
        unsigned *dst = bits + ((long long)(u + G.pad) * G.ev + (v + G.pad)) * G.wpr + k0;
        if (VEC4) {
            for (int w0 = 0; w0 < G.n[2]; w0 += 128) {
                const int w = w0 + 4 * lane;
                unsigned q = 0u;  // n[2] % 4 == 0 on this path: a quad is inside the row or outside of it
                if (w < G.n[2]) q = *reinterpret_cast<const unsigned *>(reinterpret_cast<const unsigned char *>(src) + w);
                seen_bad |= (q & 0xFEFEFEFEu) != 0u;
                const unsigned nib = (q & 1u) | ((q >> 7) & 2u) | ((q >> 14) & 4u) | ((q >> 21) & 8u);
                unsigned part = nib << (4 * (lane & 7));  // the eight lanes of a group hold the eight nibbles of one word
                part |= __shfl_xor_sync(FULL, part, 1);
                part |= __shfl_xor_sync(FULL, part, 2);
                part |= __shfl_xor_sync(FULL, part, 4);
                const int kw = (w0 >> 5) + (lane >> 3);
                if ((lane & 7) == 0 && part && kw * 32 < G.n[2]) dst[kw] = part;  // the buffer starts zeroed
            }
        } else {
            for (int w0 = 0; w0 < G.n[2]; w0 += 32) {
                const int w = w0 + lane;
                T val = 0;
                if (w < G.n[2]) val = src[w * G.st[2]];
                seen_bad |= val > (T)1;
                const unsigned word = __ballot_sync(FULL, val != 0);
                if (lane == 0 && word) dst[w0 >> 5] = word;
            }
        }
    }
    if (seen_bad) *bad = 1;
}

// bits -> volume (0 / 1), every voxel; same work split as the packing
template <typename T, bool VEC4>
__global__ void k_vol_unpack(T *__restrict__ vol, VolGeom G, const unsigned *__restrict__ bits) {
    const int lane = threadIdx.x & 31;
    const int nrows = G.n[0] * G.n[1];
    const int k0 = G.wbit0 >> 5;
    const int wstride = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += wstride) {
        const int u = row / G.n[1], v = row - u * G.n[1];
        T *dstv = vol + (u * G.st[0] + v * G.st[1]);
        const unsigned *src = bits + ((long long)(u + G.pad) * G.ev + (v + G.pad)) * G.wpr + k0;
        if (VEC4) {
            for (int w0 = 0; w0 < G.n[2]; w0 += 128) {
                const int w = w0 + 4 * lane;
                if (w < G.n[2]) {
                    const unsigned nib = (src[w >> 5] >> (w & 31)) & 15u;
                    const unsigned q = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                    *reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(dstv) + w) = q;
                }
            }
        } else {
            for (int w0 = 0; w0 < G.n[2]; w0 += 32) {
                const int w = w0 + lane;
                if (w < G.n[2]) dstv[w * G.st[2]] = (T)((src[w0 >> 5] >> lane) & 1u);
            }
        }
    }
}

// bounding box of the set bits: box[0..2] = min (u, v, bit), box[3..5] = max (inclusive); untouched when empty
__global__ void k_bits_bbox(const unsigned *__restrict__ bits, VolGeom G, int *__restrict__ box) {
    const long long nwords = (long long)G.eu * G.ev * G.wpr;
    int lo0 = 1 << 30, lo1 = 1 << 30, lo2 = 1 << 30, hi0 = -1, hi1 = -1, hi2 = -1;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < nwords; x += (long long)gridDim.x * blockDim.x) {
        const unsigned word = bits[x];
        if (word) {
            const int k = (int)(x % G.wpr);
            const long long r = x / G.wpr;
            const int v = (int)(r % G.ev), u = (int)(r / G.ev);
            lo0 = min(lo0, u), hi0 = max(hi0, u);
            lo1 = min(lo1, v), hi1 = max(hi1, v);
            lo2 = min(lo2, k * 32 + __ffs(word) - 1), hi2 = max(hi2, k * 32 + 31 - __clz(word));
        }
    }
    // warp -> block (shared memory) -> one set of global atomics per block: thousands of warps on six addresses would serialise
    __shared__ int sb[6];
    if (threadIdx.x < 6) sb[threadIdx.x] = threadIdx.x < 3 ? (1 << 30) : -1;
    __syncthreads();
    lo0 = __reduce_min_sync(FULL, lo0), lo1 = __reduce_min_sync(FULL, lo1), lo2 = __reduce_min_sync(FULL, lo2);
    hi0 = __reduce_max_sync(FULL, hi0), hi1 = __reduce_max_sync(FULL, hi1), hi2 = __reduce_max_sync(FULL, hi2);
    if ((threadIdx.x & 31) == 0 && hi0 >= 0) {
        atomicMin(&sb[0], lo0), atomicMin(&sb[1], lo1), atomicMin(&sb[2], lo2);
        atomicMax(&sb[3], hi0), atomicMax(&sb[4], hi1), atomicMax(&sb[5], hi2);
    }
    __syncthreads();
    if (threadIdx.x < 6 && sb[3] >= 0) {
        if (threadIdx.x < 3) atomicMin(&box[threadIdx.x], sb[threadIdx.x]);
        else atomicMax(&box[threadIdx.x], sb[threadIdx.x]);
    }
}

// mask of the bits of word k that lie inside [lo, hi)
__device__ __forceinline__ unsigned span_mask(int k, int lo, int hi) {
    const int a = max(lo - k * 32, 0), b = min(hi - k * 32, 32);
    if (b <= a) return 0u;
    const unsigned upto_b = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    return upto_b & ~((1u << a) - 1u);  // a <= 31 here
}

// one dilation (ERODE = false: OR of the element's translates) or erosion (ERODE = true: AND) step over the domain box D;
// words outside of D become zero.  The element is symmetric (checked on the host), so scipy's reflection does not matter.
template <bool ERODE>
__global__ void k_bits_step(const unsigned *__restrict__ src, unsigned *__restrict__ dst, VolGeom G, Structure S, Box D) {
    const long long nwords = (long long)G.eu * G.ev * G.wpr;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < nwords; x += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(x % G.wpr);
        const long long r = x / G.wpr;
        const int v = (int)(r % G.ev), u = (int)(r / G.ev);
        unsigned dm = 0u;
        if (u >= D.lo[0] && u < D.hi[0] && v >= D.lo[1] && v < D.hi[1]) dm = span_mask(k, D.lo[2], D.hi[2]);
        unsigned acc = 0u;
        if (dm) {
            acc = ERODE ? 0xFFFFFFFFu : 0u;
            for (int i = 0; i < S.nrows; ++i) {
                const int uu = u + S.du[i], vv = v + S.dv[i];
                unsigned long long m = S.wmask[i];
                unsigned lo = 0u, mid = 0u, hi = 0u;
                if (uu >= 0 && uu < G.eu && vv >= 0 && vv < G.ev) {
                    const unsigned *row = src + ((long long)uu * G.ev + vv) * G.wpr;
                    mid = row[k];
                    if ((m & 0x7FFFFFFFull) && k > 0) lo = row[k - 1];  // only rows with offsets along the packed axis
                    if ((m >> 32) && k + 1 < G.wpr) hi = row[k + 1];
                }
                while (m) {
                    const int dw = __ffsll((long long)m) - 1 - 31;
                    m &= m - 1ull;
                    // bit b of the translate = source bit b + dw
                    const unsigned t = dw == 0 ? mid : dw > 0 ? __funnelshift_r(mid, hi, (unsigned)dw) : __funnelshift_l(lo, mid, (unsigned)(-dw));
                    acc = ERODE ? (acc & t) : (acc | t);
                }
                if (ERODE ? ((acc & dm) == 0u) : ((acc & dm) == dm)) break;  // saturated
            }
        }
        dst[x] = acc & dm;
    }
}

int launch_blocks(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

SYK_API int syk_binary_morph_ops(void *vol_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                 const uint8_t *structure_host, const int64_t structure_shape[3], const int32_t *ops_host,
                                 const int32_t *iters_host, int n_ops, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 1 || elem_bytes == 2 || elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 1, 2, 4 or 8");
    SYK_CHECK_ARG(shape && strides && structure_host && structure_shape && (n_ops == 0 || (ops_host && iters_host)), "NULL argument");
    SYK_CHECK_ARG(n_ops >= 0, "negative op count");
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0 || n_ops == 0) return SYK_OK;
    SYK_CHECK_ARG(vol_dev, "NULL volume");
    for (int a = 0; a < 3; ++a) {
        SYK_CHECK_ARG(structure_shape[a] >= 1 && structure_shape[a] % 2 == 1 && structure_shape[a] <= 63, "structure extents must be odd and <= 63");
        SYK_CHECK_ARG(shape[a] < (1 << 28), "volume extent too large");
    }
    int pad = 0;
    for (int i = 0; i < n_ops; ++i) {
        SYK_CHECK_ARG(ops_host[i] >= 0 && ops_host[i] <= 3, "op code must be 0 (erosion), 1 (dilation), 2 (opening) or 3 (closing)");
        SYK_CHECK_ARG(iters_host[i] >= 1 && iters_host[i] <= 4096, "iterations must be >= 1");
        if ((ops_host[i] == 1 || ops_host[i] == 3) && iters_host[i] > pad) pad = iters_host[i];
    }
    // internal axes: largest |stride| first, the contiguous axis is packed
    int ax[3] = {0, 1, 2};
    auto key = [&](int a) { return strides[a] < 0 ? -strides[a] : strides[a]; };
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (key(ax[j]) > key(ax[i])) {
                const int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    VolGeom G;
    for (int a = 0; a < 3; ++a) {
        G.n[a] = (int)shape[ax[a]];
        G.st[a] = strides[ax[a]];
    }
    SYK_CHECK_ARG((long long)G.n[0] * G.n[1] < (1ll << 31) - (1 << 20), "too many rows");
    G.pad = pad;
    G.wbit0 = ((pad + 31) / 32) * 32;  // word aligned: a volume word is a buffer word whenever the row length allows
    G.eu = G.n[0] + 2 * pad;
    G.ev = G.n[1] + 2 * pad;
    G.wpr = (G.wbit0 + G.n[2] + pad + 31) / 32;
    // structuring element -> rows (du, dv) with a bit mask of dw; must be point-symmetric
    const int64_t *ss = structure_shape;
    auto sat = [&](int64_t x, int64_t y, int64_t z) { return structure_host[(x * ss[1] + y) * ss[2] + z] != 0; };
    Structure S;
    S.nrows = 0;
    const int64_t es[3] = {ss[ax[0]], ss[ax[1]], ss[ax[2]]};
    bool any = false;
    for (int64_t a = 0; a < es[0]; ++a)
        for (int64_t b = 0; b < es[1]; ++b) {
            unsigned long long m = 0ull;
            for (int64_t c = 0; c < es[2]; ++c) {
                int64_t idx[3];
                idx[ax[0]] = a, idx[ax[1]] = b, idx[ax[2]] = c;
                const bool on = sat(idx[0], idx[1], idx[2]);
                const bool mirror = sat(ss[0] - 1 - idx[0], ss[1] - 1 - idx[1], ss[2] - 1 - idx[2]);
                SYK_CHECK_ARG(on == mirror, "the structuring element must be point-symmetric");
                if (on) m |= 1ull << (c - es[2] / 2 + 31);
            }
            if (m) {
                SYK_CHECK_ARG(S.nrows < MAX_ROWS, "structuring element has too many rows");
                S.du[S.nrows] = (signed char)(a - es[0] / 2);
                S.dv[S.nrows] = (signed char)(b - es[1] / 2);
                S.wmask[S.nrows] = m;
                ++S.nrows;
                any = true;
            }
        }
    SYK_CHECK_ARG(any, "empty structuring element");

    struct Scratch {
        cudaStream_t s;
        void *p[3] = {nullptr, nullptr, nullptr};
        ~Scratch() {
            for (void *x : p)
                if (x) cudaFreeAsync(x, s);
        }
    } sc{s};
    syk_pool_keep_warm();
    const long long nwords = (long long)G.eu * G.ev * G.wpr;
    SYK_CUDA(cudaMallocAsync(&sc.p[0], sizeof(unsigned) * (size_t)nwords, s));
    SYK_CUDA(cudaMallocAsync(&sc.p[1], sizeof(unsigned) * (size_t)nwords, s));
    SYK_CUDA(cudaMallocAsync(&sc.p[2], 8 * sizeof(int), s));
    unsigned *cur = (unsigned *)sc.p[0], *nxt = (unsigned *)sc.p[1];
    int *ctl = (int *)sc.p[2];  // [0..5] bounding box, [6] non-binary flag
    SYK_CUDA(cudaMemsetAsync(cur, 0, sizeof(unsigned) * (size_t)nwords, s));
    SYK_CUDA(cudaMemsetAsync(ctl, 0, 8 * sizeof(int), s));
    const int pb = launch_blocks((long long)G.n[0] * G.n[1] * 32, 256), wb = launch_blocks(nwords, 256);
    const bool vec4 = elem_bytes == 1 && G.st[2] == 1 && G.n[2] % 4 == 0 && ((uintptr_t)vol_dev & 3u) == 0 && G.st[0] % 4 == 0 &&
                      G.st[1] % 4 == 0;
    if (vec4) k_vol_pack<unsigned char, true><<<pb, 256, 0, s>>>((const unsigned char *)vol_dev, G, cur, ctl + 6);
    else if (elem_bytes == 1) k_vol_pack<unsigned char, false><<<pb, 256, 0, s>>>((const unsigned char *)vol_dev, G, cur, ctl + 6);
    else if (elem_bytes == 2) k_vol_pack<unsigned short, false><<<pb, 256, 0, s>>>((const unsigned short *)vol_dev, G, cur, ctl + 6);
    else if (elem_bytes == 4) k_vol_pack<unsigned, false><<<pb, 256, 0, s>>>((const unsigned *)vol_dev, G, cur, ctl + 6);
    else k_vol_pack<unsigned long long, false><<<pb, 256, 0, s>>>((const unsigned long long *)vol_dev, G, cur, ctl + 6);
    SYK_CUDA(cudaGetLastError());
    for (int i = 0; i < n_ops; ++i) {
        const int init[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};
        int hb[7];
        SYK_CUDA(cudaMemcpyAsync(ctl, init, sizeof(init), cudaMemcpyHostToDevice, s));
        k_bits_bbox<<<wb, 256, 0, s>>>(cur, G, ctl);
        SYK_CUDA(cudaGetLastError());
        SYK_CUDA(cudaMemcpyAsync(hb, ctl, sizeof(hb), cudaMemcpyDeviceToHost, s));
        SYK_CUDA(cudaStreamSynchronize(s));
        if (hb[6]) {
            syk_set_error("syk_binary_morph_ops: the volume holds values other than 0 and 1 (multi-label overlays are not supported)");
            return SYK_EINVAL;
        }
        if (hb[3] < 0) break;  // no foreground left: every further op is a no-op, like the reference's empty id loop
        const int op = ops_host[i], n = iters_host[i];
        Box B, D;
        for (int a = 0; a < 3; ++a) {
            B.lo[a] = hb[a];
            B.hi[a] = hb[3 + a] + 1;
            const int grow = (op == 1 || op == 3) ? n : 0;
            D.lo[a] = B.lo[a] - grow;  // stays inside the buffer: grow <= pad
            D.hi[a] = B.hi[a] + grow;
        }
        // sequence of steps: erosion (0), dilation (1), opening (2) = n erosions + n dilations, closing (3) = n dilations + n erosions
        const int n_steps = (op >= 2) ? 2 * n : n;
        for (int t = 0; t < n_steps; ++t) {
            const bool erode = op == 0 || (op == 2 && t < n) || (op == 3 && t >= n);
            const Box &dom = (t + 1 == n_steps) ? B : D;  // the last step also crops to the object's box
            if (erode) k_bits_step<true><<<wb, 256, 0, s>>>(cur, nxt, G, S, dom);
            else k_bits_step<false><<<wb, 256, 0, s>>>(cur, nxt, G, S, dom);
            unsigned *t2 = cur;
            cur = nxt;
            nxt = t2;
        }
        SYK_CUDA(cudaGetLastError());
    }
    {  // the non-binary check of a call whose loop ended early or whose last op needs no further bbox
        int flag = 0;
        SYK_CUDA(cudaMemcpyAsync(&flag, ctl + 6, sizeof(int), cudaMemcpyDeviceToHost, s));
        SYK_CUDA(cudaStreamSynchronize(s));
        if (flag) {
            syk_set_error("syk_binary_morph_ops: the volume holds values other than 0 and 1 (multi-label overlays are not supported)");
            return SYK_EINVAL;
        }
    }
    if (vec4) k_vol_unpack<unsigned char, true><<<pb, 256, 0, s>>>((unsigned char *)vol_dev, G, cur);
    else if (elem_bytes == 1) k_vol_unpack<unsigned char, false><<<pb, 256, 0, s>>>((unsigned char *)vol_dev, G, cur);
    else if (elem_bytes == 2) k_vol_unpack<unsigned short, false><<<pb, 256, 0, s>>>((unsigned short *)vol_dev, G, cur);
    else if (elem_bytes == 4) k_vol_unpack<unsigned, false><<<pb, 256, 0, s>>>((unsigned *)vol_dev, G, cur);
    else k_vol_unpack<unsigned long long, false><<<pb, 256, 0, s>>>((unsigned long long *)vol_dev, G, cur);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}
