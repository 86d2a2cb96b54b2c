// syk_morph.cu -- per-contact binary closing + dilation ("next" row f2 of SURVEY.md section 8).
//
// Replaces the per-id Python loop of the contact-site worker (syconn/extraction/cs_extraction_steps.py:439-461):
//   for every contact id, in the order of the id list: the id's mask inside its bounding box padded by n_closings
//   (clipped to the volume) goes through scipy.ndimage.binary_closing(iterations = n_closings) and
//   binary_dilation(iterations = cs_dilation), both with the 6-neighbourhood cross and border_value 0; the result is
//   written to the voxels of the box that are still background.
// The closed mask of an id only depends on the id's own voxels (no step ever writes or removes another id's label), so
// all ids are processed in parallel; the sequential "first id in list order wins a background voxel" rule is restored
// with an atomicMin of the id's list position into a per-voxel rank volume, resolved by one final pass.
//
// Masks are bit-packed along the memory-contiguous axis (32 voxels per word); one morphological step of a word is
// OR / AND of the word, its two funnel-shifted row neighbours and the four words above/below/before/behind it.
//   * boxes of up to SMALL_WORDS words (the typical contact): ONE CTA keeps both ping-pong bit buffers in shared memory
//     and does mask build, all 2 n + m steps and the write-back in a single launch;
//   * larger boxes: bit buffers in HBM, one launch per step over all words of all large boxes.
// All arithmetic is integer/bitwise; results are bit-identical to the reference loop run in the same id order.
#include <stdlib.h>

#include <vector>

#include "syk_common.cuh"

namespace {

#ifndef SYK_MORPH_BW
#define SYK_MORPH_BW 4
#endif
constexpr int MT = 256;              // threads per CTA
constexpr int SMALL_WORDS = 6136;    // 2 buffers x 24 KB of shared memory (minus the queue word: 48 KB static limit), 256 threads
constexpr int MID_WORDS = 3072;      // 2 buffers x 12 KB, 256 threads
constexpr int TINY_WORDS = 1024;     // 2 buffers x 4 KB, 128 threads: the typical contact (sizes include the zero halo)
constexpr unsigned NO_RANK = 0xFFFFFFFFu;

struct MorphBox {  // 64 bytes
    unsigned long long id;
    unsigned long long word0;  // first word of the box in the HBM bit buffers (large boxes)
    int lo[3];                 // padded, clipped box origin (internal axes u, v, w)
    int ext[3];                // padded, clipped box extents
    int ilo[3];                // the id's own bounding box relative to lo ...
    int ihi[3];                // ... (exclusive): the mask is empty outside of it
};
static_assert(sizeof(MorphBox) == 64, "MorphBox layout");

struct MorphCta {
    unsigned box;   // index into the box array
    unsigned base;  // first word (inside the box) of this CTA
};

struct MorphGeom {
    long long st[3];  // contact volume strides (elements) along the internal axes
    int n[3];         // contact volume extents along the internal axes
    int elem_bytes;
};

__device__ __forceinline__ unsigned long long ld_label(const void *vol, int elem_bytes, long long a) {
    return elem_bytes == 8 ? ((const unsigned long long *)vol)[a] : (unsigned long long)((const unsigned *)vol)[a];
}

// one dilation (ERODE = false) or erosion (ERODE = true) step of word (u, v, k); voxels outside the box count as 0
template <bool ERODE>
__device__ __forceinline__ unsigned morph_word(const unsigned *src, int u, int v, int k, int eu, int ev, int wpr, unsigned valid) {
    const int row = ev * wpr;
    const unsigned *p = src + ((long long)u * ev + v) * wpr + k;
    const unsigned c = p[0];
    const unsigned l = k > 0 ? p[-1] : 0u, r = k + 1 < wpr ? p[1] : 0u;
    const unsigned wl = __funnelshift_l(l, c, 1);  // bit i = voxel i - 1
    const unsigned wr = __funnelshift_r(c, r, 1);  // bit i = voxel i + 1 (bits past the box end are kept 0 in storage)
    const unsigned vm = v > 0 ? p[-wpr] : 0u, vp = v + 1 < ev ? p[wpr] : 0u;
    const unsigned um = u > 0 ? p[-row] : 0u, up = u + 1 < eu ? p[row] : 0u;
    if (ERODE) return c & wl & wr & vm & vp & um & up;
    return (c | wl | wr | vm | vp | um | up) & valid;
}

__device__ __forceinline__ unsigned valid_bits(int k, int ew) {
    const int left = ew - k * 32;
    return left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
}

// mask word (u, v, k) of the box: bit i = (contacts[lo + (u, v, 32 k + i)] == id).  Warp-cooperative: lane i tests voxel i.
__device__ __forceinline__ unsigned build_word(const void *vol, const MorphGeom &G, const MorphBox &B, int u, int v, int k, int lane) {
    if (u < B.ilo[0] || u >= B.ihi[0] || v < B.ilo[1] || v >= B.ihi[1] || k * 32 >= B.ihi[2] || k * 32 + 32 <= B.ilo[2]) return 0u;
    const int w = k * 32 + lane;
    bool hit = false;
    if (w >= B.ilo[2] && w < B.ihi[2])
        hit = ld_label(vol, G.elem_bytes, (long long)(B.lo[0] + u) * G.st[0] + (long long)(B.lo[1] + v) * G.st[1] +
                                              (long long)(B.lo[2] + w) * G.st[2]) == B.id;
    return __ballot_sync(0xFFFFFFFFu, hit);
}

// result word (u, v, k): every set voxel takes part in the "first id in list order" vote (fire-and-forget RED.MIN; the
// final pass only looks at the votes of voxels that are still background)
__device__ __forceinline__ void apply_word(const MorphGeom &G, const MorphBox &B, unsigned rank, unsigned bits, int u, int v, int k,
                                           int lane, unsigned *rankvol) {
    if (!((bits >> lane) & 1u)) return;
    const long long gu = B.lo[0] + u, gv = B.lo[1] + v, gw = B.lo[2] + k * 32 + lane;
    atomicMin(&rankvol[(gu * G.n[1] + gv) * G.n[2] + gw], rank);
}

// ---- small boxes: everything in one launch, bit buffers in shared memory ---------------------------------------------------
// WORDS = capacity of one bit buffer, NT = threads per CTA, WPR1 = every row of the box is a single word (box at most 32
// voxels wide along w).  The buffers carry a zero halo of one row / plane along v and u, so a step is five (WPR1) or
// seven loads without any bounds test.  Size classes are instantiated so that the typical contact (a few hundred words)
// runs with many CTAs per SM.
template <int WORDS, int NT, bool WPR1>
__global__ void __launch_bounds__(NT) k_morph_small(const void *__restrict__ vol, MorphGeom G, const MorphBox *__restrict__ boxes,
                                                    const unsigned *__restrict__ list, unsigned nlist, int n_close, int n_dil,
                                                    unsigned *__restrict__ rankvol, unsigned *__restrict__ queue) {
    __shared__ unsigned buf[2][WORDS];
    __shared__ unsigned s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // boxes differ a lot in size: after its first box a CTA takes the next one from a device counter (a static round robin
    // leaves a long tail)
    for (unsigned bi = blockIdx.x; bi < nlist;) {
        const unsigned rank = list[bi];
        const MorphBox B = boxes[rank];
        const int eu = B.ext[0], ev = B.ext[1], ew = B.ext[2];
        const int wpr = WPR1 ? 1 : (ew + 31) >> 5;
        const int rowp = wpr, planep = (ev + 2) * wpr;  // pitches of the haloed layout
        const int words = eu * ev * wpr, padded = (eu + 2) * planep;
        auto index_of = [&](int wi, int &u, int &v, int &k) {  // interior word wi -> coordinates and buffer index
            k = WPR1 ? 0 : wi % wpr;
            const int r = WPR1 ? wi : wi / wpr;
            v = r % ev;
            u = r / ev;
            return (u + 1) * planep + (v + 1) * rowp + k;
        };
        __syncthreads();  // the previous box is done with the buffers (and everybody has read s_next)
        if (tid == 0) s_next = gridDim.x + atomicAdd(queue, 1u);
        for (int i = tid; i < padded; i += NT) {
            buf[0][i] = 0u;
            buf[1][i] = 0u;
        }
        __syncthreads();
        bi = s_next;
        // mask of the id: only the rows of its own bounding box can hold voxels; BW words (loads) in flight per warp
        {
            constexpr int BW = SYK_MORPH_BW;
            const int k0 = B.ilo[2] >> 5, nk = ((B.ihi[2] + 31) >> 5) - k0;
            const int iv = B.ihi[1] - B.ilo[1];
            const int nin = (B.ihi[0] - B.ilo[0]) * iv * nk;
            for (int i0 = warp * BW; i0 < nin; i0 += (NT / 32) * BW) {
                bool hit[BW];
                int dst[BW];
#pragma unroll
                for (int j = 0; j < BW; ++j) {
                    const int i = i0 + j;
                    hit[j] = false;
                    dst[j] = -1;
                    if (i < nin) {
                        const int k = k0 + i % nk, r = i / nk;
                        const int v = B.ilo[1] + r % iv, u = B.ilo[0] + r / iv;
                        const int w = k * 32 + lane;
                        dst[j] = (u + 1) * planep + (v + 1) * rowp + k;
                        if (w >= B.ilo[2] && w < B.ihi[2])
                            hit[j] = ld_label(vol, G.elem_bytes, (long long)(B.lo[0] + u) * G.st[0] + (long long)(B.lo[1] + v) * G.st[1] +
                                                                     (long long)(B.lo[2] + w) * G.st[2]) == B.id;
                    }
                }
#pragma unroll
                for (int j = 0; j < BW; ++j) {
                    const unsigned b = __ballot_sync(0xFFFFFFFFu, hit[j]);
                    if (lane == 0 && dst[j] >= 0) buf[0][dst[j]] = b;
                }
            }
        }
        // step loop: a thread owns a column (v, k) and a contiguous range of planes u; marching along u keeps the words
        // below / at the current plane in registers (three loads per word when the row is a single word)
        const int ncol = ev * wpr;
        const int ncolp = ncol < NT ? ncol : NT;       // columns handled concurrently
        const int nph = NT / ncolp;                    // plane ranges per column
        const int ulen = (eu + nph - 1) / nph;
        const int c0 = tid % ncolp, ph = tid / ncolp;
        const int u_beg = ph * ulen, u_end = ph < nph ? min(eu, u_beg + ulen) : 0;
        __syncthreads();
        int cur = 0;
        const int steps = 2 * n_close + n_dil;
        const unsigned last_valid = valid_bits(wpr - 1, ew);
        for (int s = 0; s < steps; ++s) {
            const bool erode = s >= n_close && s < 2 * n_close;
            const unsigned *src = buf[cur];
            unsigned *dstb = buf[cur ^ 1];
            for (int c = c0; c < ncol && u_beg < u_end; c += ncolp) {  // one round unless the box has more than NT columns
                const int k = WPR1 ? 0 : c % wpr, v = WPR1 ? c : c / wpr;
                int idx = (u_beg + 1) * planep + (v + 1) * rowp + k;
                const unsigned vmask = (k + 1 == wpr) ? last_valid : 0xFFFFFFFFu;
                unsigned um = src[idx - planep], cc = src[idx];
                for (int u = u_beg; u < u_end; ++u, idx += planep) {
                    const unsigned up = src[idx + planep], vm = src[idx - rowp], vp = src[idx + rowp];
                    unsigned wl, wr;
                    if (WPR1) {
                        wl = cc << 1;
                        wr = cc >> 1;
                    } else {
                        wl = __funnelshift_l(k > 0 ? src[idx - 1] : 0u, cc, 1);
                        wr = __funnelshift_r(cc, k + 1 < wpr ? src[idx + 1] : 0u, 1);
                    }
                    dstb[idx] = erode ? (cc & wl & wr & vm & vp & um & up) : ((cc | wl | wr | vm | vp | um | up) & vmask);
                    um = cc;
                    cc = up;
                }
            }
            cur ^= 1;
            __syncthreads();
        }
        for (int w0 = warp * 32; w0 < words; w0 += NT) {
            int mu = 0, mv = 0, mk = 0;
            unsigned mine = 0u;
            if (w0 + lane < words) mine = buf[cur][index_of(w0 + lane, mu, mv, mk)];
            unsigned todo = __ballot_sync(0xFFFFFFFFu, mine != 0u);
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1u;
                const unsigned bits = __shfl_sync(0xFFFFFFFFu, mine, j);
                const int u = __shfl_sync(0xFFFFFFFFu, mu, j), v = __shfl_sync(0xFFFFFFFFu, mv, j), k = __shfl_sync(0xFFFFFFFFu, mk, j);
                apply_word(G, B, rank, bits, u, v, k, lane, rankvol);
            }
        }
    }
}

// ---- large boxes: bit buffers in HBM, one launch per step --------------------------------------------------------------------
__global__ void __launch_bounds__(MT) k_morph_build(const void *__restrict__ vol, MorphGeom G, const MorphBox *__restrict__ boxes,
                                                    const MorphCta *__restrict__ ctas, unsigned *__restrict__ dst) {
    const MorphCta c = ctas[blockIdx.x];
    const MorphBox B = boxes[c.box];
    const int ev = B.ext[1], wpr = (B.ext[2] + 31) >> 5;
    const long long words = (long long)B.ext[0] * ev * wpr;
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)c.base + (threadIdx.x >> 5) * 32;
    if (w0 >= words) return;
    const int nw = (int)min(32ll, words - w0);
    unsigned mine = 0u;
    for (int j = 0; j < nw; ++j) {
        const long long wi = w0 + j;
        const int k = (int)(wi % wpr);
        const long long r = wi / wpr;
        const unsigned b = build_word(vol, G, B, (int)(r / ev), (int)(r % ev), k, lane);
        if (lane == j) mine = b;
    }
    if (lane < nw) dst[B.word0 + w0 + lane] = mine;
}

template <bool ERODE>
__global__ void __launch_bounds__(MT) k_morph_step(const MorphBox *__restrict__ boxes, const MorphCta *__restrict__ ctas,
                                                   const unsigned *__restrict__ src, unsigned *__restrict__ dst) {
    const MorphCta c = ctas[blockIdx.x];
    const MorphBox B = boxes[c.box];
    const int eu = B.ext[0], ev = B.ext[1], wpr = (B.ext[2] + 31) >> 5;
    const long long words = (long long)eu * ev * wpr;
    const long long wi = (long long)c.base + threadIdx.x;
    if (wi >= words) return;
    const int k = (int)(wi % wpr);
    const long long r = wi / wpr;
    dst[B.word0 + wi] = morph_word<ERODE>(src + B.word0, (int)(r / ev), (int)(r % ev), k, eu, ev, wpr, valid_bits(k, B.ext[2]));
}

__global__ void __launch_bounds__(MT) k_morph_apply(MorphGeom G, const MorphBox *__restrict__ boxes,
                                                    const MorphCta *__restrict__ ctas, const unsigned *__restrict__ src,
                                                    unsigned *__restrict__ rankvol) {
    const MorphCta c = ctas[blockIdx.x];
    const MorphBox B = boxes[c.box];
    const int ev = B.ext[1], wpr = (B.ext[2] + 31) >> 5;
    const long long words = (long long)B.ext[0] * ev * wpr;
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)c.base + (threadIdx.x >> 5) * 32;
    if (w0 >= words) return;
    const int nw = (int)min(32ll, words - w0);
    const unsigned mine = lane < nw ? src[B.word0 + w0 + lane] : 0u;
    unsigned todo = __ballot_sync(0xFFFFFFFFu, mine != 0u);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1u;
        const unsigned bits = __shfl_sync(0xFFFFFFFFu, mine, j);
        const long long wi = w0 + j;
        const int k = (int)(wi % wpr);
        const long long r = wi / wpr;
        apply_word(G, B, c.box, bits, (int)(r / ev), (int)(r % ev), k, lane, rankvol);
    }
}

// background voxels covered by a closed mask take the id of the first such id in list order
__global__ void __launch_bounds__(MT) k_morph_final(void *__restrict__ vol, MorphGeom G, const unsigned *__restrict__ rankvol,
                                                    const MorphBox *__restrict__ boxes) {
    // the pass is a chain of dependent loads (vote -> label -> id of the winning box): four voxels per thread in flight
    const long long total = (long long)G.n[0] * G.n[1] * G.n[2];
    const long long stride = (long long)gridDim.x * MT;
    constexpr int ILP = 4;
    for (long long i0 = (long long)blockIdx.x * MT + threadIdx.x; i0 < total; i0 += ILP * stride) {
        unsigned r[ILP];
        long long a[ILP];
        unsigned long long cur[ILP], id[ILP];
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            const long long i = i0 + k * stride;
            r[k] = i < total ? rankvol[i] : NO_RANK;
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            cur[k] = 1ull;
            a[k] = 0;
            if (r[k] != NO_RANK) {
                const long long i = i0 + k * stride;
                const long long w = i % G.n[2], q = i / G.n[2];
                a[k] = (q / G.n[1]) * G.st[0] + (q % G.n[1]) * G.st[1] + w * G.st[2];
                cur[k] = ld_label(vol, G.elem_bytes, a[k]);
                id[k] = boxes[r[k]].id;
            }
        }
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (r[k] == NO_RANK || cur[k] != 0ull) continue;  // only background is ever written (cs_extraction_steps.py:460)
            if (G.elem_bytes == 8) ((unsigned long long *)vol)[a[k]] = id[k];
            else ((unsigned *)vol)[a[k]] = (unsigned)id[k];
        }
    }
}

struct Scratch {  // stream-ordered device scratch, released when the call returns
    void *p = nullptr;
    cudaStream_t s = nullptr;
    ~Scratch() {
        if (p) cudaFreeAsync(p, s);
    }
    cudaError_t alloc(size_t bytes, cudaStream_t st) {
        syk_pool_keep_warm();
        s = st;
        return cudaMallocAsync(&p, bytes ? bytes : 16, st);
    }
};


// large boxes in batches of at most BATCH_WORDS words per bit buffer (host-planned: CTA lists and word offsets)
static int run_large_boxes(void *cs_dev, const MorphGeom &G, std::vector<MorphBox> &boxes, const std::vector<unsigned> &large,
                           int n_closings, int n_dilations, unsigned *rankvol, cudaStream_t s) {
    const unsigned long long BATCH_WORDS = getenv("SYK_MORPH_BATCH") ? strtoull(getenv("SYK_MORPH_BATCH"), nullptr, 10) : (1ull << 26);
    size_t li = 0;
    std::vector<MorphCta> ctas;
    while (li < large.size()) {
        unsigned long long total = 0;
        ctas.clear();
        size_t lj = li;
        for (; lj < large.size(); ++lj) {
            MorphBox &B = boxes[large[lj]];
            const unsigned long long words = (unsigned long long)B.ext[0] * B.ext[1] * ((B.ext[2] + 31) / 32);
            if (lj > li && total + words > BATCH_WORDS) break;
            B.word0 = total;
            total += words;
            for (unsigned long long b = 0; b < words; b += MT) ctas.push_back(MorphCta{large[lj], (unsigned)b});
        }
        li = lj;
        Scratch bx, ct, bufA, bufB;  // word0 is per batch: upload the boxes with this batch's offsets
        SYK_CUDA(bx.alloc(boxes.size() * sizeof(MorphBox), s));
        SYK_CUDA(cudaMemcpyAsync(bx.p, boxes.data(), boxes.size() * sizeof(MorphBox), cudaMemcpyHostToDevice, s));
        SYK_CUDA(ct.alloc(ctas.size() * sizeof(MorphCta), s));
        SYK_CUDA(cudaMemcpyAsync(ct.p, ctas.data(), ctas.size() * sizeof(MorphCta), cudaMemcpyHostToDevice, s));
        SYK_CUDA(bufA.alloc(total * 4, s));
        SYK_CUDA(bufB.alloc(total * 4, s));
        const unsigned grid = (unsigned)ctas.size();
        unsigned *a = (unsigned *)bufA.p, *b = (unsigned *)bufB.p;
        const MorphBox *dbx = (const MorphBox *)bx.p;
        const MorphCta *dct = (const MorphCta *)ct.p;
        k_morph_build<<<grid, MT, 0, s>>>(cs_dev, G, dbx, dct, a);
        for (int it = 0; it < 2 * n_closings + n_dilations; ++it) {
            if (it >= n_closings && it < 2 * n_closings) k_morph_step<true><<<grid, MT, 0, s>>>(dbx, dct, a, b);
            else k_morph_step<false><<<grid, MT, 0, s>>>(dbx, dct, a, b);
            unsigned *t = a;
            a = b;
            b = t;
        }
        k_morph_apply<<<grid, MT, 0, s>>>(G, dbx, dct, a, rankvol);
        SYK_CUDA(cudaGetLastError());
        SYK_CUDA(cudaStreamSynchronize(s));  // the host vectors of this batch are reused by the next one
    }
    return SYK_OK;
}

// the six shared-memory size classes: lst[c] = device list of the cnt[c] boxes of class c
static int run_small_boxes(void *cs_dev, const MorphGeom &G, const MorphBox *dbx, const unsigned *const lst[6], const size_t cnt[6],
                           int n_closings, int n_dilations, unsigned *rk, int sms, cudaStream_t s) {
    Scratch d_queue;  // one box counter per class (dynamic hand-out after the first wave)
    SYK_CUDA(d_queue.alloc(8 * sizeof(unsigned), s));
    SYK_CUDA(cudaMemsetAsync(d_queue.p, 0, 8 * sizeof(unsigned), s));
    unsigned *queue = (unsigned *)d_queue.p;
#define SYK_MORPH_LAUNCH(c, WORDS, NT, WPR1, PER_SM)                                                                              \
    if (cnt[c]) {                                                                                                                \
        const unsigned long long n = cnt[c];                                                                                     \
        const unsigned long long grid = n < (unsigned long long)sms * PER_SM ? n : (unsigned long long)sms * PER_SM;             \
        k_morph_small<WORDS, NT, WPR1><<<(unsigned)grid, NT, 0, s>>>(cs_dev, G, dbx, lst[c], (unsigned)n, n_closings, n_dilations, \
                                                                     rk, queue + c);                                             \
    }
    SYK_MORPH_LAUNCH(0, TINY_WORDS, 128, true, 64)
    SYK_MORPH_LAUNCH(1, TINY_WORDS, 128, false, 64)
    SYK_MORPH_LAUNCH(2, MID_WORDS, 256, true, 32)
    SYK_MORPH_LAUNCH(3, MID_WORDS, 256, false, 32)
    SYK_MORPH_LAUNCH(4, SMALL_WORDS, MT, true, 16)
    SYK_MORPH_LAUNCH(5, SMALL_WORDS, MT, false, 16)
#undef SYK_MORPH_LAUNCH
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

// device-side planning for syk_close_contacts_records: record i (ids ascending) -> MorphBox i, appended to the list of its
// size class (class 6 = too large for shared memory).  err: a box outside the volume or id 0.
__global__ void k_morph_plan(const syk_record_t *__restrict__ recs, unsigned n, MorphGeom G, int a0, int a1, int a2, int n_close,
                             long long small_words, MorphBox *__restrict__ boxes, unsigned *__restrict__ lists,
                             unsigned *__restrict__ counts, int *__restrict__ err) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const syk_record_t r = recs[i];
    const int ax[3] = {a0, a1, a2};
    MorphBox B;
    B.id = r.id;
    B.word0 = 0;
    bool bad = r.id == 0ull;
    for (int a = 0; a < 3; ++a) {
        const long long mn = r.bb_min[ax[a]], mx = r.bb_max[ax[a]];
        bad |= !(mn >= 0 && mn < mx && mx <= G.n[a]);
        const long long lo = mn - n_close < 0 ? 0 : mn - n_close;
        const long long hi = mx + n_close > G.n[a] ? G.n[a] : mx + n_close;
        B.lo[a] = (int)lo;
        B.ext[a] = (int)(hi - lo);
        B.ilo[a] = (int)(mn - lo);
        B.ihi[a] = (int)(mx - lo);
    }
    if (bad) {
        *err = 1;
        return;
    }
    boxes[i] = B;
    const long long wpr = (B.ext[2] + 31) / 32;
    const long long padded = (long long)(B.ext[0] + 2) * (B.ext[1] + 2) * wpr;
    const int c = padded > small_words ? 6 : (padded <= TINY_WORDS ? 0 : padded <= MID_WORDS ? 2 : 4) + (wpr == 1 ? 0 : 1);
    lists[(size_t)c * n + atomicAdd(&counts[c], 1u)] = i;
}

}  // namespace

SYK_API int syk_close_contacts(void *cs_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3], const uint64_t *ids_host,
                               const int32_t *bbox_host, uint64_t n_ids, int n_closings, int n_dilations, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(shape && strides, "NULL argument");
    SYK_CHECK_ARG(n_closings >= 0 && n_dilations >= 0 && n_closings <= 64 && n_dilations <= 64, "iterations must be in 0..64");
    SYK_CHECK_ARG(n_ids < 0xFFFFFFFFull, "too many ids");
    if (n_ids == 0 || (n_closings == 0 && n_dilations == 0)) return SYK_OK;
    SYK_CHECK_ARG(cs_dev && ids_host && bbox_host, "NULL argument");
    for (int a = 0; a < 3; ++a) SYK_CHECK_ARG(shape[a] > 0 && shape[a] < (1ll << 30), "bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    // internal axes: u = slowest, w = memory-contiguous (bit-packing / coalescing axis)
    int ax[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (llabs(strides[ax[j]]) > llabs(strides[ax[i]])) {
                const int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    MorphGeom G;
    for (int a = 0; a < 3; ++a) {
        G.st[a] = strides[ax[a]];
        G.n[a] = (int)shape[ax[a]];
    }
    G.elem_bytes = elem_bytes;

    // test hooks: SYK_MORPH_SMALL lowers the shared-memory threshold (words), SYK_MORPH_BATCH the HBM batch size (words)
    long long small_words = SMALL_WORDS;
    if (const char *e = getenv("SYK_MORPH_SMALL")) {
        const long long v = atoll(e);
        if (v >= 0 && v < small_words) small_words = v;
    }
    std::vector<MorphBox> boxes(n_ids);
    std::vector<unsigned> cls[6];  // shared-memory classes: {tiny, mid, small} x {one word per row, several}
    std::vector<unsigned> large;
    for (uint64_t i = 0; i < n_ids; ++i) {
        MorphBox &B = boxes[i];
        B.id = ids_host[i];
        B.word0 = 0;
        SYK_CHECK_ARG(B.id != 0, "id 0 in the id list");
        for (int a = 0; a < 3; ++a) {
            const long long mn = bbox_host[i * 6 + ax[a]], mx = bbox_host[i * 6 + 3 + ax[a]];
            SYK_CHECK_ARG(mn >= 0 && mn < mx && mx <= G.n[a], "bounding box outside the volume");
            const long long lo = mn - n_closings < 0 ? 0 : mn - n_closings;              // cs_extraction_steps.py:443-444
            const long long hi = mx + n_closings > G.n[a] ? G.n[a] : mx + n_closings;    // :445, clipped by the slice
            B.lo[a] = (int)lo;
            B.ext[a] = (int)(hi - lo);
            B.ilo[a] = (int)(mn - lo);
            B.ihi[a] = (int)(mx - lo);
        }
        const long long wpr = (B.ext[2] + 31) / 32;
        const long long padded = (long long)(B.ext[0] + 2) * (B.ext[1] + 2) * wpr;  // with the zero halo along u and v
        if (padded > small_words) large.push_back((unsigned)i);
        else cls[(padded <= TINY_WORDS ? 0 : padded <= MID_WORDS ? 2 : 4) + (wpr == 1 ? 0 : 1)].push_back((unsigned)i);
    }
    const long long nvox = (long long)G.n[0] * G.n[1] * G.n[2];
    Scratch d_boxes, d_rank, d_small;
    SYK_CUDA(d_rank.alloc((size_t)nvox * 4, s));
    SYK_CUDA(cudaMemsetAsync(d_rank.p, 0xFF, (size_t)nvox * 4, s));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    rc = run_large_boxes(cs_dev, G, boxes, large, n_closings, n_dilations, (unsigned *)d_rank.p, s);
    if (rc) return rc;
    SYK_CUDA(d_boxes.alloc(boxes.size() * sizeof(MorphBox), s));
    SYK_CUDA(cudaMemcpyAsync(d_boxes.p, boxes.data(), boxes.size() * sizeof(MorphBox), cudaMemcpyHostToDevice, s));
    {
        std::vector<unsigned> all;
        size_t first[7] = {0};
        for (int c = 0; c < 6; ++c) {
            all.insert(all.end(), cls[c].begin(), cls[c].end());
            first[c + 1] = all.size();
        }
        if (!all.empty()) {
            SYK_CUDA(d_small.alloc(all.size() * sizeof(unsigned), s));
            SYK_CUDA(cudaMemcpyAsync(d_small.p, all.data(), all.size() * sizeof(unsigned), cudaMemcpyHostToDevice, s));
            const unsigned *lst[6];
            size_t cnt[6];
            for (int c = 0; c < 6; ++c) {
                lst[c] = (const unsigned *)d_small.p + first[c];
                cnt[c] = first[c + 1] - first[c];
            }
            rc = run_small_boxes(cs_dev, G, (const MorphBox *)d_boxes.p, lst, cnt, n_closings, n_dilations, (unsigned *)d_rank.p, sms, s);
            if (rc) return rc;
        }
    }
    k_morph_final<<<sms * 8, MT, 0, s>>>(cs_dev, G, (const unsigned *)d_rank.p, (const MorphBox *)d_boxes.p);
    SYK_CUDA(cudaGetLastError());
    SYK_CUDA(cudaStreamSynchronize(s));  // `boxes` / `small` are pageable host memory owned by this call
    return SYK_OK;
}


SYK_API int syk_close_contacts_records(void *cs_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                       const syk_record_t *records_dev, uint64_t n_ids, int n_closings, int n_dilations, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(shape && strides, "NULL argument");
    SYK_CHECK_ARG(n_closings >= 0 && n_dilations >= 0 && n_closings <= 64 && n_dilations <= 64, "iterations must be in 0..64");
    SYK_CHECK_ARG(n_ids < 0x7FFFFFFFull, "too many ids");
    if (n_ids == 0 || (n_closings == 0 && n_dilations == 0)) return SYK_OK;
    SYK_CHECK_ARG(cs_dev && records_dev, "NULL argument");
    for (int a = 0; a < 3; ++a) SYK_CHECK_ARG(shape[a] > 0 && shape[a] < (1ll << 30), "bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    int ax[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (llabs(strides[ax[j]]) > llabs(strides[ax[i]])) {
                const int t = ax[i];
                ax[i] = ax[j];
                ax[j] = t;
            }
    MorphGeom G;
    for (int a = 0; a < 3; ++a) {
        G.st[a] = strides[ax[a]];
        G.n[a] = (int)shape[ax[a]];
    }
    G.elem_bytes = elem_bytes;
    long long small_words = SMALL_WORDS;
    if (const char *e = getenv("SYK_MORPH_SMALL")) {
        const long long v = atoll(e);
        if (v >= 0 && v < small_words) small_words = v;
    }
    const unsigned n = (unsigned)n_ids;
    const long long nvox = (long long)G.n[0] * G.n[1] * G.n[2];
    Scratch d_boxes, d_lists, d_ctl, d_rank;
    SYK_CUDA(d_boxes.alloc((size_t)n * sizeof(MorphBox), s));
    SYK_CUDA(d_lists.alloc((size_t)n * 7 * sizeof(unsigned), s));
    SYK_CUDA(d_ctl.alloc(8 * sizeof(unsigned), s));
    SYK_CUDA(d_rank.alloc((size_t)nvox * 4, s));
    SYK_CUDA(cudaMemsetAsync(d_ctl.p, 0, 8 * sizeof(unsigned), s));
    SYK_CUDA(cudaMemsetAsync(d_rank.p, 0xFF, (size_t)nvox * 4, s));
    unsigned *counts = (unsigned *)d_ctl.p;
    k_morph_plan<<<(n + 255) / 256, 256, 0, s>>>(records_dev, n, G, ax[0], ax[1], ax[2], n_closings, small_words, (MorphBox *)d_boxes.p,
                                                 (unsigned *)d_lists.p, counts, (int *)(counts + 7));
    SYK_CUDA(cudaGetLastError());
    unsigned h[8];
    SYK_CUDA(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, s));  // 32 bytes: the class sizes set the grids
    SYK_CUDA(cudaStreamSynchronize(s));
    if (h[7]) {
        syk_set_error("invalid argument: a record with id 0 or a bounding box outside the volume");
        return SYK_EINVAL;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (h[6]) {  // rare: boxes beyond the shared-memory classes are planned on the host
        std::vector<MorphBox> boxes(n);
        std::vector<unsigned> large(h[6]);
        SYK_CUDA(cudaMemcpyAsync(boxes.data(), d_boxes.p, (size_t)n * sizeof(MorphBox), cudaMemcpyDeviceToHost, s));
        SYK_CUDA(cudaMemcpyAsync(large.data(), (unsigned *)d_lists.p + (size_t)6 * n, (size_t)h[6] * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
        SYK_CUDA(cudaStreamSynchronize(s));
        rc = run_large_boxes(cs_dev, G, boxes, large, n_closings, n_dilations, (unsigned *)d_rank.p, s);
        if (rc) return rc;
    }
    const unsigned *lst[6];
    size_t cnt[6];
    for (int c = 0; c < 6; ++c) {  // the class lists live n entries apart
        lst[c] = (const unsigned *)d_lists.p + (size_t)c * n;
        cnt[c] = h[c];
    }
    rc = run_small_boxes(cs_dev, G, (const MorphBox *)d_boxes.p, lst, cnt, n_closings, n_dilations, (unsigned *)d_rank.p, sms, s);
    if (rc) return rc;
    k_morph_final<<<sms * 8, MT, 0, s>>>(cs_dev, G, (const unsigned *)d_rank.p, (const MorphBox *)d_boxes.p);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}
