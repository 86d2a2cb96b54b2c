// syk_host.cu -- synthetic label generator and the HOST-buffer entry points (one synchronous call per chunk, the shape
// of the reference's Cython `def` functions; host<->device copies happen inside the call).
#include <stdlib.h>

#include "syk_common.cuh"

namespace {

// ---- synthetic labels: integer-only, coordinate hashed => identical bits on CPU (syconn_b200/synth.py) and GPU ----
__host__ __device__ __forceinline__ long long tri(long long t, long long P) {
    long long m = t % (2 * P);
    m -= P;
    return m < 0 ? -m : m;
}

__host__ __device__ __forceinline__ unsigned long long synth_label(long long x, long long y, long long z, int px, int py,
                                                                   int pz, int amp, unsigned long long seed, int kind,
                                                                   int density16, int elem_bytes) {
    const long long B = 1ll << 20;
    const long long X = x + B, Y = y + B, Z = z + B;
    const long long xw = X + (((tri(Y, 41) + tri(Z, 29)) * amp) >> 4);
    const long long yw = Y + (((tri(Z, 37) + tri(X, 43)) * amp) >> 4);
    const long long zw = Z + (((tri(X, 31) + tri(Y, 47)) * amp) >> 5);
    const unsigned long long cx = (unsigned long long)(xw / px), cy = (unsigned long long)(yw / py),
                             cz = (unsigned long long)(zw / pz);
    const unsigned long long h =
        syk_mix64((cx * 0x9E3779B97F4A7C15ULL) ^ (cy * 0xC2B2AE3D27D4EB4FULL) ^ (cz * 0x165667B19E3779F9ULL) ^
                  (seed * 0xD6E8FEB86659FD93ULL + (unsigned long long)kind * 0xA0761D6478BD642FULL));
    if (kind == 0) {
        if ((h & 31ull) == 0ull) return 0ull;
        unsigned long long id = h >> 32;
        return id ? id : 1ull;
    }
    if (((h >> 8) & 15ull) >= (unsigned long long)density16) return 0ull;
    return elem_bytes == 8 ? (h | 1ull) : ((h >> 32) | 1ull);
}

__global__ void k_synth(void *out, int elem_bytes, long long nx, long long ny, long long nz, long long sx, long long sy,
                        long long sz, long long ox, long long oy, long long oz, int px, int py, int pz, int amp,
                        unsigned long long seed, int kind, int density16, int fast_axis) {
    const long long total = nx * ny * nz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long x, y, z;
        if (fast_axis == 2) {  // z fastest in memory
            z = i % nz;
            const long long r = i / nz;
            y = r % ny;
            x = r / ny;
        } else {  // x fastest in memory
            x = i % nx;
            const long long r = i / nx;
            y = r % ny;
            z = r / ny;
        }
        const unsigned long long v = synth_label(x + ox, y + oy, z + oz, px, py, pz, amp, seed, kind, density16, elem_bytes);
        const long long a = x * sx + y * sy + z * sz;
        if (elem_bytes == 8) ((unsigned long long *)out)[a] = v;
        else ((unsigned *)out)[a] = (unsigned)v;
    }
}

// Device scratch of the host entry points comes from the stream-ordered pool (kept warm between calls: a call per
// chunk must not pay cudaMalloc/cudaFree round trips for gigabyte buffers).
// Everything a host call enqueues goes to the calling thread's own stream (syk_host_stream): concurrent calls from
// several worker threads overlap their uploads, kernels and downloads.
struct DevBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    ~DevBuf() {
        if (p) cudaFreeAsync(p, s);
    }
};
static cudaError_t dev_alloc(DevBuf &b, size_t bytes, cudaStream_t s) {
    syk_pool_keep_warm();
    b.s = s;
    return cudaMallocAsync(&b.p, bytes ? bytes : 16, s);
}
#define SYK_H2D(dst, src, bytes, s) SYK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s))
// device -> host, complete on return
#define SYK_D2H(dst, src, bytes, s)                                            \
    do {                                                                       \
        SYK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s)); \
        SYK_CUDA(syk_stream_wait(s));                                    \
    } while (0)

// a dense block viewed through `strides`: nbytes = elem_bytes * prod(shape); base offset must be 0
static int check_dense(const int64_t *shape, const int64_t *strides, int nd) {
    // sort by stride, check that strides are the running products
    int idx[4] = {0, 1, 2, 3};
    for (int i = 0; i < nd; ++i)
        for (int j = i + 1; j < nd; ++j)
            if (strides[idx[j]] < strides[idx[i]]) {
                int t = idx[i];
                idx[i] = idx[j];
                idx[j] = t;
            }
    int64_t expect = 1;
    for (int i = 0; i < nd; ++i) {
        const int a = idx[i];
        if (shape[a] == 1) continue;
        if (strides[a] != expect) {
            syk_set_error("host array must be a dense block (permuted C layout); make it contiguous first");
            return SYK_EINVAL;
        }
        expect *= shape[a];
    }
    return SYK_OK;
}

static uint64_t pick_capacity(uint64_t hint, uint64_t nvox) {
    uint64_t c = hint ? hint * 2 : (nvox / 64 < (1u << 16) ? (1u << 16) : nvox / 64);
    if (c > nvox * 2 + 1024) c = nvox * 2 + 1024;
    return c;
}

static int export_to_host(syk_table_t *t, const syk_chunk_geom_t *geom, syk_record_t **out, uint64_t *n_out, cudaStream_t hs) {
    uint64_t n = 0;
    int ovf = 0;
    int rc = syk_table_count(t, hs, &n, &ovf);
    if (rc) return rc;
    if (ovf) {
        syk_set_error("id table overflow");
        return SYK_EOVERFLOW;
    }
    *n_out = n;
    *out = nullptr;
    if (n == 0) return SYK_OK;
    DevBuf recs;
    SYK_CUDA(dev_alloc(recs, n * sizeof(syk_record_t), hs));
    rc = syk_table_export(t, geom, 1, (syk_record_t *)recs.p, n, &n, hs);
    if (rc) return rc;
    *out = (syk_record_t *)malloc(n * sizeof(syk_record_t));
    if (!*out) return SYK_ENOMEM;
    SYK_D2H(*out, recs.p, n * sizeof(syk_record_t), hs);
    return SYK_OK;
}

// (label0 << 32) | label1 per voxel of a dense [X, Y, Z, 2] label block (0 where label0 == 0), written in C order
__global__ void k_pack_pairs(const unsigned *__restrict__ labels, long long nx, long long ny, long long nz, long long sx, long long sy,
                             long long sz, long long sc, unsigned long long *__restrict__ out) {
    const long long total = nx * ny * nz;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long z = i % nz;
        const long long r = i / nz;
        const long long y = r % ny;
        const long long x = r / ny;
        const long long a = x * sx + y * sy + z * sz;
        const unsigned l0 = labels[a], l1 = labels[a + sc];
        out[i] = l0 ? (((unsigned long long)l0 << 32) | (unsigned long long)l1) : 0ull;
    }
}

// dense uint32 labels of a device block of n ids (retry with a larger table on overflow); ids.p[label - 1] = id
static int relabel_block(const void *vol_dev, int elem_bytes, uint64_t n, DevBuf &labels, DevBuf &ids, uint64_t *n_ids, cudaStream_t hs) {
    SYK_CUDA(dev_alloc(labels, n * sizeof(uint32_t), hs));
    uint64_t cap = n / 32 < (1u << 16) ? (1u << 16) : n / 32;
    for (;;) {
        syk_table_t *t = nullptr;
        int rc = syk_table_create_on(&t, cap, hs);
        if (rc) return rc;
        if (ids.p) {
            cudaFreeAsync(ids.p, hs);
            ids.p = nullptr;
        }
        SYK_CUDA(dev_alloc(ids, t->capacity * sizeof(uint64_t), hs));
        rc = syk_dense_relabel(t, vol_dev, elem_bytes, n, (uint32_t *)labels.p, (uint64_t *)ids.p, t->capacity, n_ids, hs);
        syk_table_destroy(t);
        if (rc != SYK_EOVERFLOW || cap >= n * 2) return rc;
        cap *= 4;
    }
}

}  // namespace

SYK_API int syk_synth_labels(void *out_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                             const int64_t origin[3], const int32_t pitch[3], int32_t warp_amp, uint64_t seed, int kind,
                             int density16, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(pitch[0] > 0 && pitch[1] > 0 && pitch[2] > 0, "pitch must be positive");
    const long long total = shape[0] * shape[1] * shape[2];
    if (total == 0) return SYK_OK;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const int fast_axis = (strides[0] < strides[2]) ? 0 : 2;
    k_synth<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out_dev, elem_bytes, shape[0], shape[1], shape[2], strides[0],
                                                               strides[1], strides[2], origin[0], origin[1], origin[2],
                                                               pitch[0], pitch[1], pitch[2], warp_amp, seed, kind, density16,
                                                               fast_axis);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

SYK_API int syk_find_object_properties_host(const void *labels_host, int elem_bytes, const int64_t shape[3],
                                            const int64_t strides[3], uint64_t capacity_hint, syk_record_t **records_out,
                                            uint64_t *n_out) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(records_out && n_out, "output pointers are NULL");
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    *records_out = nullptr;
    *n_out = 0;
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if (nvox == 0) return SYK_OK;
    rc = check_dense(shape, strides, 3);
    if (rc) return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf lab;
    SYK_CUDA(dev_alloc(lab, nvox * elem_bytes, hs));
    SYK_H2D(lab.p, labels_host, nvox * elem_bytes, hs);
    const int64_t origin[3] = {0, 0, 0};
    syk_chunk_geom_t geom;
    for (int a = 0; a < 3; ++a) {
        geom.origin[a] = 0;
        geom.shape[a] = shape[a];
    }
    uint64_t cap = pick_capacity(capacity_hint, nvox);
    for (;;) {
        syk_table_t *t = nullptr;
        rc = syk_table_create_on(&t, cap, hs);
        if (rc) return rc;
        rc = syk_find_object_properties(t, lab.p, elem_bytes, shape, strides, origin, 0, hs);
        if (!rc) rc = export_to_host(t, &geom, records_out, n_out, hs);
        syk_table_destroy(t);
        if (rc != SYK_EOVERFLOW || cap >= nvox * 2) return rc;
        cap *= 4;
    }
}

SYK_API int syk_map_subcell_extract_props_host(const void *cell_host, const int64_t cell_strides[3], const void *subcell_host,
                                               const int64_t sub_strides[4], int n_sub, int elem_bytes, const int64_t shape[3],
                                               int props_too, uint64_t capacity_hint, syk_record_t **cell_records_out,
                                               uint64_t *n_cell_out, syk_record_t **sub_records_out, uint64_t *n_sub_out,
                                               syk_pair_t **pairs_out, uint64_t *n_pairs_out) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(n_sub >= 0 && n_sub <= 4, "n_sub must be in [0, 4] per call");
    if (cell_records_out) *cell_records_out = nullptr;
    if (n_cell_out) *n_cell_out = 0;
    for (int c = 0; c < n_sub; ++c) {
        if (sub_records_out) sub_records_out[c] = nullptr;
        if (n_sub_out) n_sub_out[c] = 0;
        pairs_out[c] = nullptr;
        n_pairs_out[c] = 0;
    }
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if (nvox == 0) return SYK_OK;
    rc = check_dense(shape, cell_strides, 3);
    if (rc) return rc;
    if (n_sub) {
        const int64_t sshape[4] = {n_sub, shape[0], shape[1], shape[2]};
        rc = check_dense(sshape, sub_strides, 4);
        if (rc) return rc;
    }
    cudaStream_t hs = syk_host_stream();
    DevBuf cell, sub;
    SYK_CUDA(dev_alloc(cell, nvox * elem_bytes, hs));
    SYK_H2D(cell.p, cell_host, nvox * elem_bytes, hs);
    const void *subp[4] = {nullptr, nullptr, nullptr, nullptr};
    if (n_sub) {
        SYK_CUDA(dev_alloc(sub, nvox * elem_bytes * n_sub, hs));
        SYK_H2D(sub.p, subcell_host, nvox * elem_bytes * n_sub, hs);
        for (int c = 0; c < n_sub; ++c) subp[c] = (const char *)sub.p + (size_t)c * sub_strides[0] * elem_bytes;
    }
    const int64_t origin[3] = {0, 0, 0};
    syk_chunk_geom_t geom;
    for (int a = 0; a < 3; ++a) {
        geom.origin[a] = 0;
        geom.shape[a] = shape[a];
    }
    uint64_t cap = pick_capacity(capacity_hint, nvox);
    for (;;) {
        syk_table_t *ct = nullptr, *st[4] = {nullptr, nullptr, nullptr, nullptr};
        syk_pairs_t *pt[4] = {nullptr, nullptr, nullptr, nullptr};
        rc = SYK_OK;
        if (props_too) rc = syk_table_create_on(&ct, cap, hs);
        for (int c = 0; c < n_sub && !rc; ++c) {
            if (props_too) rc = syk_table_create_on(&st[c], cap, hs);
            if (!rc) rc = syk_pairs_create_on(&pt[c], cap, hs);
        }
        if (!rc)
            rc = syk_map_subcell_extract_props(props_too ? ct : nullptr, props_too ? st : nullptr, pt, cell.p, cell_strides, subp,
                                               sub_strides + 1, n_sub, elem_bytes, shape, origin, 0, hs);
        if (!rc && props_too) rc = export_to_host(ct, &geom, cell_records_out, n_cell_out, hs);
        for (int c = 0; c < n_sub && !rc; ++c) {
            if (props_too) rc = export_to_host(st[c], &geom, &sub_records_out[c], &n_sub_out[c], hs);
            if (rc) break;
            uint64_t np = 0;
            DevBuf pb;
            const uint64_t maxp = pt[c]->capacity;
            SYK_CUDA(dev_alloc(pb, maxp * sizeof(syk_pair_t), hs));
            rc = syk_pairs_export(pt[c], (syk_pair_t *)pb.p, maxp, &np, hs);
            if (rc) break;
            n_pairs_out[c] = np;
            if (np) {
                pairs_out[c] = (syk_pair_t *)malloc(np * sizeof(syk_pair_t));
                SYK_D2H(pairs_out[c], pb.p, np * sizeof(syk_pair_t), hs);
            }
        }
        syk_table_destroy(ct);
        for (int c = 0; c < n_sub; ++c) {
            syk_table_destroy(st[c]);
            syk_pairs_destroy(pt[c]);
        }
        if (rc != SYK_EOVERFLOW || cap >= nvox * 2) {
            if (rc) {  // release partial results
                if (cell_records_out && *cell_records_out) { free(*cell_records_out); *cell_records_out = nullptr; }
                for (int c = 0; c < n_sub; ++c) {
                    if (sub_records_out && sub_records_out[c]) { free(sub_records_out[c]); sub_records_out[c] = nullptr; }
                    if (pairs_out[c]) { free(pairs_out[c]); pairs_out[c] = nullptr; }
                }
            }
            return rc;
        }
        if (cell_records_out && *cell_records_out) { free(*cell_records_out); *cell_records_out = nullptr; }
        for (int c = 0; c < n_sub; ++c) {
            if (sub_records_out && sub_records_out[c]) { free(sub_records_out[c]); sub_records_out[c] = nullptr; }
            if (pairs_out[c]) { free(pairs_out[c]); pairs_out[c] = nullptr; }
        }
        cap *= 4;
    }
}

static int cs_host_impl(const void *edges_host, int edge_bytes, const int64_t *edge_strides, const void *arr_host, int elem_bytes,
                        const int64_t shape[3], const int64_t strides[3], const int32_t stencil[3], uint64_t *out_host,
                        syk_record_t **records_out = nullptr, uint64_t *n_out = nullptr) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    for (int a = 0; a < 3; ++a)
        SYK_CHECK_ARG(stencil[a] >= 1 && (stencil[a] % 2) == 1, "stencil must be odd along every axis (block_processing_C.pyx:57)");
    int64_t oshape[3];
    uint64_t nout = 1;
    for (int a = 0; a < 3; ++a) {
        oshape[a] = shape[a] - stencil[a] + 1;
        if (oshape[a] <= 0) return SYK_OK;
        nout *= (uint64_t)oshape[a];
    }
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    rc = check_dense(shape, strides, 3);
    if (rc) return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf arr, edg, out;
    SYK_CUDA(dev_alloc(arr, nvox * elem_bytes, hs));
    SYK_H2D(arr.p, arr_host, nvox * elem_bytes, hs);
    if (edges_host) {
        rc = check_dense(shape, edge_strides, 3);
        if (rc) return rc;
        SYK_CUDA(dev_alloc(edg, nvox * edge_bytes, hs));
        SYK_H2D(edg.p, edges_host, nvox * edge_bytes, hs);
    }
    SYK_CUDA(dev_alloc(out, nout * 8, hs));
    const int64_t ost[3] = {oshape[1] * oshape[2], oshape[2], 1};
    if (edges_host)
        rc = syk_process_block_nonzero(edg.p, edge_bytes, edge_strides, arr.p, elem_bytes, strides, shape, stencil, (uint64_t *)out.p,
                                       ost, hs);
    else
        rc = syk_detect_cs(arr.p, elem_bytes, shape, strides, stencil, (uint64_t *)out.p, ost, hs);
    if (rc) return rc;
    // the contact volume goes home while its properties are computed (second stream of the thread would be needed for a
    // true overlap inside one call; across worker threads the copies overlap anyway)
    if (records_out) {  // properties of the contact volume while it is still on the device (cs_extraction_steps.py:439)
        const int64_t origin[3] = {0, 0, 0};
        syk_chunk_geom_t geom;
        for (int a = 0; a < 3; ++a) {
            geom.origin[a] = 0;
            geom.shape[a] = oshape[a];
        }
        uint64_t cap = pick_capacity(0, nout);
        for (;;) {
            syk_table_t *t = nullptr;
            rc = syk_table_create_on(&t, cap, hs);
            if (rc) return rc;
            rc = syk_find_object_properties(t, out.p, 8, oshape, ost, origin, 0, hs);
            if (!rc) rc = export_to_host(t, &geom, records_out, n_out, hs);
            syk_table_destroy(t);
            if (rc != SYK_EOVERFLOW || cap >= nout * 2) break;
            cap *= 4;
        }
        if (rc) return rc;
    }
    SYK_D2H(out_host, out.p, nout * 8, hs);
    return SYK_OK;
}

SYK_API int syk_detect_cs_props_host(const void *arr_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                     const int32_t stencil[3], uint64_t *out_host, syk_record_t **records_out, uint64_t *n_out) {
    SYK_CHECK_ARG(records_out && n_out, "output pointers are NULL");
    *records_out = nullptr;
    *n_out = 0;
    return cs_host_impl(nullptr, 0, nullptr, arr_host, elem_bytes, shape, strides, stencil, out_host, records_out, n_out);
}

SYK_API int syk_detect_cs_host(const void *arr_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                               const int32_t stencil[3], uint64_t *out_host) {
    return cs_host_impl(nullptr, 0, nullptr, arr_host, elem_bytes, shape, strides, stencil, out_host);
}

SYK_API int syk_process_block_nonzero_host(const void *edges_host, int edge_bytes, const int64_t edge_strides[3],
                                           const void *arr_host, int elem_bytes, const int64_t arr_strides[3],
                                           const int64_t shape[3], const int32_t stencil[3], uint64_t *out_host) {
    SYK_CHECK_ARG(edges_host != nullptr && edge_strides != nullptr, "edges is NULL");
    SYK_CHECK_ARG(edge_bytes == 1 || edge_bytes == 4, "edge_bytes must be 1 or 4");
    return cs_host_impl(edges_host, edge_bytes, edge_strides, arr_host, elem_bytes, shape, arr_strides, stencil, out_host);
}

SYK_API int syk_extract_cs_syntype_host(const void *cs_host, int elem_bytes, const int64_t shape[3], const int64_t cs_strides[3],
                                        const uint8_t *syn_host, const int64_t syn_strides[3], const uint8_t *asym_host,
                                        const int64_t asym_strides[3], const uint8_t *sym_host, const int64_t sym_strides[3],
                                        syk_record_t **cs_records_out, uint64_t *n_cs_out, syk_synvox_t **vox_out, uint64_t *n_vox_out) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(cs_records_out && n_cs_out && vox_out && n_vox_out, "output pointers are NULL");
    *cs_records_out = nullptr;
    *vox_out = nullptr;
    *n_cs_out = *n_vox_out = 0;
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if (nvox == 0) return SYK_OK;
    if ((rc = check_dense(shape, cs_strides, 3)) || (rc = check_dense(shape, syn_strides, 3)) ||
        (rc = check_dense(shape, asym_strides, 3)) || (rc = check_dense(shape, sym_strides, 3)))
        return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf cs, syn, asym, sym, vox, cnt;
    SYK_CUDA(dev_alloc(cs, nvox * elem_bytes, hs));
    SYK_CUDA(dev_alloc(syn, nvox, hs));
    SYK_CUDA(dev_alloc(asym, nvox, hs));
    SYK_CUDA(dev_alloc(sym, nvox, hs));
    SYK_CUDA(dev_alloc(cnt, 16, hs));
    SYK_H2D(cs.p, cs_host, nvox * elem_bytes, hs);
    SYK_H2D(syn.p, syn_host, nvox, hs);
    SYK_H2D(asym.p, asym_host, nvox, hs);
    SYK_H2D(sym.p, sym_host, nvox, hs);
    const int64_t origin[3] = {0, 0, 0};
    syk_chunk_geom_t geom;
    for (int a = 0; a < 3; ++a) {
        geom.origin[a] = 0;
        geom.shape[a] = shape[a];
    }
    uint64_t cap = pick_capacity(0, nvox);
    uint64_t max_vox = nvox / 16 + 4096;
    for (;;) {
        syk_table_t *t = nullptr;
        rc = syk_table_create_on(&t, cap, hs);
        if (rc) return rc;
        if (vox.p) {
            cudaFreeAsync(vox.p, hs);
            vox.p = nullptr;
        }
        SYK_CUDA(dev_alloc(vox, max_vox * sizeof(syk_synvox_t), hs));
        SYK_CUDA(cudaMemsetAsync(cnt.p, 0, 16, hs));
        rc = syk_extract_cs_syntype(t, cs.p, elem_bytes, shape, cs_strides, (const uint8_t *)syn.p, syn_strides, (const uint8_t *)asym.p,
                                    asym_strides, (const uint8_t *)sym.p, sym_strides, origin, 0, (syk_synvox_t *)vox.p, max_vox,
                                    (uint64_t *)cnt.p, hs);
        unsigned long long nv = 0;
        if (!rc) SYK_D2H(&nv, cnt.p, sizeof(nv), hs);
        if (!rc && nv > max_vox) {  // voxel buffer too small: retry with the exact size
            syk_table_destroy(t);
            max_vox = nv;
            continue;
        }
        if (!rc) rc = export_to_host(t, &geom, cs_records_out, n_cs_out, hs);
        syk_table_destroy(t);
        if (rc == SYK_EOVERFLOW && cap < nvox * 2) {
            cap *= 4;
            continue;
        }
        if (rc) return rc;
        *n_vox_out = nv;
        if (nv) {
            *vox_out = (syk_synvox_t *)malloc(nv * sizeof(syk_synvox_t));
            if (!*vox_out) return SYK_ENOMEM;
            SYK_D2H(*vox_out, vox.p, nv * sizeof(syk_synvox_t), hs);
        }
        return SYK_OK;
    }
}

SYK_API int syk_detect_seg_boundaries_host(const void *arr_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                           uint8_t *out_host) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if (nvox == 0) return SYK_OK;
    rc = check_dense(shape, strides, 3);
    if (rc) return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf arr, out;
    SYK_CUDA(dev_alloc(arr, nvox * elem_bytes, hs));
    SYK_H2D(arr.p, arr_host, nvox * elem_bytes, hs);
    SYK_CUDA(dev_alloc(out, nvox, hs));
    rc = syk_detect_seg_boundaries(arr.p, elem_bytes, shape, strides, (uint8_t *)out.p, hs);
    if (rc) return rc;
    SYK_D2H(out_host, out.p, nvox, hs);
    return SYK_OK;
}

SYK_API int syk_detect_contact_partners_host(const void *edges_host, int edge_bytes, const int64_t edge_strides[3], const void *arr_host,
                                             int elem_bytes, const int64_t strides[3], const int64_t shape[3], const int32_t stencil[3],
                                             uint64_t *out_host) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(arr_host && out_host && shape && strides && stencil, "NULL argument");
    for (int a = 0; a < 3; ++a)
        SYK_CHECK_ARG(stencil[a] >= 1 && (stencil[a] % 2) == 1, "stencil must be odd along every axis");
    int64_t oshape[3];
    uint64_t nout = 1;
    for (int a = 0; a < 3; ++a) {
        oshape[a] = shape[a] - stencil[a] + 1;
        if (oshape[a] <= 0) return SYK_OK;
        nout *= (uint64_t)oshape[a];
    }
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if ((rc = check_dense(shape, strides, 3))) return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf arr, labels, ids, edg, packed, out;
    SYK_CUDA(dev_alloc(arr, nvox * elem_bytes, hs));
    SYK_H2D(arr.p, arr_host, nvox * elem_bytes, hs);
    const void *lab = arr.p;
    if (elem_bytes == 8) {  // 64-bit ids -> dense uint32 labels (a bijection keeps boundaries, counts and scan order)
        uint64_t n_ids = 0;
        if ((rc = relabel_block(arr.p, 8, nvox, labels, ids, &n_ids, hs))) return rc;
        lab = labels.p;
    }
    if (edges_host) {
        SYK_CHECK_ARG(edge_bytes == 1 || edge_bytes == 4, "edge_bytes must be 1 or 4");
        if ((rc = check_dense(shape, edge_strides, 3))) return rc;
        SYK_CUDA(dev_alloc(edg, nvox * edge_bytes, hs));
        SYK_H2D(edg.p, edges_host, nvox * edge_bytes, hs);
    }
    SYK_CUDA(dev_alloc(packed, nout * 8, hs));
    SYK_CUDA(dev_alloc(out, nout * 16, hs));
    const int64_t ost[3] = {oshape[1] * oshape[2], oshape[2], 1};
    rc = syk_detect_contact_partners(edg.p, edge_bytes, edge_strides, lab, strides, shape, stencil, (uint64_t *)packed.p, ost, hs);
    if (!rc) rc = syk_cs64_unpack((const uint64_t *)packed.p, nout, (const uint64_t *)ids.p, (uint64_t *)out.p, hs);
    if (rc) return rc;
    SYK_D2H(out_host, out.p, nout * 16, hs);
    return SYK_OK;
}

SYK_API int syk_find_object_properties_cs_64bit_host(const uint64_t *cs_host, const int64_t shape[3], const int64_t strides[4],
                                                     syk_record_t **records_out, uint64_t **partners_out, uint64_t *n_out) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(cs_host && shape && strides && records_out && partners_out && n_out, "NULL argument");
    *records_out = nullptr;
    *partners_out = nullptr;
    *n_out = 0;
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if (nvox == 0) return SYK_OK;
    const int64_t shape4[4] = {shape[0], shape[1], shape[2], 2};
    if ((rc = check_dense(shape4, strides, 4))) return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf cs, labels, ids, packed;
    SYK_CUDA(dev_alloc(cs, nvox * 16, hs));
    SYK_H2D(cs.p, cs_host, nvox * 16, hs);
    uint64_t n_ids = 0;
    if ((rc = relabel_block(cs.p, 8, nvox * 2, labels, ids, &n_ids, hs))) return rc;
    SYK_CUDA(dev_alloc(packed, nvox * 8, hs));
    long long blocks = (long long)((nvox + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_pack_pairs<<<(unsigned)blocks, 256, 0, hs>>>((const unsigned *)labels.p, shape[0], shape[1], shape[2], strides[0], strides[1],
                                                   strides[2], strides[3], (unsigned long long *)packed.p);
    SYK_CUDA(cudaGetLastError());
    const int64_t origin[3] = {0, 0, 0};
    const int64_t pst[3] = {shape[1] * shape[2], shape[2], 1};
    syk_chunk_geom_t geom;
    for (int a = 0; a < 3; ++a) {
        geom.origin[a] = 0;
        geom.shape[a] = shape[a];
    }
    uint64_t cap = pick_capacity(0, nvox);
    for (;;) {
        syk_table_t *t = nullptr;
        if ((rc = syk_table_create_on(&t, cap, hs))) return rc;
        rc = syk_find_object_properties(t, packed.p, 8, shape, pst, origin, 0, hs);
        if (!rc) rc = export_to_host(t, &geom, records_out, n_out, hs);
        syk_table_destroy(t);
        if (rc != SYK_EOVERFLOW || cap >= nvox * 2) break;
        cap *= 4;
    }
    if (rc || *n_out == 0) return rc;
    uint64_t *idh = (uint64_t *)malloc((n_ids ? n_ids : 1) * sizeof(uint64_t));
    uint64_t *pr = (uint64_t *)malloc(*n_out * 2 * sizeof(uint64_t));
    if (!idh || !pr) {
        free(idh);
        free(pr);
        free(*records_out);
        *records_out = nullptr;
        *n_out = 0;
        return SYK_ENOMEM;
    }
    cudaError_t e = cudaMemcpyAsync(idh, ids.p, n_ids * sizeof(uint64_t), cudaMemcpyDeviceToHost, hs);
    if (e == cudaSuccess) e = syk_stream_wait(hs);
    if (e != cudaSuccess) {
        syk_set_error("copy of the id list failed: %s", cudaGetErrorString(e));
        free(idh);
        free(pr);
        free(*records_out);
        *records_out = nullptr;
        *n_out = 0;
        return SYK_ECUDA;
    }
    for (uint64_t i = 0; i < *n_out; ++i) {  // packed dense labels -> the two partner ids of the record
        const uint64_t k = (*records_out)[i].id;
        const uint32_t l0 = (uint32_t)(k >> 32), l1 = (uint32_t)k;
        pr[2 * i] = l0 ? idh[l0 - 1] : 0;
        pr[2 * i + 1] = l1 ? idh[l1 - 1] : 0;
    }
    free(idh);
    *partners_out = pr;
    return SYK_OK;
}

SYK_API int syk_close_contacts_host(void *cs_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                    const uint64_t *ids_host, const int32_t *bbox_host, uint64_t n_ids, int n_closings,
                                    int n_dilations) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    SYK_CHECK_ARG(cs_host && shape && strides, "NULL argument");
    const uint64_t nvox = (uint64_t)shape[0] * shape[1] * shape[2];
    if (nvox == 0 || n_ids == 0) return SYK_OK;
    if ((rc = check_dense(shape, strides, 3))) return rc;
    cudaStream_t hs = syk_host_stream();
    DevBuf cs;
    SYK_CUDA(dev_alloc(cs, nvox * elem_bytes, hs));
    SYK_H2D(cs.p, cs_host, nvox * elem_bytes, hs);
    if ((rc = syk_close_contacts(cs.p, elem_bytes, shape, strides, ids_host, bbox_host, n_ids, n_closings, n_dilations, hs))) return rc;
    SYK_D2H(cs_host, cs.p, nvox * elem_bytes, hs);
    return SYK_OK;
}
