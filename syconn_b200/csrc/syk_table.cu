// syk_table.cu -- device hash tables of per-id records / overlap pairs: create, clear, export, merge, bucket.
// Replaces the reference's Python dict plumbing between chunks and workers:
//   merge_prop_dicts  syconn/proc/sd_proc.py:1248-1273     -> syk_table_merge_records
//   merge_map_dicts   syconn/proc/sd_proc.py:1300-1322     -> syk_pairs_merge
//   id -> reducer hash  syconn/reps/rep_helper.py:143-163  -> syk_records_bucket / syk_pairs_bucket
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include <cub/device/device_radix_sort.cuh>

#include "syk_common.cuh"

// ---- errors ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void syk_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int syk_require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        syk_set_error("no CUDA device available (%s); libsyk has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return SYK_ENODEV;
    }
    return SYK_OK;
}

// Tables and host-call scratch come from the stream-ordered pool and are kept there between calls: a call per chunk
// must not pay cudaMalloc/cudaFree round trips for buffers of hundreds of megabytes.
void syk_pool_keep_warm() {
    static bool done = false;
    if (done) return;
    done = true;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        // keep up to SYK_POOL_KEEP_MB (default 32768, i.e. 32 of the 180 GB) of freed scratch in the pool; beyond that it goes back to the driver
        // so that other allocators in the process (e.g. torch's caching allocator) are not starved
        const char *e = getenv("SYK_POOL_KEEP_MB");
        unsigned long long thr = (unsigned long long)(e ? atoll(e) : 32768) << 20;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
}

int syk_ensure_dyn_smem(const void *func, int bytes) {
    static std::mutex mu;
    static std::unordered_map<const void *, int> cur[16];
    int dev = 0;
    SYK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int &have = cur[dev & 15][func];
    if (bytes > have) {
        SYK_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        have = bytes;
    }
    return SYK_OK;
}

typedef CUresult (*SykEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool syk_make_tmap3(CUtensorMap *m, const void *base, int elem_bytes, const long long n[3], const long long st[3], int box_w,
                    int box_v) {
    static SykEncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (SykEncodeTiledFn)p;
        cudaGetLastError();
    });
    if (!fn || getenv("SYK_NO_TMA")) return false;
    if (st[2] != 1 || ((uintptr_t)base & 15) || box_w > 256 || box_v > 256 || ((box_w * elem_bytes) & 15)) return false;
    for (int a = 0; a < 2; ++a)
        if (st[a] <= 0 || ((st[a] * elem_bytes) & 15) || st[a] * elem_bytes >= (1ll << 40)) return false;
    for (int a = 0; a < 3; ++a)
        if (n[a] <= 0 || n[a] >= (1ll << 31)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)n[2], (cuuint64_t)n[1], (cuuint64_t)n[0]};
    cuuint64_t strides[2] = {(cuuint64_t)(st[1] * elem_bytes), (cuuint64_t)(st[0] * elem_bytes)};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_v, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    return fn(m, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims,
              strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t syk_stream_wait(cudaStream_t s) {
    // Default: cudaStreamSynchronize (the driver spins; measured fastest for the *_host calls at 1 .. 8 ranks per box).
    // SYK_YIELD_WAIT=1: poll cudaStreamQuery and yield the core in between -- for hosts with more waiting threads than cores.
    static const bool yield = getenv("SYK_YIELD_WAIT") != nullptr;
    if (!yield) return cudaStreamSynchronize(s);
    cudaError_t e;
    while ((e = cudaStreamQuery(s)) == cudaErrorNotReady) sched_yield();
    return e;
}

cudaStream_t syk_host_stream() {
    static thread_local cudaStream_t streams[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return (cudaStream_t)0;
    if (!streams[dev] && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        return (cudaStream_t)0;
    }
    return streams[dev];
}

SYK_API int syk_version(void) { return SYK_VERSION; }
SYK_API const char *syk_last_error(void) { return g_err; }
SYK_API int syk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
SYK_API int syk_set_device(int device) {
    int rc = syk_require_device();
    if (rc) return rc;
    SYK_CUDA(cudaSetDevice(device));
    return SYK_OK;
}
SYK_API void syk_free(void *p) { free(p); }

static uint64_t round_pow2(uint64_t c) {
    uint64_t p = 1024;
    while (p < c) p <<= 1;
    return p;
}

// ---- id tables ------------------------------------------------------------------------------------------------
int syk_table_create_on(syk_table **out, uint64_t capacity, cudaStream_t s) {
    SYK_CHECK_ARG(out != nullptr, "out is NULL");
    int rc = syk_require_device();
    if (rc) return rc;
    syk_table *t = (syk_table *)calloc(1, sizeof(syk_table));
    if (!t) return SYK_ENOMEM;
    t->capacity = round_pow2(capacity);
    t->stream = s;
    SYK_CUDA(cudaGetDevice(&t->device));
    syk_pool_keep_warm();
    void *ctl = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&t->slots, t->capacity * sizeof(SykSlot), s);
    if (e == cudaSuccess) e = cudaMallocAsync(&ctl, 64, s);
    if (e != cudaSuccess) {
        syk_set_error("cudaMalloc of %llu table slots failed: %s", (unsigned long long)t->capacity, cudaGetErrorString(e));
        if (t->slots) cudaFreeAsync(t->slots, s);
        free(t);
        cudaGetLastError();
        return SYK_ENOMEM;
    }
    t->flags = (int *)ctl;
    t->counter = (unsigned long long *)((char *)ctl + 32);
    SYK_CUDA(cudaMemsetAsync(t->slots, 0, t->capacity * sizeof(SykSlot), s));
    SYK_CUDA(cudaMemsetAsync(ctl, 0, 64, s));
    // the table is used from whatever stream the caller passes later (torch side streams do not synchronise with the
    // creating stream): creation is rare, so simply finish the allocation + clear here
    SYK_CUDA(cudaStreamSynchronize(s));
    *out = t;
    return SYK_OK;
}
SYK_API int syk_table_create(syk_table_t **out, uint64_t capacity) { return syk_table_create_on(out, capacity, (cudaStream_t)0); }

SYK_API int syk_table_destroy(syk_table_t *t) {
    if (!t) return SYK_OK;
    cudaDeviceSynchronize();  // kernels on any stream may still use the slots; destruction is rare
    cudaGetLastError();
    cudaFreeAsync(t->slots, t->stream);
    cudaFreeAsync(t->flags, t->stream);
    free(t);
    return SYK_OK;
}

SYK_API uint64_t syk_table_capacity(const syk_table_t *t) { return t ? t->capacity : 0; }

SYK_API int syk_table_clear(syk_table_t *t, void *stream) {
    SYK_CHECK_ARG(t != nullptr, "table is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    SYK_CUDA(cudaMemsetAsync(t->slots, 0, t->capacity * sizeof(SykSlot), s));
    SYK_CUDA(cudaMemsetAsync(t->flags, 0, 4 * sizeof(int), s));
    return SYK_OK;
}

__global__ void k_table_count(const SykSlot *slots, uint64_t cap, unsigned long long *counter) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n = 0;
    for (; i < cap; i += (uint64_t)gridDim.x * blockDim.x) n += slots[i].key != 0ull;
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, n);
}

SYK_API int syk_table_count(syk_table_t *t, void *stream, uint64_t *n_out, int *overflow_out) {
    SYK_CHECK_ARG(t != nullptr, "table is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    SYK_CUDA(cudaMemsetAsync(t->counter, 0, sizeof(unsigned long long), s));
    int blocks = (int)((t->capacity + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_table_count<<<blocks, 256, 0, s>>>(t->slots, t->capacity, t->counter);
    unsigned long long n = 0;
    int fl = 0;
    SYK_CUDA(cudaMemcpyAsync(&n, t->counter, sizeof(n), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaMemcpyAsync(&fl, t->flags, sizeof(fl), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaStreamSynchronize(s));
    if (n_out) *n_out = n;
    if (overflow_out) *overflow_out = fl;
    return SYK_OK;
}

// n_geoms == 0xFFFFFFFF: geoms[0] applies to every chunk_seq (single-chunk export)
__device__ __forceinline__ void decode_rep(syk_record_t &r, const syk_chunk_geom_t *geoms, uint32_t n_geoms) {
    uint32_t seq = (uint32_t)(r.rep_key >> 40);
    r.chunk_seq = seq;
    if (n_geoms == 0xFFFFFFFFu && geoms != nullptr) seq = 0;
    else if (geoms == nullptr || seq >= n_geoms) {
        r.rep[0] = r.rep[1] = r.rep[2] = -1;
        return;
    }
    const syk_chunk_geom_t g = geoms[seq];
    unsigned long long lin = SYK_REP_MASK - (r.rep_key & SYK_REP_MASK);
    unsigned long long z = lin % (unsigned long long)g.shape[2];
    unsigned long long xy = lin / (unsigned long long)g.shape[2];
    unsigned long long y = xy % (unsigned long long)g.shape[1];
    unsigned long long x = xy / (unsigned long long)g.shape[1];
    r.rep[0] = (int32_t)((long long)x + g.origin[0]);
    r.rep[1] = (int32_t)((long long)y + g.origin[1]);
    r.rep[2] = (int32_t)((long long)z + g.origin[2]);
}

// min_vx > 1: the worker's small-object drop (syconn/proc/sd_proc.py:650-661, :667-680): an object that lies purely inside the
// chunk (its box touches none of the six chunk faces -- equivalent to "its id is on no face") and has fewer than min_vx
// voxels is not reported.  `geoms[0]` must then be the chunk of the call.
__device__ __forceinline__ bool small_inside(const syk_record_t &r, const syk_chunk_geom_t &g, unsigned long long min_vx) {
    if (r.count >= min_vx) return false;
    for (int a = 0; a < 3; ++a)
        if ((long long)r.bb_min[a] <= g.origin[a] || (long long)r.bb_max[a] >= g.origin[a] + g.shape[a]) return false;
    return true;
}

__global__ void k_table_export(const SykSlot *slots, uint64_t cap, syk_record_t *out, unsigned long long max_out,
                               unsigned long long *counter, const syk_chunk_geom_t *geoms, uint32_t n_geoms,
                               unsigned long long min_vx = 0ull) {
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t base = i0 - lane; base < cap; base += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i = base + lane;
        SykSlot s;
        syk_record_t r;
        bool occ = false;
        if (i < cap) {
            s = slots[i];
            occ = s.key != 0ull;
        }
        if (occ) {
            r.id = s.key;
            r.count = s.count;
            r.rep_key = s.rep_enc - 1ull;
            for (int a = 0; a < 3; ++a) {
                r.bb_min[a] = (int32_t)((0xFFFFFFFFu - s.min_enc[a]) - SYK_COORD_BIAS);
                r.bb_max[a] = (int32_t)(s.max_enc[a] - SYK_COORD_BIAS);
            }
            if (min_vx > 1ull && geoms != nullptr && small_inside(r, geoms[0], min_vx)) occ = false;
        }
        unsigned m = __ballot_sync(0xffffffffu, occ);
        if (!m) continue;
        unsigned long long pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(counter, (unsigned long long)__popc(m));
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        if (occ) {
            unsigned long long pos = pos0 + __popc(m & ((1u << lane) - 1u));
            if (pos < max_out) {
                decode_rep(r, geoms, n_geoms);
                out[pos] = r;
            }
        }
    }
}

static int upload_geoms(const syk_chunk_geom_t *geoms_host, uint32_t n_geoms, cudaStream_t s, syk_chunk_geom_t **dev) {
    *dev = nullptr;
    if (!geoms_host || n_geoms == 0) return SYK_OK;
    SYK_CUDA(cudaMallocAsync((void **)dev, sizeof(syk_chunk_geom_t) * n_geoms, s));
    SYK_CUDA(cudaMemcpyAsync(*dev, geoms_host, sizeof(syk_chunk_geom_t) * n_geoms, cudaMemcpyHostToDevice, s));
    return SYK_OK;
}

SYK_API int syk_table_export(syk_table_t *t, const syk_chunk_geom_t *geoms_host, uint32_t n_geoms, syk_record_t *records_dev,
                             uint64_t max_records, uint64_t *n_out, void *stream) {
    SYK_CHECK_ARG(t != nullptr, "table is NULL");
    SYK_CHECK_ARG(records_dev != nullptr || max_records == 0, "records_dev is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    syk_chunk_geom_t *gd = nullptr;
    int rc = upload_geoms(geoms_host, n_geoms, s, &gd);
    if (rc) return rc;
    SYK_CUDA(cudaMemsetAsync(t->counter, 0, sizeof(unsigned long long), s));
    int blocks = (int)((t->capacity + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_table_export<<<blocks, 256, 0, s>>>(t->slots, t->capacity, records_dev, max_records, t->counter, gd, n_geoms);
    SYK_CUDA(cudaGetLastError());
    unsigned long long n = 0;
    int fl = 0;
    SYK_CUDA(cudaMemcpyAsync(&n, t->counter, sizeof(n), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaMemcpyAsync(&fl, t->flags, sizeof(fl), cudaMemcpyDeviceToHost, s));
    if (gd) SYK_CUDA(cudaFreeAsync(gd, s));
    SYK_CUDA(cudaStreamSynchronize(s));
    if (n_out) *n_out = n;
    if (fl) {
        syk_set_error("id table overflow (capacity %llu): retry with a larger capacity", (unsigned long long)t->capacity);
        return SYK_EOVERFLOW;
    }
    if (n > max_records) {
        syk_set_error("export buffer too small: %llu records, room for %llu", n, (unsigned long long)max_records);
        return SYK_EOVERFLOW;
    }
    return SYK_OK;
}

// append mode has no host round trip: a table that overflowed poisons the log counter (bit 62) so that the host
// sees an impossible record count when it finally reads it
__global__ void k_poison_on_overflow(const int *flags, unsigned long long *counter) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && flags[0]) atomicOr(counter, 1ull << 62);
}

static int append_records(syk_table_t *t, const syk_chunk_geom_t *geom_host, syk_record_t *log_dev, uint64_t max_records,
                          uint64_t *counter_dev, uint64_t min_vx, void *stream);
SYK_API int syk_table_append_records(syk_table_t *t, const syk_chunk_geom_t *geom_host, syk_record_t *log_dev,
                                     uint64_t max_records, uint64_t *counter_dev, void *stream) {
    return append_records(t, geom_host, log_dev, max_records, counter_dev, 0, stream);
}
SYK_API int syk_table_append_records_min_vx(syk_table_t *t, const syk_chunk_geom_t *geom_host, syk_record_t *log_dev,
                                            uint64_t max_records, uint64_t *counter_dev, uint64_t min_vx, void *stream) {
    SYK_CHECK_ARG(geom_host != nullptr || min_vx <= 1, "the small-object drop needs the chunk geometry");
    return append_records(t, geom_host, log_dev, max_records, counter_dev, min_vx, stream);
}
static int append_records(syk_table_t *t, const syk_chunk_geom_t *geom_host, syk_record_t *log_dev, uint64_t max_records,
                          uint64_t *counter_dev, uint64_t min_vx, void *stream) {
    SYK_CHECK_ARG(t != nullptr && log_dev != nullptr && counter_dev != nullptr, "NULL argument");
    cudaStream_t s = (cudaStream_t)stream;
    syk_chunk_geom_t *gd = nullptr;
    int rc = upload_geoms(geom_host, geom_host ? 1 : 0, s, &gd);
    if (rc) return rc;
    int blocks = (int)((t->capacity + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_table_export<<<blocks, 256, 0, s>>>(t->slots, t->capacity, log_dev, max_records, (unsigned long long *)counter_dev, gd,
                                          gd ? 0xFFFFFFFFu : 0u, (unsigned long long)min_vx);
    k_poison_on_overflow<<<1, 32, 0, s>>>(t->flags, (unsigned long long *)counter_dev);
    SYK_CUDA(cudaGetLastError());
    if (gd) SYK_CUDA(cudaFreeAsync(gd, s));
    return SYK_OK;
}

__global__ void k_decode_rep(syk_record_t *recs, uint64_t n, const syk_chunk_geom_t *geoms, uint32_t n_geoms) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    syk_record_t r = recs[i];
    decode_rep(r, geoms, n_geoms);
    recs[i] = r;
}

SYK_API int syk_records_decode_rep(syk_record_t *records_dev, uint64_t n, const syk_chunk_geom_t *geoms_host, uint32_t n_geoms,
                                   void *stream) {
    if (n == 0) return SYK_OK;
    cudaStream_t s = (cudaStream_t)stream;
    syk_chunk_geom_t *gd = nullptr;
    int rc = upload_geoms(geoms_host, n_geoms, s, &gd);
    if (rc) return rc;
    k_decode_rep<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(records_dev, n, gd, n_geoms);
    SYK_CUDA(cudaGetLastError());
    if (gd) SYK_CUDA(cudaFreeAsync(gd, s));
    return SYK_OK;
}

__global__ void k_merge_records(TableView t, const syk_record_t *recs, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const syk_record_t r = recs[i];
    if (r.id == 0ull) return;
    syk_table_update(t, r.id, r.count, r.rep_key, r.bb_min[0], r.bb_min[1], r.bb_min[2], r.bb_max[0], r.bb_max[1],
                     r.bb_max[2]);
}

SYK_API int syk_table_merge_records(syk_table_t *t, const syk_record_t *records_dev, uint64_t n, void *stream) {
    SYK_CHECK_ARG(t != nullptr, "table is NULL");
    if (n == 0) return SYK_OK;
    k_merge_records<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(view_of(t), records_dev, n);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

// ---- bucketing by owner (for the hash-owner all-to-all) ------------------------------------------------------
template <typename R>
__device__ __forceinline__ uint64_t owner_key(const R &r);
template <>
__device__ __forceinline__ uint64_t owner_key<syk_record_t>(const syk_record_t &r) { return r.id; }
template <>
__device__ __forceinline__ uint64_t owner_key<syk_pair_t>(const syk_pair_t &r) { return r.sub_id; }

__host__ __device__ __forceinline__ uint32_t syk_owner_of(uint64_t id, uint32_t n_owners) {
    return (uint32_t)((syk_mix64(id ^ 0x5bd1e9955bd1e995ULL) >> 20) % n_owners);
}

// Block-level staging: the number of owners is tiny (<= #GPUs), so per-row global atomics would all hit the same few
// addresses.  Each block counts into shared memory and issues one global atomic per owner.
constexpr int BUCKET_MAX_SMEM_OWNERS = 64;

template <typename R>
__global__ void k_bucket_count(const R *recs, uint64_t n, uint32_t n_owners, unsigned long long *counts) {
    __shared__ unsigned int local[BUCKET_MAX_SMEM_OWNERS];
    const bool use_smem = n_owners <= BUCKET_MAX_SMEM_OWNERS;
    if (use_smem) {
        for (unsigned i = threadIdx.x; i < n_owners; i += blockDim.x) local[i] = 0u;
        __syncthreads();
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t o = syk_owner_of(owner_key(recs[i]), n_owners);
        if (use_smem) atomicAdd(&local[o], 1u);
        else atomicAdd(&counts[o], 1ull);
    }
    if (use_smem) {
        __syncthreads();
        for (unsigned i = threadIdx.x; i < n_owners; i += blockDim.x)
            if (local[i]) atomicAdd(&counts[i], (unsigned long long)local[i]);
    }
}
// counts[0..n) -> cursors[0..n) = exclusive prefix
__global__ void k_bucket_scan(const unsigned long long *counts, unsigned long long *cursors, uint32_t n_owners) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t o = 0; o < n_owners; ++o) {
            cursors[o] = acc;
            acc += counts[o];
        }
    }
}
// every block handles one contiguous slice: count per owner in shared memory, reserve the block's ranges with one
// global atomic per owner, then scatter (order inside a bucket is unspecified)
template <typename R>
__global__ void k_bucket_scatter(const R *recs, uint64_t n, uint32_t n_owners, unsigned long long *cursors, R *out) {
    __shared__ unsigned int local[BUCKET_MAX_SMEM_OWNERS];
    __shared__ unsigned long long base[BUCKET_MAX_SMEM_OWNERS];
    const bool use_smem = n_owners <= BUCKET_MAX_SMEM_OWNERS;
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = (uint64_t)blockIdx.x * per_block;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    if (!use_smem) {
        for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const R r = recs[i];
            out[atomicAdd(&cursors[syk_owner_of(owner_key(r), n_owners)], 1ull)] = r;
        }
        return;
    }
    for (unsigned i = threadIdx.x; i < n_owners; i += blockDim.x) local[i] = 0u;
    __syncthreads();
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&local[syk_owner_of(owner_key(recs[i]), n_owners)], 1u);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < n_owners; i += blockDim.x) {
        base[i] = local[i] ? atomicAdd(&cursors[i], (unsigned long long)local[i]) : 0ull;
        local[i] = 0u;
    }
    __syncthreads();
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const R r = recs[i];
        const uint32_t o = syk_owner_of(owner_key(r), n_owners);
        out[base[o] + atomicAdd(&local[o], 1u)] = r;
    }
}

template <typename R>
static int bucket_impl(const R *recs, uint64_t n, uint32_t n_owners, R *out, uint64_t *counts_dev, cudaStream_t s) {
    SYK_CHECK_ARG(n_owners >= 1 && n_owners <= 4096, "n_owners out of range");
    SYK_CHECK_ARG(counts_dev != nullptr, "counts_dev is NULL");
    SYK_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * n_owners, s));
    if (n == 0) return SYK_OK;
    unsigned long long *cursors = nullptr;
    SYK_CUDA(cudaMallocAsync((void **)&cursors, sizeof(unsigned long long) * n_owners, s));
    unsigned blocks = (unsigned)((n + 1023) / 1024);
    if (blocks > 148 * 4) blocks = 148 * 4;
    k_bucket_count<R><<<blocks, 256, 0, s>>>(recs, n, n_owners, (unsigned long long *)counts_dev);
    k_bucket_scan<<<1, 32, 0, s>>>((unsigned long long *)counts_dev, cursors, n_owners);
    k_bucket_scatter<R><<<blocks, 256, 0, s>>>(recs, n, n_owners, cursors, out);
    SYK_CUDA(cudaGetLastError());
    SYK_CUDA(cudaFreeAsync(cursors, s));
    return SYK_OK;
}

SYK_API int syk_records_bucket(const syk_record_t *records_dev, uint64_t n, uint32_t n_owners, syk_record_t *out_dev,
                               uint64_t *counts_dev, void *stream) {
    return bucket_impl<syk_record_t>(records_dev, n, n_owners, out_dev, counts_dev, (cudaStream_t)stream);
}
SYK_API int syk_pairs_bucket(const syk_pair_t *pairs_dev, uint64_t n, uint32_t n_owners, syk_pair_t *out_dev,
                             uint64_t *counts_dev, void *stream) {
    return bucket_impl<syk_pair_t>(pairs_dev, n, n_owners, out_dev, counts_dev, (cudaStream_t)stream);
}

// ---- dense relabelling of 64-bit ids (support of the 64-bit contact-site variants) -------------------------------------
__global__ void k_dense_insert(const void *__restrict__ vol, int elem_bytes, unsigned long long n, TableView t) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = elem_bytes == 8 ? ((const unsigned long long *)vol)[i] : (unsigned long long)((const unsigned *)vol)[i];
        if (k != 0ull) syk_table_slot(t, k);
    }
}
__global__ void k_dense_number(SykSlot *slots, uint64_t cap, unsigned long long *counter, unsigned long long *ids, unsigned long long max_ids) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = slots[i].key;
        if (k == 0ull) continue;
        const unsigned long long idx = atomicAdd(counter, 1ull);
        slots[i].count = idx + 1ull;  // dense label
        if (idx < max_ids) ids[idx] = k;
    }
}
__global__ void k_dense_apply(const void *__restrict__ vol, int elem_bytes, unsigned long long n, TableView t, unsigned *__restrict__ out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = elem_bytes == 8 ? ((const unsigned long long *)vol)[i] : (unsigned long long)((const unsigned *)vol)[i];
        unsigned lab = 0u;
        if (k != 0ull) {
            const SykSlot *s = syk_table_slot(t, k);
            lab = s ? (unsigned)s->count : 0u;
        }
        out[i] = lab;
    }
}
// labels_out[i] = dense label (1..n_ids, 0 for id 0) of vol[i]; ids_out[label - 1] = id.  `t` must be empty; it is left
// holding id -> label in the count field.  SYK_EOVERFLOW: table or ids_out too small.
SYK_API int syk_dense_relabel(syk_table_t *t, const void *vol_dev, int elem_bytes, uint64_t n, uint32_t *labels_out_dev,
                              uint64_t *ids_out_dev, uint64_t max_ids, uint64_t *n_ids_out_host, void *stream) {
    SYK_CHECK_ARG(t && vol_dev && labels_out_dev && ids_out_dev && n_ids_out_host, "NULL argument");
    SYK_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    cudaStream_t s = (cudaStream_t)stream;
    *n_ids_out_host = 0;
    if (n == 0) return SYK_OK;
    unsigned long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_dense_insert<<<(unsigned)blocks, 256, 0, s>>>(vol_dev, elem_bytes, n, view_of(t));
    SYK_CUDA(cudaMemsetAsync(t->counter, 0, sizeof(unsigned long long), s));
    unsigned long long cb = (t->capacity + 255) / 256;
    if (cb > 148 * 16) cb = 148 * 16;
    k_dense_number<<<(unsigned)cb, 256, 0, s>>>(t->slots, t->capacity, t->counter, (unsigned long long *)ids_out_dev, max_ids);
    k_dense_apply<<<(unsigned)blocks, 256, 0, s>>>(vol_dev, elem_bytes, n, view_of(t), labels_out_dev);
    SYK_CUDA(cudaGetLastError());
    unsigned long long nid = 0;
    int fl = 0;
    SYK_CUDA(cudaMemcpyAsync(&nid, t->counter, sizeof(nid), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaMemcpyAsync(&fl, t->flags, sizeof(fl), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaStreamSynchronize(s));
    *n_ids_out_host = nid;
    if (fl || nid > max_ids || nid >= 0xFFFFFFFFull) {
        syk_set_error("dense relabel: table / id buffer too small (%llu ids)", nid);
        return SYK_EOVERFLOW;
    }
    return SYK_OK;
}

// ---- pair tables ------------------------------------------------------------------------------------------------
int syk_pairs_create_on(syk_pairs **out, uint64_t capacity, cudaStream_t s) {
    SYK_CHECK_ARG(out != nullptr, "out is NULL");
    int rc = syk_require_device();
    if (rc) return rc;
    syk_pairs *t = (syk_pairs *)calloc(1, sizeof(syk_pairs));
    if (!t) return SYK_ENOMEM;
    t->capacity = round_pow2(capacity);
    t->stream = s;
    SYK_CUDA(cudaGetDevice(&t->device));
    syk_pool_keep_warm();
    void *ctl = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&t->slots, t->capacity * sizeof(SykPairSlot), s);
    if (e == cudaSuccess) e = cudaMallocAsync(&ctl, 64, s);
    if (e != cudaSuccess) {
        syk_set_error("cudaMalloc of %llu pair slots failed: %s", (unsigned long long)t->capacity, cudaGetErrorString(e));
        if (t->slots) cudaFreeAsync(t->slots, s);
        free(t);
        cudaGetLastError();
        return SYK_ENOMEM;
    }
    t->flags = (int *)ctl;
    t->counter = (unsigned long long *)((char *)ctl + 32);
    SYK_CUDA(cudaMemsetAsync(t->slots, 0, t->capacity * sizeof(SykPairSlot), s));
    SYK_CUDA(cudaMemsetAsync(ctl, 0, 64, s));
    SYK_CUDA(cudaStreamSynchronize(s));  // see syk_table_create_on
    *out = t;
    return SYK_OK;
}
SYK_API int syk_pairs_create(syk_pairs_t **out, uint64_t capacity) { return syk_pairs_create_on(out, capacity, (cudaStream_t)0); }
SYK_API uint64_t syk_pairs_capacity(const syk_pairs_t *t) { return t ? t->capacity : 0; }
SYK_API int syk_pairs_destroy(syk_pairs_t *t) {
    if (!t) return SYK_OK;
    cudaDeviceSynchronize();
    cudaGetLastError();
    cudaFreeAsync(t->slots, t->stream);
    cudaFreeAsync(t->flags, t->stream);
    free(t);
    return SYK_OK;
}
SYK_API int syk_pairs_clear(syk_pairs_t *t, void *stream) {
    SYK_CHECK_ARG(t != nullptr, "pair table is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    SYK_CUDA(cudaMemsetAsync(t->slots, 0, t->capacity * sizeof(SykPairSlot), s));
    SYK_CUDA(cudaMemsetAsync(t->flags, 0, 4 * sizeof(int), s));
    return SYK_OK;
}

// read-only lookup (no insertion)
__device__ __forceinline__ const SykSlot *table_find(const TableView &t, unsigned long long key) {
    if (t.slots == nullptr) return nullptr;
    uint64_t h = syk_mix64(key) & t.mask;
    for (uint64_t probes = 0; probes <= t.mask; ++probes) {
        const SykSlot *s = t.slots + h;
        const unsigned long long cur = s->key;
        if (cur == key) return s;
        if (cur == 0ull) return nullptr;
        h = (h + 1) & t.mask;
    }
    return nullptr;
}

// sub_t / geom / min_vx: drop the pairs of organelle objects that the small-object drop removed (sd_proc.py:678-679)
__global__ void k_pairs_export(const SykPairSlot *slots, uint64_t cap, syk_pair_t *out, unsigned long long max_out,
                               unsigned long long *counter, TableView sub_t = TableView{nullptr, 0, nullptr},
                               const syk_chunk_geom_t *geom = nullptr, unsigned long long min_vx = 0ull) {
    uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    for (uint64_t base = i0 - lane; base < cap; base += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i = base + lane;
        SykPairSlot s;
        bool occ = false;
        if (i < cap) {
            s = slots[i];
            occ = s.sub != 0ull;
        }
        if (occ && min_vx > 1ull && geom != nullptr) {
            const SykSlot *o = table_find(sub_t, s.sub);
            if (o != nullptr) {
                syk_record_t r;
                r.count = o->count;
                for (int a = 0; a < 3; ++a) {
                    r.bb_min[a] = (int32_t)((0xFFFFFFFFu - o->min_enc[a]) - SYK_COORD_BIAS);
                    r.bb_max[a] = (int32_t)(o->max_enc[a] - SYK_COORD_BIAS);
                }
                if (small_inside(r, geom[0], min_vx)) occ = false;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, occ);
        if (!m) continue;
        unsigned long long pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(counter, (unsigned long long)__popc(m));
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        if (occ) {
            unsigned long long pos = pos0 + __popc(m & ((1u << lane) - 1u));
            if (pos < max_out) {
                syk_pair_t r;
                r.sub_id = s.sub;
                r.cell_id = s.cell;
                r.count = s.count;
                r._pad = 0;
                out[pos] = r;
            }
        }
    }
}

SYK_API int syk_pairs_export(syk_pairs_t *t, syk_pair_t *pairs_dev, uint64_t max_pairs, uint64_t *n_out, void *stream) {
    SYK_CHECK_ARG(t != nullptr, "pair table is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    SYK_CUDA(cudaMemsetAsync(t->counter, 0, sizeof(unsigned long long), s));
    int blocks = (int)((t->capacity + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pairs_export<<<blocks, 256, 0, s>>>(t->slots, t->capacity, pairs_dev, max_pairs, t->counter);
    SYK_CUDA(cudaGetLastError());
    unsigned long long n = 0;
    int fl = 0;
    SYK_CUDA(cudaMemcpyAsync(&n, t->counter, sizeof(n), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaMemcpyAsync(&fl, t->flags, sizeof(fl), cudaMemcpyDeviceToHost, s));
    SYK_CUDA(cudaStreamSynchronize(s));
    if (n_out) *n_out = n;
    if (fl) {
        syk_set_error("pair table overflow (capacity %llu): retry with a larger capacity", (unsigned long long)t->capacity);
        return SYK_EOVERFLOW;
    }
    if (n > max_pairs) {
        syk_set_error("pair export buffer too small: %llu pairs, room for %llu", n, (unsigned long long)max_pairs);
        return SYK_EOVERFLOW;
    }
    return SYK_OK;
}

SYK_API int syk_pairs_append(syk_pairs_t *t, syk_pair_t *log_dev, uint64_t max_pairs, uint64_t *counter_dev, void *stream) {
    SYK_CHECK_ARG(t != nullptr && log_dev != nullptr && counter_dev != nullptr, "NULL argument");
    int blocks = (int)((t->capacity + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pairs_export<<<blocks, 256, 0, (cudaStream_t)stream>>>(t->slots, t->capacity, log_dev, max_pairs,
                                                             (unsigned long long *)counter_dev);
    k_poison_on_overflow<<<1, 32, 0, (cudaStream_t)stream>>>(t->flags, (unsigned long long *)counter_dev);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

// syk_pairs_append with the worker's small-object drop: pairs whose organelle object (looked up in the chunk's organelle table
// `sub_t`) lies inside the chunk `geom_host` with fewer than min_vx voxels are skipped (sd_proc.py:667-680)
SYK_API int syk_pairs_append_min_vx(syk_pairs_t *t, syk_table_t *sub_t, const syk_chunk_geom_t *geom_host, uint64_t min_vx,
                                    syk_pair_t *log_dev, uint64_t max_pairs, uint64_t *counter_dev, void *stream) {
    SYK_CHECK_ARG(t != nullptr && log_dev != nullptr && counter_dev != nullptr, "NULL argument");
    SYK_CHECK_ARG(min_vx <= 1 || (sub_t != nullptr && geom_host != nullptr), "the small-object drop needs the organelle table and the chunk geometry");
    cudaStream_t s = (cudaStream_t)stream;
    syk_chunk_geom_t *gd = nullptr;
    int rc = upload_geoms(geom_host, geom_host ? 1 : 0, s, &gd);
    if (rc) return rc;
    int blocks = (int)((t->capacity + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pairs_export<<<blocks, 256, 0, s>>>(t->slots, t->capacity, log_dev, max_pairs, (unsigned long long *)counter_dev, view_of(sub_t), gd,
                                          (unsigned long long)min_vx);
    k_poison_on_overflow<<<1, 32, 0, s>>>(t->flags, (unsigned long long *)counter_dev);
    SYK_CUDA(cudaGetLastError());
    if (gd) SYK_CUDA(cudaFreeAsync(gd, s));
    return SYK_OK;
}

// Second, smaller exchange of the reduce (sd_proc.py:1054-1084): every overlap pair learns the total size of its organelle
// object from the owner's final organelle table (written to the pair's `_pad` field; 0 = organelle not in the table, e.g.
// removed by the size threshold), so that after re-bucketing by cell id the receiver can form count / size ratios.
__global__ void k_pairs_attach_size(syk_pair_t *pairs, uint64_t n, TableView sub_t) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SykSlot *o = table_find(sub_t, pairs[i].sub_id);
    pairs[i]._pad = o ? o->count : 0ull;
}
SYK_API int syk_pairs_attach_size(syk_pair_t *pairs_dev, uint64_t n, syk_table_t *sub_final, void *stream) {
    SYK_CHECK_ARG(sub_final != nullptr && (pairs_dev != nullptr || n == 0), "NULL argument");
    if (n == 0) return SYK_OK;
    k_pairs_attach_size<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pairs_dev, n, view_of(sub_final));
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}

__global__ void k_pairs_merge(PairView t, const syk_pair_t *p, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const syk_pair_t r = p[i];
    if (r.sub_id == 0ull || r.cell_id == 0ull) return;
    syk_pairs_update(t, r.sub_id, r.cell_id, r.count);
}

SYK_API int syk_pairs_merge(syk_pairs_t *t, const syk_pair_t *pairs_dev, uint64_t n, void *stream) {
    SYK_CHECK_ARG(t != nullptr, "pair table is NULL");
    if (n == 0) return SYK_OK;
    k_pairs_merge<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(view_of(t), pairs_dev, n);
    SYK_CUDA(cudaGetLastError());
    return SYK_OK;
}


// ---- records sorted by id (the contact-site worker walks its objects in ascending id order) ----------------
__global__ void k_record_keys(const syk_record_t *__restrict__ recs, uint64_t n, unsigned long long *__restrict__ keys, unsigned *__restrict__ idx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = recs[i].id;
        idx[i] = (unsigned)i;
    }
}
__global__ void k_record_gather(const syk_record_t *__restrict__ src, const unsigned *__restrict__ idx, uint64_t n, syk_record_t *__restrict__ dst) {
    // 64-byte records as four 16-byte pieces: four consecutive lanes move one record
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t i = t >> 2;
    if (i < n) reinterpret_cast<uint4 *>(dst + i)[t & 3] = reinterpret_cast<const uint4 *>(src + idx[i])[t & 3];
}

SYK_API int syk_records_sort_by_id(syk_record_t *records_dev, uint64_t n, void *stream) {
    int rc = syk_require_device();
    if (rc) return rc;
    if (n < 2) return SYK_OK;
    SYK_CHECK_ARG(records_dev, "NULL records");
    SYK_CHECK_ARG(n < (1ull << 31), "too many records");
    cudaStream_t s = (cudaStream_t)stream;
    struct Scratch {
        cudaStream_t s;
        void *p[3] = {nullptr, nullptr, nullptr};
        ~Scratch() {
            for (void *x : p)
                if (x) cudaFreeAsync(x, s);
        }
    } sc{s};
    syk_pool_keep_warm();
    // [keys | sorted keys | idx | sorted idx], the record copy, cub's temporary storage
    SYK_CUDA(cudaMallocAsync(&sc.p[0], n * (2 * sizeof(unsigned long long) + 2 * sizeof(unsigned)), s));
    SYK_CUDA(cudaMallocAsync(&sc.p[1], n * sizeof(syk_record_t), s));
    unsigned long long *keys = (unsigned long long *)sc.p[0], *keys2 = keys + n;
    unsigned *idx = (unsigned *)(keys2 + n), *idx2 = idx + n;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_record_keys<<<blocks, 256, 0, s>>>(records_dev, n, keys, idx);
    size_t tmp_bytes = 0;
    SYK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int)n, 0, 64, s));
    SYK_CUDA(cudaMallocAsync(&sc.p[2], tmp_bytes ? tmp_bytes : 16, s));
    SYK_CUDA(cub::DeviceRadixSort::SortPairs(sc.p[2], tmp_bytes, keys, keys2, idx, idx2, (int)n, 0, 64, s));
    k_record_gather<<<(unsigned)((4 * n + 255) / 256), 256, 0, s>>>(records_dev, idx2, n, (syk_record_t *)sc.p[1]);
    SYK_CUDA(cudaGetLastError());
    SYK_CUDA(cudaMemcpyAsync(records_dev, sc.p[1], n * sizeof(syk_record_t), cudaMemcpyDeviceToDevice, s));
    return SYK_OK;
}
