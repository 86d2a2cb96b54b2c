// syk_common.cuh -- shared device/host helpers of libsyk (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/syk.h"

#define SYK_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing -----------------------------------------------------------------------------------------
void syk_set_error(const char *fmt, ...);
int syk_require_device();
void syk_pool_keep_warm();  // raise the release threshold of the default stream-ordered pool (once)
// Stream of the calling host thread for the *_host entry points (non-blocking, created on first use): calls made
// from different threads overlap their PCIe copies and kernels instead of serialising on the default stream.
cudaStream_t syk_host_stream();
// Completion wait of the *_host entry points: cudaStreamSynchronize, or (SYK_YIELD_WAIT=1) a cudaStreamQuery poll that yields
// the core, for hosts where more threads wait than there are cores.
cudaError_t syk_stream_wait(cudaStream_t s);
// Raise (never lower) a kernel's opt-in dynamic shared-memory limit.  The *_host API is thread-concurrent: setting the
// attribute to the size of the current call would let two threads with different stencils undo each other between
// set and launch; a monotone, mutex-guarded maximum cannot.
int syk_ensure_dyn_smem(const void *func, int bytes);
// Rank-3 TMA tensor map over a label volume given by its internal axes (u, v, w; w contiguous): dims {n[2], n[1], n[0]},
// box {box_w, box_v, 1}, no swizzle, out-of-bounds elements read as zero.  cuTensorMapEncodeTiled is resolved through the
// runtime (no link against libcuda).  false: the view cannot be described (unaligned base / strides) or TMA is disabled
// (SYK_NO_TMA) -- callers then use their LDG / cp.async path.
bool syk_make_tmap3(CUtensorMap *m, const void *base, int elem_bytes, const long long n[3], const long long st[3], int box_w,
                    int box_v);

#define SYK_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            syk_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));    \
            return SYK_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

#define SYK_CHECK_ARG(cond, msg)                   \
    do {                                           \
        if (!(cond)) {                             \
            syk_set_error("invalid argument: %s", msg); \
            return SYK_EINVAL;                     \
        }                                          \
    } while (0)

// ---- device hash tables -------------------------------------------------------------------------------------
// One 64-byte slot per id.  All value fields are encoded so that an all-zero slot is the neutral element and
// every update is a fire-and-forget atomicAdd / atomicMax (RED at the L2):
//   count   : atomicAdd
//   rep_enc : atomicMax of rep_key + 1           (0 = none)
//   min_enc : atomicMax of 0xFFFFFFFF - (c+BIAS)  (0 = none)   -> min coordinate
//   max_enc : atomicMax of (c_exclusive + BIAS)   (0 = none)   -> max coordinate (exclusive)
struct __align__(16) SykSlot {
    unsigned long long key;
    unsigned long long count;
    unsigned long long rep_enc;
    uint32_t min_enc[3];
    uint32_t max_enc[3];
    uint32_t pad[4];
};
static_assert(sizeof(SykSlot) == 64, "slot must be 64 bytes");

struct __align__(16) SykPairSlot {  // key = (sub, cell) claimed with one 128-bit CAS
    unsigned long long sub;
    unsigned long long cell;
    unsigned long long count;
    unsigned long long pad;
};
static_assert(sizeof(SykPairSlot) == 32, "pair slot must be 32 bytes");

#define SYK_COORD_BIAS 0x40000000u
#define SYK_REP_MASK ((1ull << 40) - 1ull)

struct syk_table {
    SykSlot *slots;
    uint64_t capacity;  // power of two
    int *flags;         // device: [0] overflow (first 16 bytes of the 64-byte control block)
    unsigned long long *counter;  // device scratch counter (export), control block + 32
    int device;
    cudaStream_t stream;  // allocation stream (the table is freed on it)
};
struct syk_pairs {
    SykPairSlot *slots;
    uint64_t capacity;
    int *flags;
    unsigned long long *counter;
    int device;
    cudaStream_t stream;
};
// stream-ordered construction (the public creators use the default stream)
int syk_table_create_on(syk_table **out, uint64_t capacity, cudaStream_t s);
int syk_pairs_create_on(syk_pairs **out, uint64_t capacity, cudaStream_t s);

// device-side views passed by value to kernels
struct TableView {
    SykSlot *slots;
    uint64_t mask;
    int *flags;
};
struct PairView {
    SykPairSlot *slots;
    uint64_t mask;
    int *flags;
};
static inline TableView view_of(const syk_table *t) {
    TableView v;
    v.slots = t ? t->slots : nullptr;
    v.mask = t ? t->capacity - 1 : 0;
    v.flags = t ? t->flags : nullptr;
    return v;
}
static inline PairView view_of(const syk_pairs *t) {
    PairView v;
    v.slots = t ? t->slots : nullptr;
    v.mask = t ? t->capacity - 1 : 0;
    v.flags = t ? t->flags : nullptr;
    return v;
}

__host__ __device__ __forceinline__ uint64_t syk_mix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
__host__ __device__ __forceinline__ uint32_t syk_mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t syk_hash_id32(uint64_t id) {
    return syk_mix32((uint32_t)id ^ ((uint32_t)(id >> 32) * 0x9E3779B1u));
}

#ifdef __CUDACC__
// Find-or-claim the slot of `key` (key != 0).  Returns nullptr (and raises the overflow flag) when full.
__device__ __forceinline__ SykSlot *syk_table_slot(const TableView &t, unsigned long long key) {
    uint64_t h = syk_mix64(key) & t.mask;
    for (uint64_t probes = 0; probes <= t.mask; ++probes) {
        SykSlot *s = t.slots + h;
        unsigned long long cur = *(volatile unsigned long long *)&s->key;
        if (cur == key) return s;
        if (cur == 0ull) {
            unsigned long long prev = atomicCAS(&s->key, 0ull, key);
            if (prev == 0ull || prev == key) return s;
        }
        h = (h + 1) & t.mask;
    }
    atomicExch(t.flags, 1);
    return nullptr;
}

// bb_min / bb_max_excl are signed global coordinates (|c| < 2^30)
__device__ __forceinline__ void syk_table_update(const TableView &t, unsigned long long key, unsigned long long count,
                                                 unsigned long long rep_key, int mnx, int mny, int mnz, int mxx, int mxy,
                                                 int mxz) {
    SykSlot *s = syk_table_slot(t, key);
    if (!s) return;
    atomicAdd(&s->count, count);
    atomicMax(&s->rep_enc, rep_key + 1ull);
    atomicMax(&s->min_enc[0], 0xFFFFFFFFu - ((uint32_t)mnx + SYK_COORD_BIAS));
    atomicMax(&s->min_enc[1], 0xFFFFFFFFu - ((uint32_t)mny + SYK_COORD_BIAS));
    atomicMax(&s->min_enc[2], 0xFFFFFFFFu - ((uint32_t)mnz + SYK_COORD_BIAS));
    atomicMax(&s->max_enc[0], (uint32_t)mxx + SYK_COORD_BIAS);
    atomicMax(&s->max_enc[1], (uint32_t)mxy + SYK_COORD_BIAS);
    atomicMax(&s->max_enc[2], (uint32_t)mxz + SYK_COORD_BIAS);
}

__device__ __forceinline__ void syk_pairs_update(const PairView &t, unsigned long long sub, unsigned long long cell,
                                                 unsigned long long count) {
    uint64_t h = syk_mix64(sub * 0x9E3779B97F4A7C15ULL ^ syk_mix64(cell)) & t.mask;
    const unsigned __int128 want = ((unsigned __int128)cell << 64) | (unsigned __int128)sub;  // little endian: sub first
    for (uint64_t probes = 0; probes <= t.mask; ++probes) {
        SykPairSlot *s = t.slots + h;
        unsigned long long cs, cc;
        do {  // one 16-byte load of the key; a half-written key cannot exist (claimed by a single 128-bit CAS),
              // the retry only guards against a split load
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(cs), "=l"(cc) : "l"(s));
        } while ((cs == 0ull) != (cc == 0ull));
        bool mine = (cs == sub && cc == cell);
        if (!mine && cs == 0ull && cc == 0ull) {
            unsigned __int128 prev = atomicCAS((unsigned __int128 *)s, (unsigned __int128)0, want);
            mine = (prev == 0) || (prev == want);
        }
        if (mine) {
            atomicAdd(&s->count, count);
            return;
        }
        h = (h + 1) & t.mask;
    }
    atomicExch(t.flags, 1);
}
#endif
