/*
 * oracle/syk_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the SyConn label-volume extraction hot path.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may load this library; the product (syconn_b200) never does.
 *
 * Each function cites the reference file:line (relative to /root/reference) whose
 * behaviour it restates.  The restatement is pinned against the reference's own
 * compiled Cython/numba code by tests/golden/make_golden.py (fixtures committed)
 * and against the reference's known-answer tests (tests/test_oracle_pins.py).
 *
 * All volumes are addressed as arr[x*sx + y*sy + z*sz] with strides given in
 * ELEMENTS, so any NumPy view (e.g. ZYX memory seen as XYZ) can be passed as is;
 * results are always in the LOGICAL (x, y, z) index order of the view, exactly as
 * the reference's typed memoryviews behave.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline uint64_t ld_label(const void *base, int elem_bytes, int64_t idx) {
    return elem_bytes == 8 ? ((const uint64_t *)base)[idx] : (uint64_t)((const uint32_t *)base)[idx];
}

static inline uint64_t mix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

/* ------------------------------------------------------------------------------------------
 * detect_seg_boundaries  -- syconn/extraction/find_object_properties.py:424-455
 * out[x,y,z] = 1 iff arr != 0 and some IN-BOUNDS face neighbour differs (0 counts as different).
 * out is C-contiguous uint8 [X,Y,Z].
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_detect_seg_boundaries(const void *arr, int elem_bytes, const int64_t shape[3],
                                      const int64_t st[3], uint8_t *out) {
    const int64_t nx = shape[0], ny = shape[1], nz = shape[2];
    for (int64_t x = 0; x < nx; ++x)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t z = 0; z < nz; ++z) {
                const int64_t i = x * st[0] + y * st[1] + z * st[2];
                const uint64_t c = ld_label(arr, elem_bytes, i);
                uint8_t b = 0;
                if (c != 0) {
                    if (x > 0 && ld_label(arr, elem_bytes, i - st[0]) != c) b = 1;
                    if (x + 1 < nx && ld_label(arr, elem_bytes, i + st[0]) != c) b = 1;
                    if (y > 0 && ld_label(arr, elem_bytes, i - st[1]) != c) b = 1;
                    if (y + 1 < ny && ld_label(arr, elem_bytes, i + st[1]) != c) b = 1;
                    if (z > 0 && ld_label(arr, elem_bytes, i - st[2]) != c) b = 1;
                    if (z + 1 < nz && ld_label(arr, elem_bytes, i + st[2]) != c) b = 1;
                }
                out[(x * ny + y) * nz + z] = b;
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * kernel + process_block_nonzero -- syconn/extraction/block_processing_C.pyx:21-49, :53-75
 *
 * "valid"-mode stencil.  For every output voxel whose centre is flagged in `edges`: histogram of
 * the sx*sy*sz window of uint32 IDs, counts of ID 0 and of the centre ID forced to 0, arg-max with
 * strict '>' while iterating keys ASCENDING (std::map) => ties go to the SMALLEST id; result
 * (min(center,key) << 32) + max(center,key) if max count > 0, else 0.
 * `arr` elements are read as uint32 (elem_bytes 4) or as uint64 truncated to uint32 (elem_bytes 8,
 * = the caller-side .astype(np.uint32) of extraction/cs_extraction_steps.py:385-387).
 * out: C-contiguous uint64 [X-sx+1, Y-sy+1, Z-sz+1].
 * ------------------------------------------------------------------------------------------ */
#define WTAB 16384 /* > 2 * max supported window (e.g. 21*21*9 = 3969) */
typedef struct { uint32_t key; int32_t cnt; uint32_t stamp; } wslot_t;

ORC_API int orc_process_block_nonzero(const void *edges, int edge_bytes, const int64_t est[3],
                                      const void *arr, int elem_bytes, const int64_t ast[3],
                                      const int64_t shape[3], const int32_t stencil[3], uint64_t *out) {
    const int sx = stencil[0], sy = stencil[1], sz = stencil[2];
    if ((sx % 2 + sy % 2 + sz % 2) != 3) return -2; /* block_processing_C.pyx:57 assert */
    if ((int64_t)sx * sy * sz * 2 > WTAB) return -3;
    const int64_t X = shape[0] - sx + 1, Y = shape[1] - sy + 1, Z = shape[2] - sz + 1;
    if (X <= 0 || Y <= 0 || Z <= 0) return 0;
    const int ox = sx / 2, oy = sy / 2, oz = sz / 2;
    wslot_t *tab = (wslot_t *)calloc(WTAB, sizeof(wslot_t));
    uint32_t *touched = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)sx * sy * sz);
    uint32_t stamp = 0;
    for (int64_t x = 0; x < X; ++x)
        for (int64_t y = 0; y < Y; ++y)
            for (int64_t z = 0; z < Z; ++z) {
                uint64_t res = 0;
                const int64_t ce = (x + ox) * est[0] + (y + oy) * est[1] + (z + oz) * est[2];
                const uint64_t e = edge_bytes == 1 ? ((const uint8_t *)edges)[ce] : ((const uint32_t *)edges)[ce];
                if (e != 0) {
                    const uint32_t center =
                        (uint32_t)ld_label(arr, elem_bytes, (x + ox) * ast[0] + (y + oy) * ast[1] + (z + oz) * ast[2]);
                    int nt = 0;
                    ++stamp;
                    for (int i = 0; i < sx; ++i)
                        for (int j = 0; j < sy; ++j) {
                            const int64_t b = (x + i) * ast[0] + (y + j) * ast[1] + z * ast[2];
                            for (int k = 0; k < sz; ++k) {
                                const uint32_t id = (uint32_t)ld_label(arr, elem_bytes, b + k * ast[2]);
                                uint32_t h = (uint32_t)(id * 2654435761u) >> 18; /* 14 bits */
                                for (;;) {
                                    wslot_t *s = &tab[h];
                                    if (s->stamp != stamp) { s->stamp = stamp; s->key = id; s->cnt = 1; touched[nt++] = h; break; }
                                    if (s->key == id) { s->cnt++; break; }
                                    h = (h + 1) & (WTAB - 1);
                                }
                            }
                        }
                    /* unique_ids[0] = 0; unique_ids[center_id] = 0; arg-max, ties -> smallest key */
                    int32_t best = 0; uint32_t key = 0;
                    for (int t = 0; t < nt; ++t) {
                        const wslot_t *s = &tab[touched[t]];
                        if (s->key == 0 || s->key == center) continue;
                        if (s->cnt > best || (s->cnt == best && best > 0 && s->key < key)) { best = s->cnt; key = s->key; }
                    }
                    if (best > 0) {
                        if (center > key) res = ((uint64_t)key << 32) + center;
                        else res = ((uint64_t)center << 32) + key;
                    } else res = key; /* == 0 */
                }
                out[(x * Y + y) * Z + z] = res;
            }
    free(tab); free(touched);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Growable open-addressing map  uint64 id -> dense index (insertion order)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t *keys; int64_t *vals; uint64_t cap, n;
} idmap_t;

static void idmap_init(idmap_t *m, uint64_t cap) {
    m->cap = cap; m->n = 0;
    m->keys = (uint64_t *)calloc(cap, sizeof(uint64_t));
    m->vals = (int64_t *)malloc(cap * sizeof(int64_t));
}
static void idmap_free(idmap_t *m) { free(m->keys); free(m->vals); }
static void idmap_grow(idmap_t *m) {
    idmap_t n; idmap_init(&n, m->cap * 2);
    for (uint64_t i = 0; i < m->cap; ++i)
        if (m->keys[i]) {
            uint64_t h = mix64(m->keys[i]) & (n.cap - 1);
            while (n.keys[h]) h = (h + 1) & (n.cap - 1);
            n.keys[h] = m->keys[i]; n.vals[h] = m->vals[i];
        }
    n.n = m->n; idmap_free(m); *m = n;
}
/* returns dense index; *is_new set when inserted (key != 0) */
static inline int64_t idmap_get(idmap_t *m, uint64_t key, int *is_new) {
    uint64_t h = mix64(key) & (m->cap - 1);
    for (;;) {
        if (m->keys[h] == key) { *is_new = 0; return m->vals[h]; }
        if (m->keys[h] == 0) break;
        h = (h + 1) & (m->cap - 1);
    }
    if ((m->n + 1) * 2 > m->cap) { idmap_grow(m); return idmap_get(m, key, is_new); }
    m->keys[h] = key; m->vals[h] = (int64_t)m->n; m->n++; *is_new = 1;
    return m->vals[h];
}

/* per-object record, see find_object_properties_C.pyx:24-49 */
typedef struct {
    uint64_t id; int64_t size; int32_t bb[6]; int32_t rep[3];
} objrec_t;

typedef struct { objrec_t *r; uint64_t n, cap; idmap_t map; } objtab_t;

static void objtab_init(objtab_t *t) {
    t->n = 0; t->cap = 1024; t->r = (objrec_t *)malloc(t->cap * sizeof(objrec_t)); idmap_init(&t->map, 2048);
}
static void objtab_free(objtab_t *t) { free(t->r); idmap_free(&t->map); }
static inline void objtab_add(objtab_t *t, uint64_t key, int32_t x, int32_t y, int32_t z) {
    int is_new; const int64_t i = idmap_get(&t->map, key, &is_new);
    if (is_new) {
        if (t->n == t->cap) { t->cap *= 2; t->r = (objrec_t *)realloc(t->r, t->cap * sizeof(objrec_t)); }
        objrec_t *r = &t->r[t->n++];
        r->id = key; r->size = 1; /* first voxel in scan order = rep_coord; bbox [[x,y,z],[x+1,y+1,z+1]] */
        r->bb[0] = x; r->bb[1] = y; r->bb[2] = z; r->bb[3] = x + 1; r->bb[4] = y + 1; r->bb[5] = z + 1;
        r->rep[0] = x; r->rep[1] = y; r->rep[2] = z;
    } else {
        objrec_t *r = &t->r[i];
        if (x < r->bb[0]) r->bb[0] = x;
        if (y < r->bb[1]) r->bb[1] = y;
        if (z < r->bb[2]) r->bb[2] = z;
        if (x + 1 > r->bb[3]) r->bb[3] = x + 1;
        if (y + 1 > r->bb[4]) r->bb[4] = y + 1;
        if (z + 1 > r->bb[5]) r->bb[5] = z + 1;
        r->size++;
    }
}

/* result handle returned to Python; arrays are read out with orc_result_* */
typedef struct { uint64_t sub, cell; int64_t cnt; } pairrec_t;
typedef struct {
    int nch;             /* number of object tables: 1 (cell) + C (organelle channels) */
    objtab_t *tabs;      /* [0] = cell, [1..C] = organelle channels */
    pairrec_t **pairs;   /* per organelle channel, insertion order */
    uint64_t *npairs;
} orc_result_t;

ORC_API void orc_result_free(orc_result_t *r) {
    if (!r) return;
    for (int i = 0; i < r->nch; ++i) objtab_free(&r->tabs[i]);
    for (int i = 0; i + 1 < r->nch; ++i) free(r->pairs[i]);
    free(r->tabs); free(r->pairs); free(r->npairs); free(r);
}
ORC_API uint64_t orc_result_nobj(const orc_result_t *r, int tab) { return r->tabs[tab].n; }
ORC_API void orc_result_objs(const orc_result_t *r, int tab, uint64_t *ids, int64_t *sizes, int32_t *bbox, int32_t *rep) {
    const objtab_t *t = &r->tabs[tab];
    for (uint64_t i = 0; i < t->n; ++i) {
        ids[i] = t->r[i].id; sizes[i] = t->r[i].size;
        memcpy(bbox + 6 * i, t->r[i].bb, 6 * sizeof(int32_t));
        memcpy(rep + 3 * i, t->r[i].rep, 3 * sizeof(int32_t));
    }
}
ORC_API uint64_t orc_result_npairs(const orc_result_t *r, int ch) { return r->npairs[ch]; }
ORC_API void orc_result_pairs(const orc_result_t *r, int ch, uint64_t *sub, uint64_t *cell, int64_t *cnt) {
    for (uint64_t i = 0; i < r->npairs[ch]; ++i) {
        sub[i] = r->pairs[ch][i].sub; cell[i] = r->pairs[ch][i].cell; cnt[i] = r->pairs[ch][i].cnt;
    }
}

/* ------------------------------------------------------------------------------------------
 * find_object_properties -- syconn/extraction/find_object_properties_C.pyx:24-49
 * ------------------------------------------------------------------------------------------ */
ORC_API orc_result_t *orc_find_object_properties(const void *arr, int elem_bytes, const int64_t shape[3],
                                                 const int64_t st[3]) {
    orc_result_t *r = (orc_result_t *)calloc(1, sizeof(orc_result_t));
    r->nch = 1; r->tabs = (objtab_t *)malloc(sizeof(objtab_t)); objtab_init(&r->tabs[0]);
    r->pairs = NULL; r->npairs = NULL;
    for (int64_t x = 0; x < shape[0]; ++x)
        for (int64_t y = 0; y < shape[1]; ++y)
            for (int64_t z = 0; z < shape[2]; ++z) {
                const uint64_t key = ld_label(arr, elem_bytes, x * st[0] + y * st[1] + z * st[2]);
                if (key == 0) continue;
                objtab_add(&r->tabs[0], key, (int32_t)x, (int32_t)y, (int32_t)z);
            }
    return r;
}

/* pair map: (sub, cell) -> dense index */
typedef struct { uint64_t *ks, *kc; int64_t *vals; uint64_t cap, n; } pairmap_t;
static void pairmap_init(pairmap_t *m, uint64_t cap) {
    m->cap = cap; m->n = 0;
    m->ks = (uint64_t *)calloc(cap, sizeof(uint64_t)); m->kc = (uint64_t *)calloc(cap, sizeof(uint64_t));
    m->vals = (int64_t *)malloc(cap * sizeof(int64_t));
}
static void pairmap_free(pairmap_t *m) { free(m->ks); free(m->kc); free(m->vals); }
static int64_t pairmap_get(pairmap_t *m, uint64_t s, uint64_t c, int *is_new);
static void pairmap_grow(pairmap_t *m) {
    pairmap_t n; pairmap_init(&n, m->cap * 2);
    for (uint64_t i = 0; i < m->cap; ++i)
        if (m->ks[i]) {
            uint64_t h = mix64(m->ks[i] * 0x9E3779B97F4A7C15ULL ^ m->kc[i]) & (n.cap - 1);
            while (n.ks[h]) h = (h + 1) & (n.cap - 1);
            n.ks[h] = m->ks[i]; n.kc[h] = m->kc[i]; n.vals[h] = m->vals[i];
        }
    n.n = m->n; pairmap_free(m); *m = n;
}
static int64_t pairmap_get(pairmap_t *m, uint64_t s, uint64_t c, int *is_new) {
    uint64_t h = mix64(s * 0x9E3779B97F4A7C15ULL ^ c) & (m->cap - 1);
    for (;;) {
        if (m->ks[h] == s && m->kc[h] == c) { *is_new = 0; return m->vals[h]; }
        if (m->ks[h] == 0) break;
        h = (h + 1) & (m->cap - 1);
    }
    if ((m->n + 1) * 2 > m->cap) { pairmap_grow(m); return pairmap_get(m, s, c, is_new); }
    m->ks[h] = s; m->kc[h] = c; m->vals[h] = (int64_t)m->n; m->n++; *is_new = 1;
    return m->vals[h];
}

/* ------------------------------------------------------------------------------------------
 * map_subcell_extract_props -- syconn/extraction/find_object_properties_C.pyx:112-192
 * props_too == 0 restates map_subcell_C (:72-109): only the overlap mapping, voxels with cell == 0 skipped.
 * subcell: [C, X, Y, Z] with element strides sst[4]; cell: [X, Y, Z] with strides cst[3].
 * Overlap (sub_id, cell_id) is counted when both are non-zero (:163-174); organelle props count every
 * organelle voxel, also over cell background (:150-162).
 * ------------------------------------------------------------------------------------------ */
ORC_API orc_result_t *orc_map_subcell_extract_props(const void *cell, const int64_t cst[3], const void *subcell,
                                                    const int64_t sst[4], int elem_bytes, int nsub,
                                                    const int64_t shape[3], int props_too) {
    orc_result_t *r = (orc_result_t *)calloc(1, sizeof(orc_result_t));
    r->nch = 1 + nsub;
    r->tabs = (objtab_t *)malloc(sizeof(objtab_t) * r->nch);
    for (int i = 0; i < r->nch; ++i) objtab_init(&r->tabs[i]);
    r->pairs = (pairrec_t **)calloc(nsub > 0 ? nsub : 1, sizeof(pairrec_t *));
    r->npairs = (uint64_t *)calloc(nsub > 0 ? nsub : 1, sizeof(uint64_t));
    pairmap_t *pm = (pairmap_t *)malloc(sizeof(pairmap_t) * (nsub > 0 ? nsub : 1));
    uint64_t *pcap = (uint64_t *)malloc(sizeof(uint64_t) * (nsub > 0 ? nsub : 1));
    for (int i = 0; i < nsub; ++i) { pairmap_init(&pm[i], 2048); pcap[i] = 1024; r->pairs[i] = (pairrec_t *)malloc(pcap[i] * sizeof(pairrec_t)); }
    for (int64_t x = 0; x < shape[0]; ++x)
        for (int64_t y = 0; y < shape[1]; ++y)
            for (int64_t z = 0; z < shape[2]; ++z) {
                const uint64_t key = ld_label(cell, elem_bytes, x * cst[0] + y * cst[1] + z * cst[2]);
                if (!props_too && key == 0) continue; /* map_subcell_C :91-92 */
                for (int ii = 0; ii < nsub; ++ii) {
                    const uint64_t sk = ld_label(subcell, elem_bytes, ii * sst[0] + x * sst[1] + y * sst[2] + z * sst[3]);
                    if (sk == 0) continue;
                    if (props_too) objtab_add(&r->tabs[1 + ii], sk, (int32_t)x, (int32_t)y, (int32_t)z);
                    if (key != 0) {
                        int is_new; const int64_t pi = pairmap_get(&pm[ii], sk, key, &is_new);
                        if (is_new) {
                            if (r->npairs[ii] == pcap[ii]) { pcap[ii] *= 2; r->pairs[ii] = (pairrec_t *)realloc(r->pairs[ii], pcap[ii] * sizeof(pairrec_t)); }
                            pairrec_t *p = &r->pairs[ii][r->npairs[ii]++]; p->sub = sk; p->cell = key; p->cnt = 1;
                        } else r->pairs[ii][pi].cnt++;
                    }
                }
                if (!props_too || key == 0) continue;
                objtab_add(&r->tabs[0], key, (int32_t)x, (int32_t)y, (int32_t)z);
            }
    for (int i = 0; i < nsub; ++i) pairmap_free(&pm[i]);
    free(pm); free(pcap);
    return r;
}

/* ------------------------------------------------------------------------------------------
 * extract_cs_syntype -- syconn/extraction/block_processing_C.pyx:78-158   ("next" row f1)
 * One scan in (x,y,z) order.  Per contact id (cs_seg != 0): props as find_object_properties (table 0).
 * Where syn_mask != 0 additionally: props of the synaptic part (table 1), the voxel (in scan order, with
 * `offset` added by the Python side), and counts of asym_mask == 1 / sym_mask == 1 voxels.
 * Result handle: tabs[0] = cs props, tabs[1] = syn props; the syn voxel tuples are returned through
 * orc_syntype_voxels (key, x, y, z, asym==1, sym==1) in scan order.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t key; int32_t x, y, z; uint8_t asym, sym; } synvox_t;
typedef struct { orc_result_t *res; synvox_t *vox; uint64_t nvox, cap; } orc_syntype_t;

ORC_API orc_syntype_t *orc_extract_cs_syntype(const void *cs, int elem_bytes, const int64_t cst[3], const uint8_t *syn,
                                              const int64_t sst[3], const uint8_t *asym, const int64_t ast[3],
                                              const uint8_t *sym, const int64_t yst[3], const int64_t shape[3]) {
    orc_syntype_t *r = (orc_syntype_t *)calloc(1, sizeof(orc_syntype_t));
    r->res = (orc_result_t *)calloc(1, sizeof(orc_result_t));
    r->res->nch = 2;
    r->res->tabs = (objtab_t *)malloc(2 * sizeof(objtab_t));
    objtab_init(&r->res->tabs[0]);
    objtab_init(&r->res->tabs[1]);
    r->res->pairs = (pairrec_t **)calloc(1, sizeof(pairrec_t *));
    r->res->npairs = (uint64_t *)calloc(1, sizeof(uint64_t));
    r->cap = 1024;
    r->vox = (synvox_t *)malloc(r->cap * sizeof(synvox_t));
    for (int64_t x = 0; x < shape[0]; ++x)
        for (int64_t y = 0; y < shape[1]; ++y)
            for (int64_t z = 0; z < shape[2]; ++z) {
                const uint64_t key = ld_label(cs, elem_bytes, x * cst[0] + y * cst[1] + z * cst[2]);
                if (key == 0) continue;
                objtab_add(&r->res->tabs[0], key, (int32_t)x, (int32_t)y, (int32_t)z);
                if (syn[x * sst[0] + y * sst[1] + z * sst[2]] == 0) continue;
                objtab_add(&r->res->tabs[1], key, (int32_t)x, (int32_t)y, (int32_t)z);
                if (r->nvox == r->cap) { r->cap *= 2; r->vox = (synvox_t *)realloc(r->vox, r->cap * sizeof(synvox_t)); }
                synvox_t *v = &r->vox[r->nvox++];
                v->key = key; v->x = (int32_t)x; v->y = (int32_t)y; v->z = (int32_t)z;
                v->asym = asym[x * ast[0] + y * ast[1] + z * ast[2]] == 1;
                v->sym = sym[x * yst[0] + y * yst[1] + z * yst[2]] == 1;
            }
    return r;
}
ORC_API orc_result_t *orc_syntype_result(orc_syntype_t *r) { return r->res; }
ORC_API uint64_t orc_syntype_nvox(const orc_syntype_t *r) { return r->nvox; }
ORC_API void orc_syntype_voxels(const orc_syntype_t *r, uint64_t *key, int32_t *xyz, uint8_t *asym, uint8_t *sym) {
    for (uint64_t i = 0; i < r->nvox; ++i) {
        key[i] = r->vox[i].key; xyz[3 * i] = r->vox[i].x; xyz[3 * i + 1] = r->vox[i].y; xyz[3 * i + 2] = r->vox[i].z;
        asym[i] = r->vox[i].asym; sym[i] = r->vox[i].sym;
    }
}
ORC_API void orc_syntype_free(orc_syntype_t *r) {
    if (!r) return;
    /* tabs[0..1] only: nch == 2 but there is a single (empty) pair slot */
    objtab_free(&r->res->tabs[0]); objtab_free(&r->res->tabs[1]);
    free(r->res->tabs); free(r->res->pairs); free(r->res->npairs); free(r->res);
    free(r->vox); free(r);
}
