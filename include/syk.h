/*
 * syk.h -- C ABI of libsyk.so: B200 (sm_100a) kernels for SyConn's chunked label-volume extraction path.
 *
 * Every entry point replaces one function of the reference's two Cython extension modules (the reference's
 * plugin boundary for this path) or one step of its Python merge logic.  Citations are file:line under the
 * reference repository (StructuralNeurobiologyLab/SyConn).
 *
 * Conventions
 *   - All functions return 0 on success, a negative SYK_E* code otherwise; syk_last_error() gives the
 *     thread-local message.  Nothing falls back to the CPU: without a usable CUDA device every compute entry
 *     point fails with SYK_ENODEV.
 *   - Plain pointers and sizes only.  "dev" pointers are device memory (e.g. torch.Tensor.data_ptr()),
 *     "host" pointers are host memory.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - Label volumes are addressed as base[x*strides[0] + y*strides[1] + z*strides[2]], strides in ELEMENTS,
 *     elem_bytes is 4 (uint32) or 8 (uint64): any NumPy view of a dense block (e.g. ZYX memory seen as XYZ,
 *     syconn/proc/sd_proc.py:629,641) is passed as is.  Results are always in the LOGICAL (x,y,z) order of
 *     the view, exactly like the reference's typed memoryviews.
 *   - Device-pointer functions are asynchronous on `stream`; *_host functions are synchronous (one call =
 *     one chunk, like the reference's `def` functions that hold the GIL for the whole call).
 */
#ifndef SYK_H
#define SYK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYK_VERSION 100

#define SYK_OK 0
#define SYK_EINVAL (-1)    /* bad argument (even stencil, shape mismatch, bad dtype size ...) */
#define SYK_ECUDA (-2)     /* CUDA runtime error, see syk_last_error() */
#define SYK_ENODEV (-3)    /* no CUDA device: there is no CPU fallback */
#define SYK_EOVERFLOW (-4) /* a hash table ran out of slots: retry with a larger capacity */
#define SYK_ENOMEM (-5)

/* Per-object record (64 bytes).  Array-level form of the reference's three dicts
 * rep_coords / bounding_box / sizes (syconn/extraction/find_object_properties_C.pyx:24-49) and the unit
 * that is exchanged between GPUs and merged (syconn/proc/sd_proc.py:1248-1273). */
typedef struct syk_record {
    uint64_t id;        /* object id (never 0) */
    uint64_t count;     /* voxel count ("size"; the reference's C int, here 64 bit) */
    uint64_t rep_key;   /* (chunk_seq << 40) | (2^40-1 - linear index (x*Sy+y)*Sz+z inside that chunk's call);
                           max over rep_key == "first voxel in scan order of the LAST chunk holding the id"
                           (find_object_properties_C.pyx:48 within a chunk, sd_proc.py:1261 across chunks) */
    int32_t bb_min[3];  /* global coordinates (origin added), inclusive */
    int32_t bb_max[3];  /* exclusive (max+1), as find_object_properties_C.pyx:39-41 */
    int32_t rep[3];     /* decoded global rep_coord; filled by syk_table_export / syk_records_decode_rep */
    uint32_t chunk_seq; /* == rep_key >> 40 */
} syk_record_t;

/* Overlap record (32 bytes): one entry of the reference's nested mapping dict
 * {sub_id: {cell_id: count}} (find_object_properties_C.pyx:163-174). */
typedef struct syk_pair {
    uint64_t sub_id;
    uint64_t cell_id;
    uint64_t count;
    uint64_t _pad;
} syk_pair_t;

/* One synaptic voxel of a contact site (extract_cs_syntype, block_processing_C.pyx:119-157): 32 bytes */
typedef struct syk_synvox {
    uint64_t id;     /* contact-site id of the voxel */
    uint64_t lin;    /* linear index (x*Sy + y)*Sz + z inside the call's volume (scan order of the reference) */
    uint64_t flags;  /* bit 0: asym_mask == 1, bit 1: sym_mask == 1 */
    uint64_t _pad;
} syk_synvox_t;

/* geometry of one chunk call, needed to decode rep_key -> rep[3] */
typedef struct syk_chunk_geom {
    int64_t origin[3];
    int64_t shape[3];
} syk_chunk_geom_t;

typedef struct syk_table syk_table_t; /* device open-addressing table id -> record */
typedef struct syk_pairs syk_pairs_t; /* device open-addressing table (sub,cell) -> count */

/* ---- library ------------------------------------------------------------------------------------------- */
int syk_version(void);
const char *syk_last_error(void);
int syk_device_count(void);
int syk_set_device(int device);

/* ---- tables -------------------------------------------------------------------------------------------- */
/* capacity is rounded up to a power of two (>= 1024); keep the load factor <= 0.5 */
int syk_table_create(syk_table_t **out, uint64_t capacity);
int syk_table_destroy(syk_table_t *t);
int syk_table_clear(syk_table_t *t, void *stream);
uint64_t syk_table_capacity(const syk_table_t *t);
/* number of occupied slots / overflow flag; synchronises `stream` */
int syk_table_count(syk_table_t *t, void *stream, uint64_t *n_out_host, int *overflow_out_host);
/* compact the table into records_dev[0..n) (order unspecified, like the reference's unordered_map iteration);
 * geoms_host[n_geoms] is indexed by chunk_seq to decode rep.  Synchronises `stream`. */
int syk_table_export(syk_table_t *t, const syk_chunk_geom_t *geoms_host, uint32_t n_geoms, syk_record_t *records_dev,
                     uint64_t max_records, uint64_t *n_out_host, void *stream);
/* Asynchronous variant for chunk loops: appends the table's records to a device log at position *counter_dev
 * (a device uint64 that accumulates across calls; records beyond max_records are dropped but still counted, so
 * the host can detect a short log; if the source table overflowed, bit 62 of the counter is set).  geom is the calling
 * chunk's geometry (one entry, applied to every chunk_seq). */
int syk_table_append_records(syk_table_t *t, const syk_chunk_geom_t *geom_host, syk_record_t *log_dev, uint64_t max_records,
                             uint64_t *counter_dev, void *stream);
/* fold records into a table: count summed, bbox min/max, rep = max rep_key
 * (merge_prop_dicts + final reduction, syconn/proc/sd_proc.py:1248-1273, :939-945, :1172-1177) */
int syk_table_merge_records(syk_table_t *t, const syk_record_t *records_dev, uint64_t n, void *stream);
/* owner(id) = hash(id) mod n_owners; reorders records into contiguous per-owner buckets (for the all-to-all).
 * counts_dev[n_owners] receives the bucket sizes.  out_dev must hold n records. */
int syk_records_bucket(const syk_record_t *records_dev, uint64_t n, uint32_t n_owners, syk_record_t *out_dev,
                       uint64_t *counts_dev, void *stream);
int syk_records_decode_rep(syk_record_t *records_dev, uint64_t n, const syk_chunk_geom_t *geoms_host, uint32_t n_geoms,
                           void *stream);
/* Sort n device records by ascending id (radix sort; the contact-site worker's per-id loop, cs_extraction_steps.py:440,
 * runs over the keys of a dict built in ascending id order).  Ids are unique in a table export, so the order is total. */
int syk_records_sort_by_id(syk_record_t *records_dev, uint64_t n, void *stream);

/* syk_table_append_records with the worker's small-object drop (syconn/proc/sd_proc.py:650-661, :667-680): an object that
 * lies purely inside the chunk `geom_host` (its box touches none of the six faces) and has fewer than min_vx voxels is not
 * appended.  min_vx <= 1: identical to syk_table_append_records. */
int syk_table_append_records_min_vx(syk_table_t *t, const syk_chunk_geom_t *geom_host, syk_record_t *log_dev, uint64_t max_records,
                                    uint64_t *counter_dev, uint64_t min_vx, void *stream);

int syk_pairs_create(syk_pairs_t **out, uint64_t capacity);
uint64_t syk_pairs_capacity(const syk_pairs_t *t);
int syk_pairs_destroy(syk_pairs_t *t);
int syk_pairs_clear(syk_pairs_t *t, void *stream);
int syk_pairs_export(syk_pairs_t *t, syk_pair_t *pairs_dev, uint64_t max_pairs, uint64_t *n_out_host, void *stream);
int syk_pairs_append(syk_pairs_t *t, syk_pair_t *log_dev, uint64_t max_pairs, uint64_t *counter_dev, void *stream);
/* syk_pairs_append that skips the pairs of organelle objects removed by the small-object drop (sd_proc.py:678-679);
 * sub_t is the chunk's organelle table (the one syk_table_append_records_min_vx exports) */
int syk_pairs_append_min_vx(syk_pairs_t *t, syk_table_t *sub_t, const syk_chunk_geom_t *geom_host, uint64_t min_vx,
                            syk_pair_t *log_dev, uint64_t max_pairs, uint64_t *counter_dev, void *stream);
/* pairs_dev[i]._pad = total size of organelle pairs_dev[i].sub_id in the owner's final table (0 = not in the table):
 * the normalisation step of the mapping inversion, sd_proc.py:1054-1084 (ratio = count / organelle size) */
int syk_pairs_attach_size(syk_pair_t *pairs_dev, uint64_t n, syk_table_t *sub_final, void *stream);
/* merge_map_dicts (syconn/proc/sd_proc.py:1300-1322): counts summed per (sub_id, cell_id) */
int syk_pairs_merge(syk_pairs_t *t, const syk_pair_t *pairs_dev, uint64_t n, void *stream);
/* owner = hash(sub_id) mod n_owners (the reference reduces per organelle object, sd_proc.py:830-853) */
int syk_pairs_bucket(const syk_pair_t *pairs_dev, uint64_t n, uint32_t n_owners, syk_pair_t *out_dev, uint64_t *counts_dev,
                     void *stream);

/* ---- hot path, device buffers -------------------------------------------------------------------------- */
/* find_object_properties(chunk)  -- syconn/extraction/find_object_properties_C.pyx:24-49
 * accumulates into `t`: voxel count, bounding box, first voxel in (x,y,z) scan order per non-zero id.
 * origin is added to all coordinates (merge_prop_dicts' offset, sd_proc.py:1259-1266). */
int syk_find_object_properties(syk_table_t *t, const void *labels_dev, int elem_bytes, const int64_t shape[3],
                               const int64_t strides[3], const int64_t origin[3], uint32_t chunk_seq, void *stream);

/* map_subcell_extract_props(ch, subcell_chs) -- find_object_properties_C.pyx:112-192
 * (cell_t == NULL and sub_t == NULL  =>  map_subcell_C, :72-109: mapping only)
 * subcell_dev[c] are n_sub device pointers sharing shape/strides `sub_strides`. */
int syk_map_subcell_extract_props(syk_table_t *cell_t, syk_table_t *const *sub_t, syk_pairs_t *const *pair_t,
                                  const void *cell_dev, const int64_t cell_strides[3], const void *const *subcell_dev,
                                  const int64_t sub_strides[3], int n_sub, int elem_bytes, const int64_t shape[3],
                                  const int64_t origin[3], uint32_t chunk_seq, void *stream);

/* detect_seg_boundaries(arr) -- syconn/extraction/find_object_properties.py:424-455; out: uint8 C-contiguous */
int syk_detect_seg_boundaries(const void *arr_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                              uint8_t *out_dev, void *stream);

/* process_block_nonzero(edges, arr, stencil) -- syconn/extraction/block_processing_C.pyx:53-75 (+ kernel :21-49)
 * edges: uint8 or uint32 (edge_bytes 1|4).  arr: uint32, or uint64 truncated to uint32 (the caller-side
 * .astype(np.uint32), cs_extraction_steps.py:385-387).  out: uint64 [shape - stencil + 1], strides in elements. */
int syk_process_block_nonzero(const void *edges_dev, int edge_bytes, const int64_t edge_strides[3], const void *arr_dev,
                              int elem_bytes, const int64_t arr_strides[3], const int64_t shape[3],
                              const int32_t stencil[3], uint64_t *out_dev, const int64_t out_strides[3], void *stream);

/* detect_cs(arr) -- syconn/extraction/find_object_properties.py:458-472: fused boundary mask + stencil */
int syk_detect_cs(const void *arr_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                  const int32_t stencil[3], uint64_t *out_dev, const int64_t out_strides[3], void *stream);

/* extract_cs_syntype(cs_seg, syn_mask, asym_mask, sym_mask, offset) -- block_processing_C.pyx:78-158 ("next" row f1)
 * cs_t accumulates the contact-site props; every voxel with cs != 0 and syn_mask != 0 is appended to vox_dev
 * (order unspecified; sort by (id, lin) to obtain the reference's per-id voxel lists).  The synaptic props, the
 * asym / sym counts and the voxel lists all follow from these tuples.  Masks are uint8 volumes of the same shape. */
int syk_extract_cs_syntype(syk_table_t *cs_t, const void *cs_dev, int elem_bytes, const int64_t shape[3],
                           const int64_t cs_strides[3], const uint8_t *syn_dev, const int64_t syn_strides[3],
                           const uint8_t *asym_dev, const int64_t asym_strides[3], const uint8_t *sym_dev,
                           const int64_t sym_strides[3], const int64_t origin[3], uint32_t chunk_seq, syk_synvox_t *vox_dev,
                           uint64_t max_vox, uint64_t *counter_dev, void *stream);

/* Same, and the props of the synaptic part of every contact -- extract_cs_syntype's second return value
 * ([rc, bb, size] over the voxels with syn_mask != 0, block_processing_C.pyx:117-137) -- accumulated in syn_t by the same
 * pass (NULL: like syk_extract_cs_syntype).  Fewer than 2^31 voxels per call when syn_t is given. */
int syk_extract_cs_syntype_props(syk_table_t *cs_t, syk_table_t *syn_t, const void *cs_dev, int elem_bytes, const int64_t shape[3],
                                 const int64_t cs_strides[3], const uint8_t *syn_dev, const int64_t syn_strides[3],
                                 const uint8_t *asym_dev, const int64_t asym_strides[3], const uint8_t *sym_dev,
                                 const int64_t sym_strides[3], const int64_t origin[3], uint32_t chunk_seq, syk_synvox_t *vox_dev,
                                 uint64_t max_vox, uint64_t *counter_dev, void *stream);

/* detect_contact_partners(seg_arr, edge_arr, offset) -- syconn/extraction/find_object_properties.py:371-421, the
 * numba twin of process_block_nonzero behind detect_cs_64bit (:347-368).  Same window histogram; ties go to the id met
 * first in the x, y, z scan of the window (numba typed.Dict keeps insertion order); out = (centre << 32) | partner of
 * uint32 labels, 0 = no partner.  edges_dev == NULL: boundary mask (detect_seg_boundaries) computed on the fly. */
int syk_detect_contact_partners(const void *edges_dev, int edge_bytes, const int64_t edge_strides[3], const void *arr_dev,
                                const int64_t arr_strides[3], const int64_t shape[3], const int32_t stencil[3],
                                uint64_t *out_dev, const int64_t out_strides[3], void *stream);
/* (centre << 32) | partner  ->  out[i][0..1] = [min, max] of the ids (ids_dev[label - 1], NULL = labels are the ids):
 * the XYZC layout of detect_cs_64bit (find_object_properties.py:418-420) */
int syk_cs64_unpack(const uint64_t *packed_dev, uint64_t n, const uint64_t *ids_dev, uint64_t *out_dev, void *stream);
/* dense uint32 labels for a block of 64-bit ids (support of the 64-bit variants): labels_out[i] in 1..n_ids (0 for id
 * 0), ids_out[label - 1] = id.  `t` must be empty.  SYK_EOVERFLOW when the table or ids_out is too small. */
int syk_dense_relabel(syk_table_t *t, const void *vol_dev, int elem_bytes, uint64_t n, uint32_t *labels_out_dev,
                      uint64_t *ids_out_dev, uint64_t max_ids, uint64_t *n_ids_out_host, void *stream);

/* Per-contact closing + dilation ("next" row f2): the per-id loop of the contact-site worker,
 * syconn/extraction/cs_extraction_steps.py:439-461.  For every id of ids_host[0..n_ids), in list order: the id's mask
 * inside its bounding box padded by n_closings (clipped to the volume) is closed (scipy.ndimage.binary_closing,
 * iterations = n_closings) and dilated (binary_dilation, iterations = n_dilations), 6-neighbourhood, border value 0;
 * voxels of the result that are still background take the id (the first id in list order wins, as in the sequential
 * loop).  cs_dev is modified in place.  bbox_host[i] = {min[3], max_exclusive[3]} in voxels of the volume passed, as
 * find_object_properties returns it (it must contain every voxel of the id).  Synchronises `stream`. */
int syk_close_contacts(void *cs_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3], const uint64_t *ids_host,
                       const int32_t *bbox_host, uint64_t n_ids, int n_closings, int n_dilations, void *stream);

/* ---- synthetic label volumes (bench/test inputs; bit-identical to syconn_b200/synth.py) ------------------ */
/* kind 0: cell supervoxels (ids < 2^32, ~3% background); kind 1..: organelle channel (sparse 64-bit ids) */
int syk_synth_labels(void *out_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                     const int64_t origin[3], const int32_t pitch[3], int32_t warp_amp, uint64_t seed, int kind,
                     int density16, void *stream);

/* ---- hot path, HOST buffers (what a binding of the reference's Cython functions calls) ------------------ */
/* Results are malloc'ed by the library; release with syk_free().  `labels` must be a dense block viewed
 * with `strides` (a permutation of a C-contiguous layout); nbytes = elem_bytes * prod(shape). */
int syk_find_object_properties_host(const void *labels_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                    uint64_t capacity_hint, syk_record_t **records_out, uint64_t *n_out);
int syk_map_subcell_extract_props_host(const void *cell_host, const int64_t cell_strides[3], const void *subcell_host,
                                       const int64_t sub_strides[4], int n_sub, int elem_bytes, const int64_t shape[3],
                                       int props_too, uint64_t capacity_hint, syk_record_t **cell_records_out,
                                       uint64_t *n_cell_out, syk_record_t **sub_records_out /*[n_sub]*/,
                                       uint64_t *n_sub_out /*[n_sub]*/, syk_pair_t **pairs_out /*[n_sub]*/,
                                       uint64_t *n_pairs_out /*[n_sub]*/);
int syk_detect_cs_host(const void *arr_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                       const int32_t stencil[3], uint64_t *out_host /* C-contiguous [shape-stencil+1] */);
/* detect_cs followed by find_object_properties(contacts) -- the pair of calls of cs_extraction_steps.py:391,439 --
 * without moving the contact volume over PCIe twice */
int syk_detect_cs_props_host(const void *arr_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                             const int32_t stencil[3], uint64_t *out_host, syk_record_t **records_out, uint64_t *n_out);
int syk_process_block_nonzero_host(const void *edges_host, int edge_bytes, const int64_t edge_strides[3],
                                   const void *arr_host, int elem_bytes, const int64_t arr_strides[3],
                                   const int64_t shape[3], const int32_t stencil[3], uint64_t *out_host);
int syk_extract_cs_syntype_host(const void *cs_host, int elem_bytes, const int64_t shape[3], const int64_t cs_strides[3],
                                const uint8_t *syn_host, const int64_t syn_strides[3], const uint8_t *asym_host,
                                const int64_t asym_strides[3], const uint8_t *sym_host, const int64_t sym_strides[3],
                                syk_record_t **cs_records_out, uint64_t *n_cs_out, syk_synvox_t **vox_out, uint64_t *n_vox_out);
int syk_detect_seg_boundaries_host(const void *arr_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                                   uint8_t *out_host);
/* detect_cs_64bit(arr) / detect_contact_partners(seg, edges, symmetric offset) -- find_object_properties.py:347-421.
 * out_host: C-contiguous uint64 [shape - stencil + 1][2] = sorted partner ids (0, 0 where there is no contact).
 * edges_host == NULL: detect_seg_boundaries(arr) is used (detect_cs_64bit). */
int syk_detect_contact_partners_host(const void *edges_host, int edge_bytes, const int64_t edge_strides[3], const void *arr_host,
                                     int elem_bytes, const int64_t strides[3], const int64_t shape[3], const int32_t stencil[3],
                                     uint64_t *out_host);
/* find_object_properties_cs_64bit(cs_seg) -- find_object_properties.py:197-269: properties per partner PAIR of an XYZC
 * contact volume (C = 2).  records_out[i].id is internal; partners_out[2 i .. 2 i + 1] are the pair's ids. */
int syk_find_object_properties_cs_64bit_host(const uint64_t *cs_host, const int64_t shape[3], const int64_t strides[4],
                                             syk_record_t **records_out, uint64_t **partners_out, uint64_t *n_out);
/* Same with the boxes taken from DEVICE records (a table export of find_object_properties(contacts) in volume-local
 * coordinates, sorted by id with syk_records_sort_by_id): box planning runs on the device, the call reads back 32 bytes and
 * does not wait for its kernels.  Processing order = record order. */
int syk_close_contacts_records(void *cs_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                               const syk_record_t *records_dev, uint64_t n_ids, int n_closings, int n_dilations, void *stream);
/* host-buffer form of syk_close_contacts: cs_host (dense block) is updated in place */
int syk_close_contacts_host(void *cs_host, int elem_bytes, const int64_t shape[3], const int64_t strides[3], const uint64_t *ids_host,
                            const int32_t *bbox_host, uint64_t n_ids, int n_closings, int n_dilations);
void syk_free(void *p);

/* ---- organelle instance segmentation, first slice (row f4) --------------------------------------------------- */
/* scipy.ndimage.label(vol > threshold) with the default 6-connectivity structure, as called by
 * _object_segmentation_thread (syconn/extraction/object_extraction_steps.py:302-303, :350-352).  labels_dev receives
 * uint32 labels 1..n in the order of each component's first voxel in LOGICAL (x, y, z) scan order -- bit-identical to
 * scipy -- whatever the memory order of the view; *n_labels_host = n.  elem_bytes 1, 2, 4 or 8; fewer than 2^32 - 16
 * voxels per call.  Synchronises `stream` (the number of components sizes the ranking buffers). */
int syk_label_components(const void *vol_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3], uint64_t threshold,
                         uint32_t *labels_dev, const int64_t label_strides[3], uint64_t *n_labels_host, void *stream);
/* The co-located label pairs of two label blocks over the same voxels (the stitch region of two neighbouring chunks,
 * _make_stitch_list_thread, object_extraction_steps.py:583-606): for every voxel with a != 0 and b != 0 the pair
 * (a + a_offset, b + b_offset) is counted in `pairs` (export it with syk_pairs_export). */
int syk_label_overlap_pairs(syk_pairs_t *pairs, const uint32_t *a_dev, const int64_t a_strides[3], const uint32_t *b_dev,
                            const int64_t b_strides[3], const int64_t shape[3], uint64_t a_offset, uint64_t b_offset, void *stream);

/* relabel_vol(vol, label_map) of syconn/extraction/block_processing_C.pyx:161-170 for dense label maps: every label l with
 * 0 < l < lut_len becomes lut_dev[l], other labels stay (the marker clean-up of _object_segmentation_thread,
 * object_extraction_steps.py:325-343).  In place, any strides. */
int syk_label_map(uint32_t *labels_dev, const int64_t shape[3], const int64_t strides[3], const uint32_t *lut_dev, uint64_t lut_len,
                  void *stream);
/* Binary morphology of a whole thresholded volume (row f4): apply_morphological_operations(vol, morph_ops,
 * mop_kwargs=dict(structure=struct)) of syconn/proc/image.py:485-507 as called on the 0/1 volume by
 * _object_segmentation_thread (object_extraction_steps.py:312-358), i.e. _multi_mop_findobjects (image.py:358-437) with the
 * single object 1: every op runs inside the bounding box of the current foreground -- dilation / closing on the box padded
 * by `iterations` zeros and cropped back, erosion / opening on the box itself, border_value 0 -- so the volume is
 * bit-identical to the reference's.  vol_dev [X,Y,Z] (any dense or strided view, elem_bytes 1/2/4/8) is updated in place and
 * must hold only 0 and 1 (SYK_EINVAL otherwise: multi-label overlays are not a device path).  structure_host: C-ordered
 * uint8 [sx,sy,sz], odd extents, point-symmetric (e.g. get_aniso_struct, image.py:522-539).  ops_host[i]: 0 binary_erosion,
 * 1 binary_dilation, 2 binary_opening, 3 binary_closing, iters_host[i] its iterations (runs of equal ops merged by the
 * caller like _count_subsequent_mops).  Synchronises `stream` once per op (bounding box). */
int syk_binary_morph_ops(void *vol_dev, int elem_bytes, const int64_t shape[3], const int64_t strides[3],
                         const uint8_t *structure_host, const int64_t structure_shape[3], const int32_t *ops_host,
                         const int32_t *iters_host, int n_ops, void *stream);

/* ---- storage codec (row f3; host code, no GPU needed) ----------------------------------------------------- */
/* LZ4 block format, the codec behind the reference's CompressedStorage / VoxelStorageDyn values
 * (python-lz4 `lz4.block.compress/decompress`, syconn/handler/compression.py:83-127, backend/storage.py:52-93).
 * Raw blocks, without python-lz4's 4-byte length prefix (syconn_b200/handler/compression.py adds it). */
uint64_t syk_lz4_compress_bound(uint64_t n);
int syk_lz4_compress_block(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *out_n);
int syk_lz4_decompress_block(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap, uint64_t *out_n);

#ifdef __cplusplus
}
#endif
#endif /* SYK_H */
