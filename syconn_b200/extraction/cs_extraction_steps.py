"""The device-side pieces of ``syconn.extraction.cs_extraction_steps`` (the contact-site worker
``_contact_site_extraction_thread``, cs_extraction_steps.py:300-500).  Only the numeric inner loops are provided; dataset
I/O (knossos_utils) stays with the reference."""
import numpy as np

from .. import _lib, global_params
from ._host import check_label_array, dense_view, estrides


def _boxes(bb_dc):
    ids = np.fromiter((int(k) for k in bb_dc.keys()), np.uint64, len(bb_dc))
    bb = np.ascontiguousarray(np.array(list(bb_dc.values()), np.int64).reshape(-1, 2, 3).astype(np.int32))
    return ids, bb


def close_contact_sites(contacts, bb_dc=None, n_closings=None, cs_dilation=None):
    """The closing / dilation loop of the contact-site worker (syconn/extraction/cs_extraction_steps.py:436-461) as one
    call: for every id of ``bb_dc`` (``find_object_properties(contacts)[1]``; computed and sorted by id when ``None``), in the dict's
    iteration order, the id's mask inside its bounding box padded by ``n_closings`` is closed
    (``scipy.ndimage.binary_closing(iterations=n_closings)``) and dilated (``binary_dilation(iterations=cs_dilation)``)
    and written to the voxels that are still background.  ``contacts`` is modified in place and returned.

    Defaults follow the worker: ``n_closings = max(cs_filtersize // 2)`` (:383, :440), ``cs_dilation`` from the config
    (:377).  Where the closed regions of two ids overlap in background, the id that comes first in ``bb_dc`` wins, as in
    the sequential loop (the reference's own order is the hash-map order of its Cython ``unordered_map``)."""
    contacts = check_label_array(contacts, "contacts", 3)
    if n_closings is None:
        n_closings = int(max(np.array(global_params.config['cell_objects']['cs_filtersize']) // 2))
    if cs_dilation is None:
        cs_dilation = int(global_params.config['cell_objects']['cs_dilation'])
    if bb_dc is None:
        from .find_object_properties import find_object_properties
        bb_dc = find_object_properties(contacts)[1]
        bb_dc = {k: bb_dc[k] for k in sorted(bb_dc)}    # a reproducible order: ascending id
    if contacts.size == 0 or len(bb_dc) == 0 or (n_closings <= 0 and cs_dilation <= 0):
        return contacts
    ids, bb = _boxes(bb_dc)
    work = dense_view(contacts)
    _lib.check(_lib.load().syk_close_contacts_host(work.ctypes.data, work.itemsize, _lib.i64(work.shape), _lib.i64(estrides(work)),
                                                   ids.ctypes.data, bb.ctypes.data, len(ids), max(int(n_closings), 0),
                                                   max(int(cs_dilation), 0)))
    if work is not contacts:        # non-dense view: the work copy goes back into the caller's array
        contacts[...] = work
    return contacts


def _to_device(a, dtype_view=None):
    """dense NumPy view -> CUDA tensor with the same strides (``zyx.swapaxes(0, 2)`` stays x-fastest on the device)."""
    import torch
    a = dense_view(np.asarray(a))
    if dtype_view is not None:
        a = a.view(dtype_view)
    return torch.from_numpy(a).cuda()


def contact_site_extraction_chunk(data, sj_d, asym_d, sym_d, offset, cs_filtersize=None, cs_dilation=None):
    """Numeric body of one iteration of the chunk loop of ``_contact_site_extraction_thread``
    (syconn/extraction/cs_extraction_steps.py:381-486), device resident between the stages:

      ``detect_cs(data)`` (:391) -> ``find_object_properties(contacts)`` (:439) -> per-id closing / dilation
      (:440-461) -> ``extract_cs_syntype`` on the volumes cropped by ``overlap`` (:465-470) -> contact-site and
      synapse segmentations of the chunk (:472-479).

    ``data``: uint32 cell supervoxels, block ``size + 2 * overlap + 2 * stencil_offset`` loaded at
    ``offset - stencil_offset`` (:385-387); ``sj_d`` / ``asym_d`` / ``sym_d``: uint8 masks of the block ``size + 2 *
    overlap`` at ``offset`` (:394-434); ``offset`` = ``chunk.coordinates - overlap`` (:382).  Ids are closed in ascending
    id order.  Returns ``(curr_cs_p, curr_syn_p, asym_cnt, sym_cnt, curr_syn_vx, cs_seg, syn_seg)``: the five results of
    ``extract_cs_syntype`` plus the cropped contact-site volume and its intersection with ``sj_d`` (XYZ, uint64) that the
    worker writes to the ``cs`` / ``syn`` KnossosDatasets."""
    import torch
    from .. import device as dev
    from ._host import syntype_to_dicts
    st = np.array(global_params.config['cell_objects']['cs_filtersize'] if cs_filtersize is None else cs_filtersize)
    assert np.sum(st % 2) == 3
    if cs_dilation is None:
        cs_dilation = int(global_params.config['cell_objects']['cs_dilation'])
    overlap = int(max(st // 2))
    data = np.asarray(data)
    if data.dtype != np.uint32:
        raise ValueError(f"Buffer dtype mismatch, expected 'uint32_t' but got '{data.dtype}' (data)")
    oshape = tuple(int(data.shape[i] - st[i] + 1) for i in range(3))
    masks = []
    for name, m in (("sj_d", sj_d), ("asym_d", asym_d), ("sym_d", sym_d)):
        m = np.asarray(m)
        if m.dtype != np.uint8:
            raise ValueError(f"Buffer dtype mismatch, expected 'uint8_t' but got '{m.dtype}' ({name})")
        assert m.shape == oshape, f"{name} must have the shape of the contact volume {oshape}, got {m.shape}"
        masks.append(_to_device(m))
    assert min(oshape) > 2 * overlap, "block smaller than the overlap"
    seg = _to_device(data, np.int32)
    contacts = dev.detect_cs(seg, [int(s) for s in st])
    del seg
    tab = dev.IdTable(max(1 << 16, contacts.numel() // 64))
    try:
        while True:  # bounding boxes of the contact ids (retry with a larger table on overflow)
            dev.find_object_properties(tab, contacts)
            n, ovf = tab.count()
            if not ovf:
                break
            tab.close()
            tab = dev.IdTable(tab.capacity * 4)
        rec = dev.records_numpy(tab.export(dev.geoms([[0, 0, 0]], [list(oshape)])))
        rec = rec[np.argsort(rec["id"])]
        if len(rec):
            dev.close_contacts(contacts, rec["id"].copy(), np.stack([rec["bb_min"], rec["bb_max"]], axis=1).astype(np.int32),
                               overlap, cs_dilation)
        crop = (slice(overlap, -overlap),) * 3
        cs_c = contacts[crop]
        tab.clear()
        cshape = tuple(cs_c.shape)
        while True:
            vox = dev.extract_cs_syntype(tab, cs_c, masks[0][crop], masks[1][crop], masks[2][crop])
            n, ovf = tab.count()
            if not ovf:
                break
            tab.close()
            tab = dev.IdTable(tab.capacity * 4)
        cs_rec = dev.records_numpy(tab.export(dev.geoms([[0, 0, 0]], [list(cshape)])))
    finally:
        tab.close()
    v = vox.cpu().numpy().view(_lib.SYNVOX_DTYPE).reshape(-1)
    off = np.asarray(offset, np.int64) + overlap                       # :470
    res = syntype_to_dicts(cs_rec, v, cshape, off)
    cs_seg = cs_c.cpu().numpy().view(np.uint64)
    syn_seg = torch.where(masks[0][crop] != 0, cs_c, torch.zeros_like(cs_c)).cpu().numpy().view(np.uint64)   # :476
    return res[0], res[1], res[2], res[3], res[4], cs_seg, syn_seg
