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
