"""Drop-in for ``syconn.extraction.find_object_properties_C`` (Cython module, reference plugin boundary).

Same names, argument meaning, return structures and error behaviour; the work runs in libsyk's sm_100a kernels
(``syk_find_object_properties_host`` / ``syk_map_subcell_extract_props_host``, include/syk.h)."""
from ._host import (find_object_properties_records, map_subcell_records, pairs_to_dict, records_to_dicts)


def find_object_properties(chunk):
    """syconn/extraction/find_object_properties_C.pyx:24-49 -> (rep_coords, bounding_box, sizes) dicts."""
    return records_to_dicts(find_object_properties_records(chunk))


def map_subcell_extract_props(ch, subcell_chs):
    """syconn/extraction/find_object_properties_C.pyx:112-192 ->
    ([rc, bb, size], [[rc_c..], [bb_c..], [size_c..]], [map_c..]) with map_c = {sub_id: {cell_id: count}}."""
    cell_rec, sub_recs, pair_recs = map_subcell_records(ch, subcell_chs, props_too=True)
    rc, bb, sz = records_to_dicts(cell_rec)
    sd = [records_to_dicts(r) for r in sub_recs]
    return [rc, bb, sz], [[d[0] for d in sd], [d[1] for d in sd], [d[2] for d in sd]], \
        [pairs_to_dict(p) for p in pair_recs]


def map_subcell_C(ch, subcell_chs):
    """syconn/extraction/find_object_properties_C.pyx:72-109 -> [map_c..] only."""
    _, _, pair_recs = map_subcell_records(ch, subcell_chs, props_too=False)
    return [pairs_to_dict(p) for p in pair_recs]
