"""Row f3: storage writers.  The lz4 block codec of libsyk (host code) against an independent pure-Python decoder that
follows the published format, and the AttributeDict / VoxelStorageDyn pickles written from reduced records read back with
a re-statement of the reference's readers (oracle/storage_ref.py <- syconn/backend/storage.py:26-93,208-421)."""
import os

import numpy as np
import pytest

from oracle import oracle, storage_ref
from syconn_b200.chunked import ChunkPlan, reduce_pairs, reduce_records
from syconn_b200.handler import compression
from syconn_b200.proc import sd_proc
from syconn_b200.synth import synth_labels


def _samples():
    rng = np.random.default_rng(0)
    yield b""
    yield b"a"
    yield b"abcdefghijkl"                      # 12 bytes: below MFLIMIT + 1, literals only
    yield b"a" * 20
    yield b"a" * 100000                        # long run: overlapping match, multi-byte lengths
    yield bytes(rng.integers(0, 256, 5000, dtype=np.uint8))           # incompressible: long literal run
    yield bytes(rng.integers(0, 4, 70000, dtype=np.uint8))            # many short matches
    yield np.arange(0, 40000, dtype=np.int64).tobytes()               # structured int64 (bounding boxes look like this)
    yield (bytes(rng.integers(0, 256, 300, dtype=np.uint8)) + b"\0" * 70000) * 2   # offsets near / beyond 65535
    bb = rng.integers(0, 2048, (37, 2, 3)).astype(np.int64)
    yield bb.tobytes()


def test_lz4_block_roundtrip_and_independent_decoder():
    for data in _samples():
        s = compression.compress(data)
        assert int.from_bytes(s[:4], "little") == len(data)
        assert compression.decompress(s) == data                              # our decoder
        assert storage_ref.lz4_block_decode(s[4:], len(data)) == data         # spec decoder (pure Python)
        if len(data) > 1000 and len(set(data)) == 1:
            assert len(s) < len(data) // 100                                  # it does compress


def test_lz4_known_answer_vectors():
    """Hand-assembled blocks from the format description (token | literals | offset | lengths): decoder only."""
    # 1 literal 'a', match offset 1 length 14 (0xA + 4), then 5 last literals
    blk = bytes([0x1A, ord("a"), 0x01, 0x00, 0x50]) + b"aaaaa"
    assert compression.decompress((20).to_bytes(4, "little") + blk) == b"a" * 20
    assert storage_ref.lz4_block_decode(blk, 20) == b"a" * 20
    # literal length 15 + 3 (extension byte), no match
    lit = bytes(range(18))
    blk = bytes([0xF0, 3]) + lit
    assert compression.decompress((18).to_bytes(4, "little") + blk) == lit
    # match length 4 + 15 + 255 + 2 with offset 4 ("abcd" repeated), 5 trailing literals
    blk = bytes([0x4F]) + b"abcd" + bytes([4, 0, 255, 2]) + bytes([0x50]) + b"vwxyz"
    want = b"abcd" + (b"abcd" * 69) + b"vwxyz"
    assert len(want) == 4 + 276 + 5
    assert compression.decompress(len(want).to_bytes(4, "little") + blk) == want
    assert storage_ref.lz4_block_decode(blk, len(want)) == want
    with pytest.raises(compression.LZ4BlockError):
        compression.decompress((20).to_bytes(4, "little") + bytes([0x1A, ord("a"), 0x05, 0x00]))   # offset before start


def test_array_string_lists():
    a = np.arange(24, dtype=np.int64).reshape(4, 2, 3)
    lst = compression.arrtolz4string_list(a)
    assert isinstance(lst, list) and len(lst) == 1
    assert np.array_equal(compression.lz4string_listtoarr(lst, dtype=np.int64, shape=(-1, 2, 3)), a)
    assert np.array_equal(storage_ref.lz4string_listtoarr(lst, np.int64, (-1, 2, 3)), a)
    assert compression.arrtolz4string_list(np.zeros((0, 2, 3), np.int64)) == [b""]
    assert compression.lz4string_listtoarr([b""], dtype=np.int64).shape == (0,)


def test_writers_read_back_like_the_reference(tmp_path):
    """reduced records of a chunked volume -> attr_dict.pkl / voxel.pkl per storage folder -> read back."""
    vol = synth_labels((48, 40, 36), pitch=(9, 8, 7), warp_amp=2, seed=4)
    sub = synth_labels((48, 40, 36), pitch=(5, 4, 4), seed=4, kind=1, density16=5)
    ids_scale = np.uint64(997)                       # spread the ids over several storage folders (ix // 1000 % n)
    vol, sub = vol * ids_scale, sub * ids_scale
    plan = ChunkPlan(vol.shape, (16, 16, 16))
    logs, pair_rows, acc, mapacc = [], [], oracle.new_prop_acc(), {}
    for s in range(len(plan)):
        off, size = plan.offsets[s], plan.sizes[s]
        sl = tuple(slice(off[i], off[i] + size[i]) for i in range(3))
        cp, sp, md = oracle.map_subcell_extract_props(vol[sl], sub[sl][None])
        r = sd_proc.prop_dicts_to_records([sp[0][0], sp[1][0], sp[2][0]], chunk_seq=s)
        for f in ("bb_min", "bb_max", "rep"):
            r[f] += np.array(off, np.int32)
        logs.append(r)
        pair_rows.append(sd_proc.map_dict_to_pairs(md[0]))
        oracle.merge_prop_dicts([acc, [sp[0][0], sp[1][0], sp[2][0]]], offset=np.array(off))
        oracle.merge_map_dicts([mapacc, md[0]])
    red = reduce_records(np.concatenate(logs))
    mapping = sd_proc.reduced_to_map_dict(reduce_pairs(np.concatenate(pair_rows)))
    min_vx = 6
    folders = sd_proc.write_segmentation_objects(str(tmp_path / "mi_0"), red, mapping=mapping, min_obj_vx=min_vx,
                                                 n_folders_fs=100, voxeldata_path="/kd/mi")
    assert len(folders) > 1
    rc, bb, sz = acc
    seen = set()
    for folder in folders:
        attr = storage_ref.read_attr_dict(os.path.join(folder, "attr_dict.pkl"))
        bbs, sizes, reps, meta = storage_ref.read_voxel_dyn(os.path.join(folder, "voxel.pkl"))
        assert meta == {"voxeldata_path": "/kd/mi"}
        assert set(attr) == set(bbs) == set(sizes) == set(reps)
        for k, a in attr.items():
            assert folder.endswith(storage_ref.subfold_from_ix(k, 100))
            want_bbs = np.array(bb[k])                                   # per-chunk boxes in chunk order (sd_proc.py:940)
            assert sz[k] >= min_vx and a["size"] == sz[k] == sizes[k]
            assert a["rep_coord"].dtype == np.int32 and a["rep_coord"].tolist() == list(rc[k]) == reps[k].tolist()
            assert np.array_equal(a["bounding_box"], [want_bbs[:, 0].min(axis=0), want_bbs[:, 1].max(axis=0)])
            assert bbs[k].dtype == np.int64 and np.array_equal(bbs[k], want_bbs)
            m = mapacc.get(k, {})
            assert dict(zip(a["mapping_ids"], a["mapping_ratios"])) == {c: n / sz[k] for c, n in m.items()}
            seen.add(k)
    assert seen == {k for k in sz if sz[k] >= min_vx}
    # the product's own classes read the files too (what a SegmentationObject would do)
    from syconn_b200.backend.storage import AttributeDict, VoxelStorageDyn
    ad = AttributeDict(os.path.join(folders[0], "attr_dict.pkl"))
    vd = VoxelStorageDyn(os.path.join(folders[0], "voxel"), voxel_mode=False)
    k = next(iter(ad.keys()))
    assert vd.object_size(k) == ad[k]["size"] and np.array_equal(vd.get_boundingdata(k), np.array(bb[k]))
    assert sorted(vd.keys()) == sorted(ad.keys())
