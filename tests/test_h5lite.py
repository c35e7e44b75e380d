"""SURVEY.md section 8f rank 4 -- the reference's on-disk HDF5 layouts (analyses/scripts/julia/bson_to_hdf.jl:18-71,
fit_matfac.jl:60-117, script_util.jl:147-180) through h5lite, the NumPy restatement of the HDF5 subset libhdf5 writes by
default (the image has no HDF5 library).

Anchor: tests/golden/matlab73_testdouble.mat is a file libhdf5 itself produced (SciPy's test data
scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD-3: a MATLAB v7.3 file = HDF5 behind a 512-byte user block); SciPy's
own test of the same variable in the other MAT formats expects ``pi/4 * arange(9)`` as a 1 x 9 matrix
(scipy/io/matlab/tests/test_mio.py, case "double").  The reader is pinned to that file; the writer is then checked (a)
byte for byte against the messages libhdf5 wrote for the same datatype / dataspace, (b) by an independent structural walk
of its output that enforces the invariants libhdf5 checks when it opens a file, (c) through the reader."""
import os
import struct

import numpy as np
import pytest

import pathmatfac_b200 as P
from pathmatfac_b200 import h5lite

GOLD = os.path.join(os.path.dirname(__file__), "golden", "matlab73_testdouble.mat")
UNDEF = 0xFFFFFFFFFFFFFFFF


def test_reader_is_pinned_to_a_libhdf5_file():
    f = h5lite.File(GOLD)
    assert f.base == 512 and f.eof + 0 == os.path.getsize(GOLD)        # user block; end-of-file address
    assert (f.leaf_k, f.internal_k) == (4, 16)
    assert f.keys() == ["testdouble"] and f.visit() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")            # MATLAB 1 x 9, dimensions reversed on disk
    assert d.attrs == {"MATLAB_class": "double"}
    x = d.read()                                                         # julia=True: column-major convention, 1 x 9
    assert x.shape == (1, 9) and x.dtype == np.float64
    assert np.array_equal(x, (np.pi / 4 * np.arange(9, dtype=float)).reshape(1, 9))
    assert np.array_equal(h5lite.read(GOLD, "testdouble", julia=False), x.T)
    with pytest.raises(KeyError):
        f["nothing/here"]
    with pytest.raises(h5lite.H5Error):
        h5lite.File(b"not an hdf5 file" * 100)


def test_writer_emits_the_messages_libhdf5_wrote():
    """Datatype (IEEE float64, little endian) and dataspace (9 x 1) messages of the anchor file, byte for byte."""
    f = h5lite.File(GOLD)
    msgs = dict((t, b) for t, b in f._messages(f._members["testdouble"]) if t in (1, 3))
    assert h5lite._dtype_message(np.dtype("<f8")) == msgs[3][:20]
    space = struct.pack("<BBBBI", 1, 2, 0, 0, 0) + struct.pack("<QQ", 9, 1)
    assert space == msgs[1]
    w = h5lite.Writer().write("testdouble", (np.pi / 4 * np.arange(9)).reshape(1, 9))
    g = h5lite.File(w.tobytes())
    mine = dict((t, b) for t, b in g._messages(g._members["testdouble"]) if t in (1, 3))
    assert mine[1] == msgs[1] and mine[3][:20] == msgs[3][:20]

    def flags(file, addr):                                   # message type -> flag byte, first header block
        hsize = struct.unpack_from("<I", file.at(addr, 16), 8)[0]
        blk, off, out = file.at(addr + 16, hsize), 0, {}
        while off + 8 <= hsize:
            t, sz, fl = struct.unpack_from("<HHB", blk, off)
            out.setdefault(t, fl)
            off += 8 + sz
        return out
    theirs, ours = flags(f, f._members["testdouble"]), flags(g, g._members["testdouble"])
    assert all(ours[t] == theirs[t] for t in (1, 3, 5))      # dataspace 0, datatype and fill value "constant"
    assert np.array_equal(g.read("testdouble"), f.read("testdouble"))
    # the local heap of the anchor: the empty string at offset 0, names null-terminated and 8-aligned, the free list
    # closed by a block whose "next" is 1 -- the conventions the writer follows
    heap = f.at(0x60, 32)
    size, free, daddr = struct.unpack_from("<QQQ", heap, 8)
    seg = f.at(daddr, size)
    assert seg[:8] == bytes(8) and seg[8:19] == b"testdouble\0" and free == 24
    assert struct.unpack_from("<QQ", seg, free) == (1, size - free)


# ---- an independent structural walk: what libhdf5 verifies (or relies on) when it opens the file -------------------

def _cstr(seg, off):
    return seg[off:seg.index(b"\0", off)]


def walk_structure(buf):
    """Returns the number of (groups, datasets, global-heap collections) visited; asserts on every violated invariant."""
    assert buf[:8] == b"\x89HDF\r\n\x1a\n"
    assert tuple(buf[8:16]) == (0, 0, 0, 0, 0, 8, 8, 0)
    leaf_k, int_k, flags = struct.unpack_from("<HHI", buf, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", buf, 24)
    assert (base, free, drv, flags) == (0, UNDEF, UNDEF, 0) and eof == len(buf)
    name_off, root, cache, _ = struct.unpack_from("<QQII", buf, 56)
    bt, hp = struct.unpack_from("<QQ", buf, 80)
    assert cache == 1 and name_off == 0
    counts = [0, 0, 0]
    seen_gcol = set()

    def header(addr):
        assert addr % 8 == 0 and addr + 16 <= eof
        ver, _r, nmsg, refs, hsize = struct.unpack_from("<BBHII", buf, addr)
        assert ver == 1 and refs == 1 and hsize % 8 == 0 and addr + 16 + hsize <= eof
        off, msgs = addr + 16, []
        for _ in range(nmsg):
            t, sz, fl = struct.unpack_from("<HHB", buf, off)
            assert sz % 8 == 0
            msgs.append((t, buf[off + 8:off + 8 + sz]))
            off += 8 + sz
        assert off == addr + 16 + hsize                                  # the messages fill the header exactly
        return msgs

    def group(oh, bt_cached, hp_cached):
        counts[0] += 1
        msgs = header(oh)
        assert [t for t, _ in msgs] == [0x11]
        btree, heap = struct.unpack_from("<QQ", msgs[0][1], 0)
        assert (btree, heap) == (bt_cached, hp_cached)                   # the symbol-table entry caches them
        assert buf[heap:heap + 4] == b"HEAP" and buf[heap + 4] == 0
        size, free_off, daddr = struct.unpack_from("<QQQ", buf, heap + 8)
        seg = buf[daddr:daddr + size]
        assert len(seg) == size and seg[:8] == bytes(8)
        nxt, fsz = struct.unpack_from("<QQ", seg, free_off)
        assert nxt == 1 and fsz >= 16 and free_off + fsz == size         # one free block, closing the segment
        names = []

        def node(a, level, lo, hi):
            """Every name below node `a` is in (lo, hi] and hi is the largest one."""
            assert buf[a:a + 4] == b"TREE" and buf[a + 4] == 0 and buf[a + 5] == level
            used = struct.unpack_from("<H", buf, a + 6)[0]
            assert used <= 2 * int_k and a + 24 + 8 + 32 * int_k <= eof  # the node owns its full size
            keys = [struct.unpack_from("<Q", buf, a + 24 + 16 * i)[0] for i in range(used + 1)]
            kids = [struct.unpack_from("<Q", buf, a + 32 + 16 * i)[0] for i in range(used)]
            knames = [_cstr(seg, k) for k in keys]
            assert knames[0] == lo and (used == 0 or knames[-1] == hi)
            assert all(x < y for x, y in zip(knames, knames[1:]))
            for i, c in enumerate(kids):
                if level > 0:
                    node(c, level - 1, knames[i], knames[i + 1])
                    continue
                assert buf[c:c + 4] == b"SNOD" and buf[c + 4] == 1
                nsym = struct.unpack_from("<H", buf, c + 6)[0]
                assert 1 <= nsym <= 2 * leaf_k and c + 8 + 80 * leaf_k <= eof
                ents = [struct.unpack_from("<QQII", buf, c + 8 + 40 * j) + struct.unpack_from("<QQ", buf, c + 32 + 40 * j)
                        for j in range(nsym)]
                en = [_cstr(seg, e[0]) for e in ents]
                assert all(x < y for x, y in zip(en, en[1:]))            # sorted inside the node
                assert knames[i] < en[0] and en[-1] == knames[i + 1]     # bracketed by the B-tree keys
                for e, nm in zip(ents, en):
                    assert e[0] % 8 == 0
                    names.append(nm)
                    if e[2] == 1:
                        group(e[1], e[4], e[5])
                    else:
                        assert e[2] == 0
                        dataset(e[1])
            return used

        top_level = buf[btree + 5]
        used = struct.unpack_from("<H", buf, btree + 6)[0]
        last = b""
        if used:
            last = _cstr(seg, struct.unpack_from("<Q", buf, btree + 24 + 16 * used)[0])
        node(btree, top_level, b"", last)
        assert names == sorted(names) and len(set(names)) == len(names)

    def dataset(oh):
        counts[1] += 1
        msgs = dict(header(oh))
        assert set(msgs) == {1, 3, 5, 8}
        sp = msgs[1]
        rank = sp[1]
        assert sp[0] == 1 and sp[2] == 0
        dims = struct.unpack_from(f"<{rank}Q", sp, 8)
        dt = msgs[3]
        cls, ver, esize = dt[0] & 15, dt[0] >> 4, struct.unpack_from("<I", dt, 4)[0]
        assert ver == 1 and cls in (0, 1, 9)
        assert tuple(msgs[5][:8]) == (2, 2, 2, 1, 0, 0, 0, 0)
        lver, lcls, addr, size = struct.unpack_from("<BBQQ", msgs[8], 0)
        n = int(np.prod(dims, dtype=np.int64)) if rank else 1
        assert (lver, lcls) == (3, 1) and size == n * esize
        assert (addr == UNDEF and size == 0) or (addr % 8 == 0 and addr + size <= eof)
        if cls == 9:
            assert esize == 16 and dt[1] == 0x01 and dt[2] == 0x01     # string, null-terminated, UTF-8
            for i in range(n):
                ln, gaddr, idx = struct.unpack_from("<IQI", buf, addr + 16 * i)
                if ln == 0:
                    assert (gaddr, idx) == (0, 0)
                    continue
                if gaddr not in seen_gcol:
                    seen_gcol.add(gaddr)
                    gcol(gaddr)
                assert 1 <= idx < 65536

    def gcol(a):
        counts[2] += 1
        assert buf[a:a + 4] == b"GCOL" and buf[a + 4] == 1 and a % 8 == 0
        size = struct.unpack_from("<Q", buf, a + 8)[0]
        assert size >= 4096 and a + size <= eof
        off, expect = a + 16, 1
        while True:
            idx, refs, _r, osz = struct.unpack_from("<HHIQ", buf, off)
            if idx == 0:
                assert off + osz == a + size and osz >= 16               # the free-space object closes the collection
                break
            assert idx == expect
            expect += 1
            off += 16 + (osz + 7) // 8 * 8

    group(root, bt, hp)
    return tuple(counts)


def _example_writer(rng):
    w = h5lite.Writer()
    arrays = {
        "X": rng.standard_normal((3, 7)).astype(np.float32),
        "Y": rng.standard_normal((3, 5)),
        "data_idx": np.arange(1, 8, dtype=np.int64),
        "small/int32": np.arange(-3, 3, dtype=np.int32).reshape(2, 3),
        "small/uint8": np.arange(6, dtype=np.uint8),
        "small/f16": np.linspace(0, 1, 5).astype(np.float16),
        "small/empty": np.zeros((0, 4), np.float32),
        "small/cube": rng.standard_normal((2, 3, 4)).astype(np.float32),
        "scalar": np.float64(3.5),
    }
    for i in range(70):                                   # > 32 symbol-table nodes: a two-level B-tree
        for j in range(4):
            arrays[f"many/values_{i}_{j}"] = rng.standard_normal((2, i % 5 + 1)).astype(np.float32)
    for k, v in arrays.items():
        w.write(k, v)
    strings = {"ids": [f"feature_{i}_é漢" for i in range(9000)] + ["", "x"],
               "few": ["a", "bb", ""], "grid": np.array([["p", "q", "r"], ["s", "t", "u"]], dtype=object)}
    for k, v in strings.items():
        w.write("str/" + k, v)
    return w, arrays, strings


def test_writer_structure_and_round_trip(tmp_path):
    w, arrays, strings = _example_writer(np.random.default_rng(5))
    path = w.save(tmp_path / "x.h5")
    buf = open(path, "rb").read()
    n_groups, n_datasets, n_gcol = walk_structure(buf)
    assert n_groups == 4 and n_datasets == len(arrays) + len(strings) and n_gcol == 3 + 1 + 1   # 9002 strings: 3 collections
    f = h5lite.File(path)
    assert f.keys() == sorted(["X", "Y", "data_idx", "small", "scalar", "many", "str"])
    assert len(f["many"].keys()) == 280
    for k, v in arrays.items():
        got = f.read(k)
        assert got.dtype == v.dtype and got.shape == np.shape(v) and np.array_equal(got, v), k
        assert f[k].shape == np.shape(v)[::-1]                              # dimensions reversed on disk, as HDF5.jl stores them
    for k, v in strings.items():
        assert np.array_equal(f.read("str/" + k), np.array(v, dtype=object)), k
    assert np.array_equal(f.read("small/int32", julia=False), arrays["small/int32"].T)
    with pytest.raises(h5lite.H5Error):
        h5lite.Writer().write("a", np.zeros(2)).write("a", np.zeros(2))        # a name is written once
    with pytest.raises(h5lite.H5Error):
        h5lite.Writer().write("a", np.zeros(2, dtype=complex))
    assert walk_structure(h5lite.Writer().tobytes()) == (1, 0, 0)               # an empty file is a root group with no members


def _chunked_file(data, chunk, deflate, shuffle):
    """A chunked (optionally shuffled + deflated) 2-D float32 dataset assembled by hand from the format specification
    (layout message v3 class 2, filter pipeline v1, v1 chunk B-tree): what `h5py.create_dataset(chunks=..., compression=...)`
    produces.  Self-consistency of the reader only -- no library-written chunked file exists in the image."""
    import zlib
    w = h5lite.Writer()
    w.buf = bytearray(96)
    keys = []
    for r0 in range(0, data.shape[0], chunk[0]):
        for c0 in range(0, data.shape[1], chunk[1]):
            blk = np.zeros(chunk, np.float32)
            part = data[r0:r0 + chunk[0], c0:c0 + chunk[1]]
            blk[:part.shape[0], :part.shape[1]] = part
            raw = blk.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(-1, 4).T.tobytes()
            if deflate:
                raw = zlib.compress(raw, 4)
            keys.append((len(raw), (r0, c0), w._alloc(raw)))
    node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
    for size, (r0, c0), addr in keys:
        node += struct.pack("<IIQQQ", size, 0, r0, c0, 0) + struct.pack("<Q", addr)
    node += struct.pack("<IIQQQ", 0, 0, data.shape[0], data.shape[1], 0)
    bt = w._alloc(node)
    filters = []
    if shuffle:
        filters.append(struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 4, 0))
    if deflate:
        filters.append(struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 4, 0))
    msgs = [h5lite._message(1, struct.pack("<BBBBI", 1, 2, 0, 0, 0) + struct.pack("<QQ", *data.shape)),
            h5lite._message(3, h5lite._dtype_message(np.dtype("<f4"))),
            h5lite._message(8, struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<III", chunk[0], chunk[1], 4))]
    if filters:
        msgs.append(h5lite._message(0xB, struct.pack("<BB", 1, len(filters)) + bytes(6) + b"".join(filters)))
    ds = w._alloc(h5lite._object_header(msgs))
    w._dataset = lambda arr: ds
    w.tree = {"data": np.zeros(1)}
    oh, gbt, hp = w._group(w.tree)
    w.buf += bytes(-len(w.buf) % 8)
    sb = h5lite.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(w.buf), UNDEF) + struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", gbt, hp)
    w.buf[:96] = sb
    return bytes(w.buf)


@pytest.mark.parametrize("deflate,shuffle", [(False, False), (True, False), (True, True)])
def test_reader_chunked_layouts(deflate, shuffle):
    data = np.random.default_rng(2).standard_normal((37, 21)).astype(np.float32)
    f = h5lite.File(_chunked_file(data, (16, 8), deflate, shuffle))
    assert np.array_equal(f.read("data", julia=False), data)


# ---- the reference's layouts --------------------------------------------------------------------------------------

def _model(rng, fsard=False):
    M, N = 12, 9
    D = rng.standard_normal((M, N)).astype(np.float32)
    views = ["methylation"] * 4 + ["mrnaseq"] * 5
    batch = {"methylation": ["b1"] * 6 + ["b2"] * 6, "mrnaseq": ["x"] * 4 + ["y"] * 4 + ["z"] * 4}
    kw = {}
    if fsard:
        fids = [f"g{j}" for j in range(N)]
        kw = dict(feature_ids=fids, feature_sets_dict={"methylation": [["g0", "g1"], ["g2", "g3", "g1"]],
                                                       "mrnaseq": [["g4", "g5", "g6"], ["g7", "g8"]]}, Y_fsard=True)
    model = P.PathMatFacModel(D.copy(), K=3, feature_views=views, batch_dict=batch,
                              sample_conditions=["c1"] * 6 + ["c2"] * 6, lambda_X_l2=1.0, **kw)
    model.matfac.X[...] = rng.standard_normal(model.matfac.X.shape)
    model.matfac.Y[...] = rng.standard_normal(model.matfac.Y.shape)
    return model


def test_write_model_to_hdf_has_the_datasets_of_bson_to_hdf(tmp_path):
    """bson_to_hdf.jl:18-71: dataset names, Julia (dimension-reversed) storage, 1-based indices."""
    model = _model(np.random.default_rng(3))
    path = P.write_model_to_hdf(tmp_path / "model.hdf", model)
    walk_structure(open(path, "rb").read())
    f = h5lite.File(path)
    expect = ["X", "Y", "data_idx", "feature_ids", "feature_views", "logsigma", "mu", "sample_conditions", "sample_ids"]
    expect += [f"logdelta/{n}_{i}" for i in (1, 2) for n in ("values", "col_range")]
    expect += [f"theta/{n}_{i}" for i in (1, 2) for n in ("values", "col_range", "batch_ids")]
    assert f.visit() == sorted(expect)
    K, M, N = 3, 12, 9
    assert f["X"].shape == (M, K) and f["Y"].shape == (N, K)              # Julia K x M / K x N, reversed on disk
    got = P.read_model_hdf(path)
    assert np.array_equal(got["X"], model.matfac.X) and np.array_equal(got["Y"], model.matfac.Y)
    assert list(got["feature_views"]) == list(model.feature_views) and list(got["sample_conditions"]) == ["c1"] * 6 + ["c2"] * 6
    assert np.array_equal(got["data_idx"], np.asarray(model.data_idx) + 1)
    assert np.array_equal(got["logdelta/col_range_1"], np.arange(1, 5)) and np.array_equal(got["theta/col_range_2"], np.arange(5, 10))
    assert list(got["theta/batch_ids_1"]) == ["b1", "b2"] and list(got["theta/batch_ids_2"]) == ["x", "y", "z"]
    assert got["theta/values_2"].shape == (3, 5) and got["logdelta/values_1"].shape == (2, 4)
    arr = P.model_arrays(model)
    assert sorted(arr) == sorted(expect)
    for k, v in arr.items():
        assert np.array_equal(got[k], v), k


def test_write_model_to_hdf_feature_set_ard(tmp_path):
    model = _model(np.random.default_rng(4), fsard=True)
    from pathmatfac_b200.regularizers import FeatureSetARDReg
    assert isinstance(model.matfac.Y_reg, FeatureSetARDReg)
    got = P.read_model_hdf(P.write_model_to_hdf(tmp_path / "m.hdf", model))
    for i, (A, S) in enumerate(zip(model.matfac.Y_reg.A, model.matfac.Y_reg.S)):
        assert np.array_equal(got[f"fsard/A/{i + 1}"], A)
        assert got[f"fsard/S/{i + 1}"].dtype == np.float32 and np.array_equal(got[f"fsard/S/{i + 1}"], np.asarray(S.todense()))


def test_omic_hdf_layout_round_trip(tmp_path):
    """fit_matfac.jl:60-117 / script_util.jl:147-180: omic_data/*, barcodes/*, the transformed output."""
    rng = np.random.default_rng(7)
    M = 6
    assays = ["mutation"] * 3 + ["methylation"] * 2 + ["mrnaseq"] * 4 + ["cna"] * 2
    genes = [f"GENE{j}" for j in range(len(assays))]
    data = rng.standard_normal((M, len(assays)))
    data[1, 4] = np.nan
    inst = [f"TCGA-XX-{i:04d}" for i in range(M)]
    groups = ["BRCA"] * 3 + ["LUAD"] * 3
    bfeat = ["mutation", "methylation", "mrnaseq", "cna"]
    barcodes = np.array([[f"TCGA-XX-{i:04d}-01A-{a[:2]}-A{(i + k) % 2}-0{k}" if (i + k) % 5 else "" for k, a in enumerate(bfeat)]
                         for i in range(M)], dtype=object)
    path = P.save_omic_data(tmp_path / "omic.hdf", assays, genes, inst, groups, data, barcodes=barcodes, barcode_features=bfeat)
    walk_structure(open(path, "rb").read())
    D, sid, cond, fg, fa = P.load_omic_data(path, ["mrnaseq", "methylation", "protein"])
    assert fa == ["methylation"] * 2 + ["mrnaseq"] * 4 and fg == genes[3:9]
    assert D.shape == (M, 6) and np.array_equal(D, data[:, 3:9], equal_nan=True) and sid == inst and cond == groups
    bd = P.load_batches(path, ["mutation", "mrnaseq", "methylation"])
    assert sorted(bd) == ["methylation", "mrnaseq"]                      # BATCHED_ASSAYS only (script_util.jl:25)
    assert bd["mrnaseq"][1] == "A1-02" and bd["methylation"][4] == ""     # last two barcode terms; "" stays ""
    assert P.load_batches(path, ["mutation", "cna"]) is None
    model = P.PathMatFacModel(D.astype(np.float32), K=2, feature_views=fa, feature_ids=fg, sample_ids=sid,
                              sample_conditions=cond, batch_dict=bd)
    assert model.matfac.X.shape == (2, M)
    out = P.save_transformed(model.matfac.X, sid, cond, np.arange(M) % 2, tmp_path / "t.hdf")
    f = h5lite.File(out)
    assert f.visit() == ["X", "instance_groups", "instances", "target"]
    assert np.array_equal(f.read("X"), model.matfac.X) and list(f.read("instances")) == sid
    assert np.array_equal(f.read("target"), np.arange(M) % 2)


def test_simulated_problem_in_the_studys_input_layout(tmp_path):
    """simulate.export_problem_hdf -> the file fit_matfac.jl's load_omic_data reads (fit_matfac.jl:60-82)."""
    from pathmatfac_b200.simulate import export_problem_hdf, simulate_problem
    blocks = [("mutation", "bernoulli", 6), ("methylation", "normal", 10), ("mrnaseq", "normal", 8)]
    model = simulate_problem(30, blocks=blocks, K=3, seed=5, missing=0.2)
    path = export_problem_hdf(model, str(tmp_path / "omic.hdf"))
    walk_structure(open(path, "rb").read())
    D, sid, cond, genes, assays = P.load_omic_data(path, ["methylation", "mrnaseq"])
    keep = [j for j, v in enumerate(model.feature_views) if v in ("methylation", "mrnaseq")]
    assert assays == [model.feature_views[j] for j in keep] and genes == [str(model.feature_ids[j]) for j in keep]
    assert np.array_equal(D, np.asarray(model.data)[:, keep], equal_nan=True) and np.isnan(D).any()
    assert sid == [str(x) for x in model.sample_ids] and len(cond) == 30
