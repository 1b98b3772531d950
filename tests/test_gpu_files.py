"""Whole-FILE parity of the host-side mirror of the reference API (minnow_b200.minnow / minh /
minp, codecs on the GPU through the C ABI): the files must be byte-identical to the golden
files produced by the reference's own Python twin (tests/golden, see make_golden.py) and to
the images of the oracle's writers, on the vectors of the reference's own tests
(go/minnow_test.go, go/minh/minh_test.go, go/minp/minp_test.go).  Run on the B200 box."""
import io
import struct

import numpy as np
import pytest

import minnow_b200 as mb
from minnow_b200 import minh, minnow, minp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = mb.Context(0)
    yield c
    c.close()


def _golden(golden, name):
    with open(golden(name), "rb") as f:
        return f.read()


def test_bit_int_record_file(ctx, golden):
    # go/minnow_test.go:242-268 TestBitIntRecord == python/minnow_test.py test_bit_int_record
    buf = io.BytesIO()
    wr = minnow.Create(buf, ctx)
    x1, x2, x3 = [100, 101, 102, 104], [[1024, 1024, 1024], [0, 1023, 500]], [-1000000, -500000]
    wr.IntGroup(4); b1 = wr.Data(np.array(x1, np.int64))
    wr.Header(np.array([2], np.int64))
    wr.IntGroup(3); wr.Data(np.array(x2[0], np.int64)); b3 = wr.Data(np.array(x2[1], np.int64))
    wr.IntGroup(2); b4 = wr.Data(np.array(x3, np.int64))
    wr.Close()
    assert (b1, b3, b4) == (0, 2, 3)
    assert buf.getvalue() == _golden(golden, "bit_int_record.minnow")
    rd = minnow.Open(buf.getvalue(), ctx)
    assert rd.Blocks() == 4 and [rd.DataType(b) for b in range(4)] == [minnow.IntGroup] * 4
    assert rd.Data(0).tolist() == x1 and rd.Data(1).tolist() == x2[0] and rd.Data(2).tolist() == x2[1]
    assert rd.Data(3).tolist() == x3 and rd.Header(0, "<i8").tolist() == [2] and rd.DataLen(2) == 3


def test_q_float_record_file(ctx, golden):
    # go/minnow_test.go:270-310 TestQFloatRecord / python/minnow_test.py test_q_float_record (the golden
    # file was written by the reference's Python twin): tolerance |x - y| <= dx on read-back
    lim, dx1, dx2 = (-50.0, 100.0), 1.0, 10.0
    blocks1 = [[-50, 0, 50, 49], [25, 25, 25, 25]]
    blocks2 = [[-50, 0, 50, 49, 0], [1, 2, 3, 4, 5], [0, 20, 0, 20, 0]]
    image = _golden(golden, "q_float_record.minnow")
    rd = minnow.Open(image, ctx, jitter=mb.Jitter.make(mb.JITTER_HASH, 9))
    assert struct.unpack("<ffffqq", rd.Header(0)) == (dx1, dx2, lim[0], lim[1], 2, 3)
    k = 0
    for dx, blocks in ((dx1, blocks1), (dx2, blocks2)):
        for blk in blocks:
            got = rd.Data(k)
            assert np.all(np.abs(got - np.array(blk, np.float32)) <= dx), (k, got)
            k += 1
    buf = io.BytesIO()
    wr = minnow.Create(buf, ctx)
    wr.Header(struct.pack("<ffffqq", dx1, dx2, lim[0], lim[1], len(blocks1), len(blocks2)))
    wr.FloatGroup(4, lim, dx1)
    for blk in blocks1:
        wr.Data(np.array(blk, np.float32))
    wr.FloatGroup(5, lim, dx2)
    for blk in blocks2:
        wr.Data(np.array(blk, np.float32))
    wr.Close()
    assert buf.getvalue() == image
    with pytest.raises(TypeError):                       # TypeMatch, go/group.go:43-71
        w2 = minnow.Create(io.BytesIO(), ctx); w2.IntGroup(2); w2.Data(np.zeros(2, np.float32))
    with pytest.raises(RuntimeError):                    # go/writer.go:91-92
        w3 = minnow.Create(io.BytesIO(), ctx); w3.Data(np.zeros(2, np.int64))


def test_fixed_groups_and_golden_files(ctx, golden):
    # go/minnow_test.go:191-240: the twin's int_record / group_record files read back exactly
    for name in ("int_record.minnow", "group_record.minnow", "int_groups_random.minnow"):
        image = _golden(golden, name)
        rd = minnow.Open(image, ctx)
        buf = io.BytesIO()
        wr = minnow.Create(buf, ctx)
        # re-write the file block by block through the mirror: header/group order follows file offsets
        events = [(int(o), "h", i) for i, o in enumerate(rd.header_offsets)] + \
                 [(int(o), "g", i) for i, o in enumerate(rd.group_offsets)]
        for _, kind, i in sorted(events):
            if kind == "h":
                wr.Header(rd.Header(i))
                continue
            g = rd.groups[i]
            if g.gt in minnow._FIXED:
                wr.FixedSizeGroup(g.gt, g.N)
            elif g.gt == minnow.IntGroup:
                wr.IntGroup(g.N)
            for k in range(len(g.sizes)):
                wr.Data(rd.Data(g.start_block + k))
        wr.Close()
        assert buf.getvalue() == image, name


def test_data_blocks_equals_repeated_data(ctx):
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 125, 6 * 1000).astype(np.float32)
    ids = rng.integers(-10 ** 12, 10 ** 12, 6 * 1000).astype(np.int64)
    a, b = io.BytesIO(), io.BytesIO()
    wa, wb = minnow.Create(a, ctx), minnow.Create(b, ctx)
    for w in (wa, wb):
        w.Header(b"hello")
    wa.FloatGroup(1000, (0, 125), 0.001); wb.FloatGroup(1000, (0, 125), 0.001)
    wa.DataBlocks(x)
    for k in range(6):
        wb.Data(x[k * 1000:(k + 1) * 1000])
    wa.IntGroup(1000); wb.IntGroup(1000)
    wa.DataBlocks(ids)
    for k in range(6):
        wb.Data(ids[k * 1000:(k + 1) * 1000])
    wa.Close(); wb.Close()
    assert a.getvalue() == b.getvalue()
    rd = minnow.Open(a.getvalue(), ctx)
    assert np.array_equal(np.concatenate([rd.Data(6 + k) for k in range(6)]), ids)
    assert np.all(np.abs(np.concatenate([rd.Data(k) for k in range(6)]) - x) <= 0.001 * 1.001)


MINH_NAMES = ["int64", "float32", "int", "float", "log"]
MINH_TEXT = ("Cats are the best. Don't we love them?!@#$%^&*(),.." + "..[]{};':\"|\\/-=_+`~meow meow meow")
MINH_COLS = [(0,), (9,), (10,), (11, 0, 100, 200, 1), (11, 1, 10, 14, 0.01)]
MINH_B1 = [[100, 200, 300, 400, 500], [150, 250, 350, 450, 550], [-30, -35, -25, -10, -20],
           [100, 200, 125, 150, 100], [1e10, 1e11, 1e11, 1e14, 3e13]]
MINH_B2 = [[125, 225, 325], [1750, 2750, 3750], [1000, 1000, 1000], [100, 100, 100], [1e14, 1e14, 1e14]]


def test_minh_reader_writer_file(ctx, golden):
    # go/minh/minh_test.go:10-117 TestReaderWriter; golden file written by the reference's python/minh.py
    buf = io.BytesIO()
    wr = minh.Create(buf, ctx)
    wr.Header(MINH_NAMES, MINH_TEXT, MINH_COLS)
    wr.Geometry(100.0, 10.0, 4)
    for blk in (MINH_B1, MINH_B2):
        wr.Block([np.array(blk[0], np.int64), np.array(blk[1], np.float32), np.array(blk[2], np.int64),
                  np.array(blk[3], np.float32), np.array(blk[4], np.float32)])
    wr.Close()
    assert buf.getvalue() == _golden(golden, "minh_reader_writer.minh")
    rd = minh.Open(buf.getvalue(), ctx, jitter=mb.Jitter.make(mb.JITTER_HASH, 3))
    assert rd.Names == MINH_NAMES and rd.Text == MINH_TEXT and rd.Blocks == 2 and rd.BlockLengths == [5, 3]
    assert (rd.L, rd.Boundary, rd.Cells, rd.Length) == (100.0, 10.0, 4, 8)
    ints = rd.Ints(["int64", "int"])
    assert ints["int64"].tolist() == MINH_B1[0] + MINH_B2[0] and ints["int"].tolist() == MINH_B1[2] + MINH_B2[2]
    fl = rd.Floats(["float32", "float", "log"])
    assert fl["float32"].tolist() == MINH_B1[1] + MINH_B2[1]
    assert np.all(np.abs(fl["float"] - np.array(MINH_B1[3] + MINH_B2[3], np.float32)) <= 1)             # float32sEq, tol Dx
    assert np.all(np.abs(np.log10(fl["log"]) - np.log10(np.array(MINH_B1[4] + MINH_B2[4]))) <= 0.01)    # log32sEq


@pytest.mark.parametrize("nside,file_cells,sub_cells", [(1, 1, 1), (2, 1, 1), (8, 1, 1), (10, 1, 1), (2, 1, 2), (10, 1, 2),
                                                        (10, 1, 5), (64, 2, 2)])
@pytest.mark.parametrize("periodic", [False, True])
def test_minp_file(ctx, orc, nside, file_cells, sub_cells, periodic):
    # go/minp/minp_test.go:7-73 TestVecReaderWriter shapes: the file image equals the oracle writer's
    rng = np.random.default_rng(nside + sub_cells)
    L, dx = 100.0, 0.1
    nfile = nside // file_cells
    g = np.stack(np.meshgrid(*[np.arange(nfile)] * 3, indexing="ij"), -1).transpose(2, 1, 0, 3).reshape(-1, 3)
    vec = np.mod(g * (L / nside) + rng.normal(0, 1.0, g.shape), L).astype(np.float32)
    vec[vec >= L] = 0
    hd = np.zeros(1, minp.Header)
    hd["L"], hd["NSide"], hd["NTotal"], hd["Z"], hd["Scale"] = L, nside, nside ** 3, 1, .5
    hd["OmegaM"], hd["OmegaL"], hd["H100"], hd["Epsilon"], hd["UniformMp"] = .27, .73, .7, 2, 1e10
    cell = np.array([(file_cells ** 3 - 1, file_cells, sub_cells)], minp.Cell)
    raw = bytes(range(130))
    buf = io.BytesIO()
    wr = minp.Create(buf, ctx)
    wr.Header(hd, raw, cell, dx, periodic)
    wr.Vectors(vec)
    wr.Close()
    want = orc.minp_write(hd.view(orc.MINP_HEADER), raw, cell.view(orc.MINP_CELL), dx, periodic, vec)
    assert buf.getvalue() == want
    for mode in (mb.JITTER_CENTER, mb.JITTER_HASH):
        rd = minp.Open(buf.getvalue(), ctx, jitter=mb.Jitter.make(mode, 31))
        assert rd.Header == hd[0] and rd.RawHeader == raw and rd.Periodic == periodic and rd.Dx == dx
        got = rd.Vectors()
        _, _, _, _, ref = orc.minp_read(want, mode, 31)
        if nside > 1 or periodic:
            assert got.tobytes() == ref.tobytes()
        ocell = np.zeros(1, orc.MINP_CELL); ocell[0] = tuple(cell[0])
        assert np.array_equal(rd.IDs(), orc.minp_ids(nside, ocell))
        assert rd.N() == sub_cells ** 3
