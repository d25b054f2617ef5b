"""Dump I/O boundary: the writer produces the record structure utils_dumpfiles.f90 defines and the reader
recovers every array and header entry; a dump written from one state restarts the hot path bit-identically (host side, no GPU)."""
import os
import struct
import tempfile

import numpy as np
import pytest

from phantom_b200 import setups, dumpio
from oraclelib import Oracle


def test_record_structure_matches_utils_dumpfiles():
    part = setups.setup_orstang(nx=12)
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "orstang_00000")
        fileid = dumpio.write_dump(fn, part, time=0.25)
        raw = open(fn, "rb").read()
        # first record: int1, r1 (default real*8), int2, iversion, int3 between 4-byte markers (open_dumpfile_w)
        assert struct.unpack("<i", raw[:4])[0] == 24 and struct.unpack("<i", raw[28:32])[0] == 24
        i1, r1, i2, iv, i3 = struct.unpack("<idiii", raw[4:28])
        assert (i1, i2, iv, i3) == (60769, 60878, 1, 690706) and r1 == 60878.0
        # second record: fileid, character(len=100), 'FT:Phantom' + '(mhd' for an MHD dump (fileident, readwrite_dumps_common.f90:32-69)
        assert struct.unpack("<i", raw[32:36])[0] == 100
        assert raw[36:46] == b"FT:Phantom" and b"(mhd+clean" in raw[36:136] and fileid.startswith("FT:Phantom")
        dd = dumpio.read_dump(fn)
    h = dd["header"]
    assert h["nparttot"] == part.npart and h["npartoftype"][0] == part.npart and h["time"] == 0.25 and h["hfact"] == part.params.hfact
    assert len(dd["blocks"]) == 4 and len(dd["blocks"][1]) == 0                 # narraylengths = 4 with MHD, empty sink block
    g, m = dd["blocks"][0], dd["blocks"][3]
    assert g["x"].dtype == np.float64 and g["h"].dtype == np.float32 and g["itype"].dtype == np.int8
    assert np.array_equal(g["x"], part.xyzh[:, 0]) and np.array_equal(g["vy"], part.vxyzu[:, 1]) and np.array_equal(g["u"], part.vxyzu[:, 3])
    assert set(("Bx", "By", "Bz", "psi")) <= set(m)


def test_wrong_endian_and_truncated_files_are_rejected():
    part = setups.setup_turb(nx=6)
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "turb_00000")
        dumpio.write_dump(fn, part)
        raw = bytearray(open(fn, "rb").read())
        bad = bytes(raw[:4]) + struct.pack(">i", 60769) + bytes(raw[8:])
        open(fn + "_be", "wb").write(bad)
        with pytest.raises(dumpio.DumpFormatError):
            dumpio.read_dump(fn + "_be")
        open(fn + "_cut", "wb").write(bytes(raw[:len(raw) // 2]))
        with pytest.raises((dumpio.DumpFormatError, EOFError, struct.error)):
            dumpio.read_dump(fn + "_cut")


def test_restart_from_dump_reproduces_derivs():
    # x, v, u are stored as default real (exact); h is real*4 in a full dump, as in the reference: restart from a dump whose h
    # is already representable in real*4 and require bit-identical derivatives
    part, _ = setups.setup_test_derivs(nx=10, lattice="random")
    part.xyzh[:, 3] = part.xyzh[:, 3].astype(np.float32).astype(np.float64)
    part.alphaind[:, 0] = 0.5
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "test_00000")
        dumpio.write_dump(fn, part)
        back = dumpio.particles_from_dump(dumpio.read_dump(fn), part.params.copy())
    assert np.array_equal(back.xyzh, part.xyzh) and np.array_equal(back.vxyzu, part.vxyzu) and np.array_equal(back.alphaind[:, 0], part.alphaind[:, 0])
    a, b = part.copy(), back.copy()
    Oracle(a.params).derivs(a)
    Oracle(b.params).derivs(b)
    assert np.array_equal(a.fxyzu, b.fxyzu) and np.array_equal(a.xyzh, b.xyzh)


def test_multi_block_dump_of_an_mpi_run_reads_back():
    # readwrite_dumps.f90:170-172, :600-665: nblocks processes each write {block headers, arrays}; the reader concatenates the pieces
    part = setups.setup_orstang(nx=10)
    with tempfile.TemporaryDirectory() as d:
        one, three = os.path.join(d, "one_00000"), os.path.join(d, "three_00000")
        dumpio.write_dump(one, part, time=0.5)
        dumpio.write_dump(three, part, time=0.5, nblocks=3)
        assert os.path.getsize(three) > os.path.getsize(one)
        a, b = dumpio.read_dump(one), dumpio.read_dump(three)
    assert b["nblocks"] == 3 and a["nblocks"] == 1 and len(b["blocks"]) == len(a["blocks"]) == 4
    for ka, kb in zip(a["blocks"], b["blocks"]):
        assert set(ka) == set(kb)
        for tag in ka:
            assert np.array_equal(ka[tag], kb[tag]), tag
    pa = dumpio.particles_from_dump(b, part.params.copy())
    assert np.array_equal(pa.xyzh[:, :3], part.xyzh[:, :3]) and np.array_equal(pa.vxyzu, part.vxyzu)
