"""Reader / writer of Phantom's native dump format: the data boundary either side of the hot path.

Format (src/main/utils_dumpfiles.f90, readwrite_dumps.f90:85-330; all records are Fortran sequential unformatted records with
4-byte length markers, `-frecord-marker=4`):

    record   int1=060769 (i4), r1=real(int2) (default real, 4 or 8 bytes), int2=060878 (i4), iversion=1 (i4), int3=690706 (i4)
    record   fileid, character(len=100): 'F'|'S' (full/small) + 'T' (tagged) + ':Phantom:version:gitsha (hydro|mhd+clean...): date'
    for each of the 8 data types  (default int, int*1, int*2, int*4, int*8, default real, real*4, real*8):
        record n (i4) ; if n > 0: record n tags character(len=16) ; record n values
    record   nblockarrays = narraylengths * nblocks (i4)       narraylengths = 2 (hydro, sinks) or 4 (+ radiation, MHD)
    for each block: record  number (i8), nums(1:8) (i4) = arrays of each type in the block
    for each block, for each data type, for each array: record tag character(len=16) ; record values(1:number)

A full dump of the gas block carries (readwrite_dumps.f90:194-269): itype (int*1), x y z, vx vy vz [u] (default real), h (real*4),
alpha, divv [curlv], poten (real*4) ...; the MHD block (4th array length) Bx By Bz (real) psi, divB, curlB (real*4).
`read_dump` returns everything by tag; `particles_from_dump` / `write_dump` convert to and from the arrays of the hot path."""
import datetime
import struct

import numpy as np

LENTAG, LENID = 16, 100
INT1, INT2, INT1O, INT2O = 60769, 60878, 690706, 780806
# data type order of utils_dumpfiles.f90:50-57 ; the default integer is 4 bytes, the default real is set per file
I_INT, I_INT1, I_INT2, I_INT4, I_INT8, I_REAL, I_REAL4, I_REAL8 = range(8)


def _dtypes(realsize):
    return [np.dtype("<i4"), np.dtype("<i1"), np.dtype("<i2"), np.dtype("<i4"), np.dtype("<i8"),
            np.dtype("<f8" if realsize == 8 else "<f4"), np.dtype("<f4"), np.dtype("<f8")]


class DumpFormatError(ValueError):
    pass


def _read_record(f):
    head = f.read(4)
    if len(head) < 4:
        raise EOFError
    n, = struct.unpack("<i", head)
    data = f.read(n)
    tail, = struct.unpack("<i", f.read(4))
    if tail != n:
        raise DumpFormatError(f"inconsistent record markers {n} / {tail} (wrong endian or -frecord-marker?)")
    return data


def _write_record(f, data):
    f.write(struct.pack("<i", len(data)))
    f.write(data)
    f.write(struct.pack("<i", len(data)))


def _tags(data):
    return [data[k:k + LENTAG].decode("ascii", "replace").strip() for k in range(0, len(data), LENTAG)]


def _tag(name):
    return name.encode("ascii")[:LENTAG].ljust(LENTAG)


def read_dump(path):
    """-> dict(fileid, realsize, header={tag: scalar | array}, blocks=[{tag: ndarray}], nblocks)"""
    with open(path, "rb") as f:
        first = _read_record(f)
        if len(first) == 20:
            realsize = 4
            i1, r1, i2, iversion, i3 = struct.unpack("<ifiii", first)
        elif len(first) == 24:
            realsize = 8
            i1, r1, i2, iversion, i3 = struct.unpack("<idiii", first)
        else:
            raise DumpFormatError("not a Phantom dump: unexpected first record length %d" % len(first))
        if i1 not in (INT1, INT1O):
            raise DumpFormatError("wrong endian? (open_dumpfile_r: ierr_endian)")
        if i2 not in (INT2, INT2O):
            raise DumpFormatError("default real size wrong (ierr_realsize)")
        if i3 != INT1O:
            raise DumpFormatError("default int size wrong (ierr_intsize)")
        fileid = _read_record(f).decode("ascii", "replace")
        if len(fileid) < 2 or fileid[1] not in "Tt":
            raise DumpFormatError("untagged (pre-2016) dump: not supported (ierr_notags)")
        dts = _dtypes(realsize)
        header = {}
        for k in range(8):
            n, = struct.unpack("<i", _read_record(f))
            if n <= 0:
                continue
            tags = _tags(_read_record(f))
            vals = np.frombuffer(_read_record(f), dtype=dts[k], count=n)
            seen = {}
            for t, v in zip(tags, vals):
                seen.setdefault(t, []).append(v)
            for t, v in seen.items():
                header[t] = v[0] if len(v) == 1 else np.array(v)       # repeated tags = arrays (npartoftype, massoftype, ...)
        nblockarrays, = struct.unpack("<i", _read_record(f))
        nblocks = max(int(header.get("nblocks", 1)), 1)
        if nblockarrays % nblocks:
            raise DumpFormatError("number of array lengths %d is not a multiple of nblocks %d" % (nblockarrays, nblocks))
        narraylengths = nblockarrays // nblocks                          # readwrite_dumps.f90:579
        # MPI dumps repeat {narraylengths block headers, then that block's arrays} once per process (:600-665); the pieces of one
        # array length are concatenated in block order
        per_block = []
        for _ in range(nblocks):
            heads = []
            for _ in range(narraylengths):
                rec = _read_record(f)
                number, = struct.unpack("<q", rec[:8])
                nums = struct.unpack("<8i", rec[8:40])
                heads.append((number, nums))
            parts = []
            for number, nums in heads:
                arrays = {}
                for k in range(8):
                    for _ in range(nums[k]):
                        tag = _tags(_read_record(f))[0]
                        arr = np.frombuffer(_read_record(f), dtype=dts[k]).copy()
                        while tag in arrays:                               # repeated tags (e.g. several dust species): suffix them
                            tag += "_"
                        arrays[tag] = arr
                parts.append(arrays)
            per_block.append(parts)
        blocks = []
        for k in range(narraylengths):
            merged = {}
            for tag in per_block[0][k]:
                merged[tag] = per_block[0][k][tag] if nblocks == 1 else np.concatenate([pb[k][tag] for pb in per_block if tag in pb[k]])
            blocks.append(merged)
    return dict(fileid=fileid, realsize=realsize, iversion=iversion, header=header, blocks=blocks, nblocks=nblocks)


def particles_from_dump(d, params):
    """the arrays of the hot path from a full dump (gas block [+ MHD block]); returns a setups.Particles"""
    from .setups import Particles
    b = d["blocks"][0]
    n = len(b["x"])
    xyzh = np.stack([b["x"], b["y"], b["z"], b["h"].astype(np.float64)], axis=1)
    iphase = b["itype"].astype(np.int8) if "itype" in b else None
    part = Particles(params, xyzh, iphase)
    for k, t in enumerate(("vx", "vy", "vz")):
        part.vxyzu[:, k] = b[t]
    if params.maxvxyzu == 4 and "u" in b:
        part.vxyzu[:, 3] = b["u"]
    if "alpha" in b:
        part.alphaind[:, 0] = b["alpha"]
    if "divv" in b:
        part.divcurlv[:, 0] = b["divv"]
    if "poten" in b:
        part.poten[:] = b["poten"]
    if params.mhd and len(d["blocks"]) >= 4 and "Bx" in d["blocks"][3]:
        m = d["blocks"][3]
        pm = np.array([params.massoftype[t] for t in range(8)])[np.abs(part.iphase.astype(np.int64))]
        rho = pm * (params.hfact / np.abs(xyzh[:, 3])) ** 3
        for k, t in enumerate(("Bx", "By", "Bz")):
            part.Bevol[:, k] = m[t] / rho                               # the dump holds B, the code evolves B/rho (readwrite_dumps.f90:309)
        if "psi" in m:
            part.Bevol[:, 3] = m["psi"]
    assert n == part.npart
    return part


def write_dump(path, part, time=0.0, extra_header=None, small=False, nblocks=1):
    """write a full dump of the particle set in the layout of write_fulldump (readwrite_dumps.f90:85-330), default real = 8 bytes.
    nblocks > 1 writes it the way an MPI run of nblocks processes does: every process its block headers and arrays in turn (:170-172)"""
    p = part.params
    n = part.npart
    mhd = bool(p.mhd)
    dts = _dtypes(8)
    ntypes = 7
    itype = np.abs(part.iphase.astype(np.int64))
    npartoftype = [int(np.sum(itype == t)) for t in range(1, ntypes + 1)]
    hdr = [[] for _ in range(8)]

    def add(k, tag, vals):
        for v in np.atleast_1d(vals):
            hdr[k].append((tag, v))
    add(I_INT, "nparttot", n); add(I_INT, "ntypes", ntypes); add(I_INT, "npartoftype", npartoftype); add(I_INT, "nblocks", nblocks)
    add(I_INT, "nptmass", 0); add(I_INT, "ndustlarge", 1 if npartoftype[6] else 0); add(I_INT, "ndustsmall", 0); add(I_INT, "idust", 7)
    add(I_INT, "majorv", 2026); add(I_INT, "minorv", 0); add(I_INT, "microv", 1)
    add(I_INT8, "nparttot", n); add(I_INT8, "ntypes", ntypes); add(I_INT8, "npartoftype", npartoftype)
    add(I_INT, "iexternalforce", 0); add(I_INT, "ieos", p.ieos)
    add(I_REAL, "time", time); add(I_REAL, "dtmax", min(p.dtmax, 1e30)); add(I_REAL, "gamma", p.gamma); add(I_REAL, "polyk", p.polyk)
    add(I_REAL, "hfact", p.hfact); add(I_REAL, "tolh", p.tolh); add(I_REAL, "C_cour", p.C_cour); add(I_REAL, "C_force", p.C_force)
    add(I_REAL, "alpha", p.alpha); add(I_REAL, "alphau", p.alphau); add(I_REAL, "alphaB", p.alphaB); add(I_REAL, "qfacdisc", p.qfacdisc)
    add(I_REAL, "massoftype", [p.massoftype[t] for t in range(1, ntypes + 1)])
    for t, v in (("xmin", p.xmin), ("xmax", p.xmax), ("ymin", p.ymin), ("ymax", p.ymax), ("zmin", p.zmin), ("zmax", p.zmax)):
        add(I_REAL, t, v)
    add(I_REAL8, "udist", 1.0); add(I_REAL8, "umass", 1.0); add(I_REAL8, "utime", 1.0); add(I_REAL8, "umagfd", 1.0)
    for k, tag, v in (extra_header or []):
        add(k, tag, v)
    opts = "+grav" if p.gravity else ""
    if npartoftype[6]:
        opts += "+dust"
    stamp = datetime.datetime(2026, 1, 1).strftime("%d/%m/%Y %H:%M:%S.0")
    fileid = ("S" if small else "F") + "T:Phantom:2026.0.1:b200sph " + ("(mhd+clean%s)  : " % opts if mhd else "(hydro%s): " % opts) + stamp
    # ---- arrays by block and type (order of readwrite_dumps.f90:194-269, :309-314)
    pm = np.array([p.massoftype[t] for t in range(8)])[itype]
    rho = pm * (p.hfact / np.abs(part.xyzh[:, 3])) ** 3
    gas = [[] for _ in range(8)]
    gas[I_INT1].append(("itype", itype.astype(np.int8)))
    for k, t in enumerate(("x", "y", "z")):
        gas[I_REAL].append((t, part.xyzh[:, k]))
    for k, t in enumerate(("vx", "vy", "vz", "u")[: p.maxvxyzu]):
        gas[I_REAL].append((t, part.vxyzu[:, k]))
    gas[I_REAL4].append(("h", part.xyzh[:, 3].astype(np.float32)))
    if not p.const_av:
        gas[I_REAL4].append(("alpha", part.alphaind[:, 0]))
    gas[I_REAL4].append(("divv", part.divcurlv[:, 0]))
    if p.gravity:
        gas[I_REAL4].append(("poten", part.poten))
    blocks = [(n, gas), (0, [[] for _ in range(8)])]
    if mhd:
        mb = [[] for _ in range(8)]
        for k, t in enumerate(("Bx", "By", "Bz")):
            mb[I_REAL].append((t, part.Bevol[:, k] * rho))
        mb[I_REAL4].append(("psi", part.Bevol[:, 3].astype(np.float32)))
        mb[I_REAL4].append(("divB", part.divcurlB[:, 0]))
        blocks += [(0, [[] for _ in range(8)]), (n, mb)]
    with open(path, "wb") as f:
        _write_record(f, struct.pack("<idiii", INT1, float(INT2), INT2, 1, INT1O))
        _write_record(f, fileid.encode("ascii")[:LENID].ljust(LENID))
        for k in range(8):
            _write_record(f, struct.pack("<i", len(hdr[k])))
            if hdr[k]:
                _write_record(f, b"".join(_tag(t) for t, _ in hdr[k]))
                _write_record(f, np.array([v for _, v in hdr[k]], dtype=dts[k]).tobytes())
        _write_record(f, struct.pack("<i", len(blocks) * nblocks))
        edges = [n * b // nblocks for b in range(nblocks + 1)]
        for b in range(nblocks):
            lo, hi = edges[b], edges[b + 1]
            for number, arrs in blocks:
                _write_record(f, struct.pack("<q8i", (hi - lo) if number else 0, *[len(a) for a in arrs]))
            for number, arrs in blocks:
                for k in range(8):
                    for tag, a in arrs[k]:
                        _write_record(f, _tag(tag))
                        _write_record(f, np.ascontiguousarray(a[lo:hi], dtype=dts[k]).tobytes())
    return fileid
