"""Reader for the SIP setup files (`.dat`) -- only what the block workloads need: the segment tables.

Format (src/sip/setup/setup_reader.cpp:329-339,486-612, src/sip/setup/io_utils.cpp:114-185), little-endian:
    int32 magic 0x0160D047, int32 version 1
    list of string                              SIAL program names
    list of (string, int32)                     predefined ints (baocc, eaocc, bavirt, eavirt, norb, ...)
    list of (string, float64)                   predefined scalars
    list of (int32 index_type, int32[] extents) segment tables; index_type 1001 ao, 1002 mo, 1003 moa, 1004 mob
    list of (string, int32 rank, int32[] dims, float64[] data)   predefined arrays
    list of (string, int32 rank, int32[] dims, int32[] data)     predefined integer arrays
    list of (string sialfile, list of (string key, string value)) per-program configuration
every list and array is prefixed by an int32 count; a string is int32 length (INCLUDING a trailing NUL) + bytes.
The files hold geometry / basis / flags only -- no integrals, no amplitudes.
"""
import struct

MAGIC, VERSION = 0x0160D047, 1
INDEX_TYPES = {1001: "ao", 1002: "mo", 1003: "moa", 1004: "mob"}


class SetupFormatError(ValueError):
    pass


class _Stream:
    def __init__(self, data):
        self.d, self.p = data, 0

    def take(self, n):
        if self.p + n > len(self.d):
            raise SetupFormatError("truncated setup file")
        b = self.d[self.p: self.p + n]
        self.p += n
        return b

    def i32(self):
        return struct.unpack("<i", self.take(4))[0]

    def f64(self):
        return struct.unpack("<d", self.take(8))[0]

    def string(self):
        n = self.i32()
        if n < 0 or n > 1 << 20:
            raise SetupFormatError("bad string length")
        return self.take(n).split(b"\0", 1)[0].decode()

    def ints(self):
        n = self.i32()
        return list(struct.unpack(f"<{n}i", self.take(4 * n)))

    def doubles(self):
        n = self.i32()
        return list(struct.unpack(f"<{n}d", self.take(8 * n)))


def read_setup(data):
    """bytes of a .dat file -> dict(programs, ints, scalars, segments{kind: extents}, arrays, int_arrays, configs)"""
    s = _Stream(data)
    if s.i32() != MAGIC:
        raise SetupFormatError("bad magic")
    if s.i32() != VERSION:
        raise SetupFormatError("unsupported version")
    out = {"programs": [s.string() for _ in range(s.i32())]}
    out["ints"] = {}
    for _ in range(s.i32()):
        k = s.string()
        out["ints"][k] = s.i32()
    out["scalars"] = {}
    for _ in range(s.i32()):
        k = s.string()
        out["scalars"][k] = s.f64()
    out["segments"] = {}
    for _ in range(s.i32()):
        t = s.i32()
        out["segments"][INDEX_TYPES.get(t, str(t))] = s.ints()
    out["arrays"] = {}
    for _ in range(s.i32()):
        k = s.string()
        s.i32()  # rank (dims carries its own count)
        dims = s.ints()
        out["arrays"][k] = (dims, s.doubles())
    out["int_arrays"] = {}
    for _ in range(s.i32()):
        k = s.string()
        s.i32()
        dims = s.ints()
        out["int_arrays"][k] = (dims, s.ints())
    out["configs"] = {}
    for _ in range(s.i32()):
        prog = s.string()
        out["configs"][prog] = {}
        for _ in range(s.i32()):
            k = s.string()
            out["configs"][prog][k] = s.string()
    out["trailing_bytes"] = len(data) - s.p
    return out


def occ_virt_segments(setup):
    """(occupied extents, virtual extents) of the alpha MO index from baocc..eaocc / bavirt..eavirt (1-based segments)"""
    seg, ints = setup["segments"]["moa"], setup["ints"]
    return seg[ints["baocc"] - 1: ints["eaocc"]], seg[ints["bavirt"] - 1: ints["eavirt"]]
