"""Minimal single-cabinet .cab reader used by the tests to cut codec units out of the
reference's fixture cabinets.

Layout facts follow libmspack/mspack/cab.h:16-45 (CFHEADER 0x24 bytes, CFFOLDER 8 bytes,
CFDATA 8 bytes, optional reserved areas) and the framing cabd.c applies before the codec sees
the bytes: payloads are concatenated per folder (cabd.c:1294-1344) and Quantum gets a 0xFF
trailer after every block (cabd.c:1330-1332).  Multi-cabinet folders are not handled here.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List


@dataclass
class CabFolder:
    comp_type: int            # raw typeCompress (method | level << 8)
    method: int               # 0 none, 1 MSZIP, 2 Quantum, 3 LZX
    window_bits: int          # (typeCompress >> 8) & 0x1f for Quantum / LZX
    blocks: List[bytes] = field(default_factory=list)      # CFDATA payloads
    usizes: List[int] = field(default_factory=list)        # uncompressed size per block

    @property
    def out_len(self) -> int:
        return sum(self.usizes)

    def unit_bytes(self) -> bytes:
        """Bytes the codec's `read` callback delivers for this folder."""
        if self.method == 2:
            return b"".join(b + b"\xff" for b in self.blocks)
        return b"".join(self.blocks)


def parse_cab(data: bytes) -> List[CabFolder]:
    if data[:4] != b"MSCF":
        raise ValueError("not a cabinet")
    (_, _, _, _, files_off, _, _, _, nfolders, nfiles, flags, _, _) = struct.unpack_from("<4sIIIIIBBHHHHH", data, 0)
    pos = 0x24
    hdr_res = fol_res = dat_res = 0
    if flags & 4:
        hdr_res, fol_res, dat_res = struct.unpack_from("<HBB", data, pos)
        pos += 4 + hdr_res
    for bit in (1, 2):                       # prev / next cabinet name + disk label
        if flags & bit:
            for _ in range(2):
                pos = data.index(b"\0", pos) + 1
    folders = []
    for _ in range(nfolders):
        off, nblocks, ctype = struct.unpack_from("<IHH", data, pos)
        pos += 8 + fol_res
        f = CabFolder(ctype, ctype & 0x0F, (ctype >> 8) & 0x1F)
        p = off
        for _b in range(nblocks):
            if p + 8 > len(data):
                break
            _csum, csize, usize = struct.unpack_from("<IHH", data, p)
            p += 8 + dat_res
            f.blocks.append(data[p:p + csize])
            f.usizes.append(usize)
            p += csize
        folders.append(f)
    return folders


# ------------------------------------------------------------------------------------------------------------
# A writer, for the cabinet-level tests: single cabinet, no reserved areas.

def cab_checksum(data: bytes, seed: int = 0) -> int:
    """cabd_checksum (cabd.c:1456-1479): XOR of the little-endian 32-bit words, the 1-3 tail bytes packed
    b0<<16 | b1<<8 | b2 (3), b0<<8 | b1 (2), b0 (1)."""
    import numpy as np
    n = len(data) & ~3
    s = seed
    if n:
        s ^= int(np.bitwise_xor.reduce(np.frombuffer(data[:n], dtype="<u4")))
    t = data[n:]
    ul = 0
    if len(t) == 3:
        ul = (t[0] << 16) | (t[1] << 8) | t[2]
    elif len(t) == 2:
        ul = (t[0] << 8) | t[1]
    elif len(t) == 1:
        ul = t[0]
    return (s ^ ul) & 0xFFFFFFFF


def build_cab(folders, with_checksums: bool = True, prev=None, next=None, set_id: int = 0x1234, set_index: int = 0) -> bytes:
    """folders: list of dicts {comp_type, blocks: [(payload bytes, uncompressed size)], files: [(name, offset, length[, folder index])]}.
    A file's optional 4th element overrides its folder index (0xFFFD continued from the previous cabinet, 0xFFFE continued in the
    next one, 0xFFFF both: cab.h:55-57); prev / next = (cabinet name, disk label) set the header flags 1 / 2 (cab.h:60-62)."""
    nfiles = sum(len(f["files"]) for f in folders)
    names = b""
    flags = 0
    for bit, pn in ((1, prev), (2, next)):
        if pn:
            flags |= bit
            names += pn[0].encode() + b"\0" + pn[1].encode() + b"\0"
    hdr_len = 0x24 + len(names) + 8 * len(folders)
    files_blob = b""
    for i, f in enumerate(folders):
        for ent in f["files"]:
            name, off, length = ent[:3]
            fidx = ent[3] if len(ent) > 3 else i
            files_blob += struct.pack("<IIHHHH", length, off, fidx, 0x2A21, 0x6000, 0x20) + name.encode() + b"\0"
    data_off = hdr_len + len(files_blob)
    fold_blob, data_blob = b"", b""
    for f in folders:
        fold_blob += struct.pack("<IHH", data_off + len(data_blob), len(f["blocks"]), f["comp_type"])
        for payload, usize in f["blocks"]:
            tail = struct.pack("<HH", len(payload), usize)
            csum = cab_checksum(tail, cab_checksum(payload)) if with_checksums else 0
            data_blob += struct.pack("<I", csum) + tail + payload
    total = data_off + len(data_blob)
    head = struct.pack("<4sIIIIIBBHHHHH", b"MSCF", 0, total, 0, hdr_len, 0, 3, 1, len(folders), nfiles, flags, set_id, set_index)
    return head + names + fold_blob + files_blob + data_blob
