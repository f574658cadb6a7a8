"""Minimal reader for the tiny subset of HDF5 the reference fixture uses
(superblock v0, old-style group with one symbol-table node, v1 object headers,
contiguous little-endian IEEE float datasets).  h5py is not installed in the
build image, and the GPU box has no /root/reference, so tests/golden/make_golden.py
uses this once to convert /root/reference/tests/testdata.hdf5 into an .npz.
Test infrastructure only.
"""
import struct
import numpy as np


def read_hdf5_flat(path):
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0, "only superblock v0"
    so, sl = b[13], b[14]                     # size of offsets / lengths
    assert so == 8 and sl == 8
    # superblock v0: 24 bytes header, then base, free, eof, driver addresses, then root symtab entry
    root = 24 + 4 * 8
    link_name_off, ohdr_addr, cache_type = struct.unpack_from("<QQI", b, root)
    btree_addr, heap_addr = struct.unpack_from("<QQ", b, root + 24)
    assert b[heap_addr:heap_addr + 4] == b"HEAP"
    heap_data = struct.unpack_from("<Q", b, heap_addr + 8 + 16)[0]

    def name_at(off):
        s = heap_data + off
        return b[s:b.index(b"\0", s)].decode()

    out = {}

    def walk_tree(addr):
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        p = addr + 8 + 16
        for k in range(nent):
            child = struct.unpack_from("<Q", b, p + 8)[0]
            p += 16
            if level > 0:
                walk_tree(child)
            else:
                walk_snod(child)

    def walk_snod(addr):
        assert b[addr:addr + 4] == b"SNOD"
        nsym = struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 8
        for k in range(nsym):
            noff, oaddr = struct.unpack_from("<QQ", b, p)
            out[name_at(noff)] = read_dataset(oaddr)
            p += 40

    def read_dataset(addr):
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        p = addr + 16
        end = p + hsize
        shape = dtype = data_addr = None
        seen = 0
        while seen < nmsg and p < end:
            mtype, msize, _ = struct.unpack_from("<HHB", b, p)
            body = p + 8
            if mtype == 0x0001:                      # dataspace
                v, rank, flags = struct.unpack_from("<BBB", b, body)
                off = body + (8 if v == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, b, off)
            elif mtype == 0x0003:                    # datatype
                cls = b[body] & 0x0F
                size = struct.unpack_from("<I", b, body + 4)[0]
                assert cls == 1, "float only"
                dtype = {4: "<f4", 8: "<f8"}[size]
            elif mtype == 0x0008:                    # layout
                v = b[body]
                assert v == 3 and b[body + 1] == 1, "contiguous v3 only"
                data_addr = struct.unpack_from("<Q", b, body + 2)[0]
            elif mtype == 0x0010:                    # continuation
                caddr, clen = struct.unpack_from("<QQ", b, body)
                p, end = caddr - 8 - 0, caddr + clen
                p = caddr
                seen += 1
                continue
            p = body + msize
            seen += 1
        n = int(np.prod(shape))
        return np.frombuffer(b, dtype=dtype, count=n, offset=data_addr).reshape(shape).copy()

    walk_tree(btree_addr)
    return out
