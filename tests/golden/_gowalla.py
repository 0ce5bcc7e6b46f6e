"""Decode /root/reference/gowalla_nevda.7z without a 7z tool (fixture generation only).

The archive is one solid LZMA2 stream (dict 16 MiB) of 12 files; offsets and
sizes were read from its (LZMA1-packed) header during the survey.
Only used by make_golden.py in the build container; never at test/run time.
"""
import lzma
import pickle

ARCHIVE = "/root/reference/gowalla_nevda.7z"
MAIN_OFF, MAIN_LEN = 32, 2696911
NAMES = ("Graph_adj.csv Graph_cat.csv Graph_dist.csv Graph_poi.csv test.pickle test_idx.pkl "
         "train.pickle train_idx.pkl pre_filter.pt pre_transform.pt test.pt train.pt").split()
SIZES = [54161455, 258520, 54161132, 162343, 10375706, 13844, 47568065, 20252, 437, 443, 908031, 3055697]


def unpack():
    raw = open(ARCHIVE, "rb").read()
    dec = lzma.LZMADecompressor(format=lzma.FORMAT_RAW,
                                filters=[{"id": lzma.FILTER_LZMA2, "dict_size": 16 << 20}])
    data = dec.decompress(raw[MAIN_OFF:MAIN_OFF + MAIN_LEN])
    assert len(data) == sum(SIZES)
    out, off = {}, 0
    for n, s in zip(NAMES, SIZES):
        out[n] = data[off:off + s]
        off += s
    return out


def sessions(blob):
    """train.pickle / test.pickle -> flat list of per-trajectory dicts (torch tensors inside)."""
    d = pickle.loads(blob)
    flat = []
    for u in d:
        for s in d[u]:
            flat.append(d[u][s])
    return flat


# ---- digests shared by make_gowalla_real.py (computed there from the REFERENCE's own outputs) and tests/test_owndata.py
ITEM_FIELDS = (("x", "int64"), ("edge_index", "int64"), ("edge_attr", "int64"), ("y", "int64"), ("time", "int64"),
               ("time_normal", "float32"), ("user", "int64"), ("cat", "int64"))


def items_digest(items):
    """sha256 over every field of every item (shape + little-endian bytes), in order."""
    import hashlib
    import numpy as np
    h = hashlib.sha256()
    for it in items:
        for name, dt in ITEM_FIELDS:
            v = getattr(it, name)
            a = np.ascontiguousarray(v.numpy() if hasattr(v, "numpy") else v).astype(dt)
            h.update(name.encode() + str(a.shape).encode())
            h.update(a.tobytes())
    return h.hexdigest()


def csr_dense_digest(csr, n):
    """sha256 of the dense float32 [n, n] matrix a CSR triple stands for (row by row: never holds more than one row block)."""
    import hashlib
    import numpy as np
    crow, col, val = (np.asarray(a) for a in csr)
    h = hashlib.sha256()
    for r0 in range(0, n, 256):
        r1 = min(n, r0 + 256)
        blk = np.zeros((r1 - r0, n), np.float32)
        for r in range(r0, r1):
            blk[r - r0, col[crow[r]:crow[r + 1]]] = val[crow[r]:crow[r + 1]]
        h.update(blk.tobytes())
    return h.hexdigest()
