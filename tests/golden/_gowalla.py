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
