"""Build-container script: the REAL Gowalla-Nevada dataset the reference ships (`/root/reference/gowalla_nevda.7z`) through
the reference's OWN dataset code, next to mobgt_b200.owndata, frozen as tests/golden/gowalla_nevda_real.npz.

  1. unpack the archive (tests/golden/_gowalla.py) into a scratch `raw/` directory;
  2. import the UNMODIFIED /root/reference/graphormer/owndata.py behind import stubs (torch_geometric.data.Data = an attribute
     bag; `collate` = identity so that `torch.save` stores the item list itself) and run `GowallaGraph.process`
     (owndata.py:375-460) on it -> the reference's train / test items in the reference's order;
  3. import the UNMODIFIED model_fqandtoyo.py and run `calculate_laplacian_matrix(., 'hat_rw_normd_lap_mat')` (:458-488) +
     `.to(torch.float)` on Graph_dist.csv / Graph_cat.csv, as its constructor does (:653-665);
  4. check mobgt_b200.owndata.load_items / load_world against 2. and 3. EXACTLY (every field of every item, every matrix
     entry), and store sha256 digests of the REFERENCE's outputs in the fixture;
  5. write the dataset in the compact form of owndata.pack_dataset (0.9 MB instead of 166 MB) — the fixture travels to the GPU
     box, where tests / bench.py --workload c4-gowalla-real train and evaluate on the real trajectories.

Run:  python tests/golden/make_gowalla_real.py        (needs /root/reference; ~1 minute)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
REF = "/root/reference/graphormer"


class Data:
    """stand-in for torch_geometric.data.Data: keyword fields + attribute assignment (owndata.py:438-444)"""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    import _gowalla
    import make_model_golden as mg
    from mobgt_b200 import owndata as mine
    tmp = tempfile.mkdtemp(prefix="gowalla_real_")
    raw = os.path.join(tmp, "raw")
    os.makedirs(raw)
    os.makedirs(os.path.join(tmp, "processed"))
    for name, blob in _gowalla.unpack().items():
        with open(os.path.join(raw, name), "wb") as f:
            f.write(blob)
    mg.install_stubs()
    del sys.modules["owndata"]                               # the model golden stubs it out; here the real one is the subject
    sys.modules["torch_geometric.data"].__dict__.update(Data=Data, InMemoryDataset=type("InMemoryDataset", (), {}),
                                                        download_url=None, extract_zip=None)
    sys.path.insert(0, REF)
    import owndata as ref_owndata                            # the unmodified reference module
    assert os.path.samefile(ref_owndata.__file__, os.path.join(REF, "owndata.py"))
    ref_owndata.collate = lambda data_list: data_list
    fake = types.SimpleNamespace(raw_dir=raw, processed_dir=os.path.join(tmp, "processed"), pre_filter=None, pre_transform=None)
    ref_owndata.GowallaGraph.process(fake)
    ref = {s: torch.load(os.path.join(tmp, "processed", f"{s}.pt"), weights_only=False) for s in ("train", "test")}
    got = {s: mine.load_items(raw, s) for s in ("train", "test")}
    for s in ("train", "test"):
        assert len(ref[s]) == len(got[s]), (s, len(ref[s]), len(got[s]))
        for a, b in zip(ref[s], got[s]):
            for name, dt in _gowalla.ITEM_FIELDS:
                va, vb = getattr(a, name).numpy(), getattr(b, name)
                assert va.shape == vb.shape and np.array_equal(va, vb), (s, b.idx, name)
        assert _gowalla.items_digest(ref[s]) == _gowalla.items_digest(got[s])
    print("items: train", len(got["train"]), "test", len(got["test"]), "== reference owndata.GowallaGraph.process")

    import pandas as pd
    import model_fqandtoyo as ref_model                      # the unmodified reference module
    world = mine.load_world(raw, "gowalla_nevda")
    digests = {}
    for name, csv, csr, n in (("C_A", "Graph_cat.csv", world.C_A, world.C), ("D_A", "Graph_dist.csv", world.D_A, world.P)):
        a = pd.read_csv(os.path.join(raw, csv)).to_numpy()
        dense = torch.from_numpy(ref_model.calculate_laplacian_matrix(a, mat_type="hat_rw_normd_lap_mat")).to(dtype=torch.float).numpy()
        crow, col, val = csr
        mine_dense = np.zeros((n, n), np.float32)
        mine_dense[np.repeat(np.arange(n), np.diff(crow)), col] = val
        assert np.array_equal(mine_dense, np.asarray(dense)), name
        import hashlib
        digests[name] = hashlib.sha256(np.ascontiguousarray(dense, np.float32).tobytes()).hexdigest()
        assert digests[name] == _gowalla.csr_dense_digest(csr, n)
        print(name, "== reference calculate_laplacian_matrix, nnz", len(val))
    packed = mine.pack_dataset(world, got)
    packed["ref_digest_train"] = np.array(_gowalla.items_digest(ref["train"]))
    packed["ref_digest_test"] = np.array(_gowalla.items_digest(ref["test"]))
    packed["ref_digest_D_A"] = np.array(digests["D_A"])
    packed["ref_digest_C_A"] = np.array(digests["C_A"])
    out = os.path.join(HERE, "gowalla_nevda_real.npz")
    np.savez_compressed(out, **packed)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
