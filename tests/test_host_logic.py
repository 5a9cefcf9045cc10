"""CPU-only checks of the host side: module surface, state_dict layout, init stream, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import seoul_tourism_recommendation_ngcf_b200 as pkg
from seoul_tourism_recommendation_ngcf_b200 import _lib, synth
from tests._golden import Golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model_from_golden(g, **kw):
    cfg = g.cfg
    nd = synth.num_dict_for(cfg["n_user"], cfg["n_item"])
    return pkg.NGCF(cfg["emb"], cfg["layers"], cfg.get("node_p", 0.3), cfg.get("mess_p", [0.1] * len(cfg["layers"])),
                    cfg.get("emb_ratio", 1.0), g.lap_list(), nd, cfg.get("B", 512), torch.device("cpu"), **kw)


def test_library_loads_and_exports_every_declared_symbol():
    """include/ngcf_b200.h <-> libngcf_b200.so <-> the ctypes table, no compute calls."""
    header = open(os.path.join(ROOT, "include", "ngcf_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ngcf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.ngcf_abi_version() == 4
    assert lib.ngcf_spmm_split_threshold() > 0
    # argument validation happens before any CUDA call
    need = ctypes.c_size_t(0)
    assert lib.ngcf_score_topk_workspace(4, 10, 64, 0, ctypes.byref(need)) != 0
    assert b"score_topk_workspace" in lib.ngcf_last_error()


@pytest.mark.parametrize("name", ["seoul_small", "node_dropout"])
def test_state_dict_layout_and_init_stream_match_reference(name):
    """Same keys, order, shapes, dtypes as the reference state_dict, and — for emb % 5 == 0 — the same initial
    values from the same torch seed (NGCF.py:39-45, 56-91 call order)."""
    g = Golden(name)
    seed = {"seoul_small": 0, "node_dropout": 3}[name]
    torch.manual_seed(seed)
    m = _model_from_golden(g)
    ref = g.params()
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
        assert torch.equal(sd[k], ref[k]), f"init stream differs at {k}"
    assert [n for n, _ in m.named_parameters()] == list(ref.keys())
    assert len(list(m.buffers())) == 0


def test_reference_checkpoint_layout_loads():
    g = Golden("ckpt_demo")                      # weights of the .pth demo.py:82 loads, users cut to 300
    m = _model_from_golden(g)
    missing, unexpected = m.load_state_dict(g.params())
    assert not missing and not unexpected
    assert m.w1_list[0].weight.shape == (64, 65)


def test_emb_not_multiple_of_five_uses_remainder_table():
    g = Golden("emb64_k3")
    m = _model_from_golden(g)
    assert m.feat_widths == [12, 12, 12, 12, 16]
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in g.params().items()}


def test_cpu_forward_fails_loudly():
    g = Golden("seoul_small")
    m = _model_from_golden(g)
    b = g.batch()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(b["year"], b["u_id"], b["age"], b["sex"], b["month"], b["day"], b["dow"], b["pos_item"], b["neg_item"], False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.BPR(0.025, 16)(torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.score_topk(torch.zeros(2, 4), torch.zeros(3, 4), 2)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "seoul_tourism_recommendation_ngcf_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_synthetic_graph_generator():
    u, i, r = synth.powerlaw_bipartite(300, 200, 5000, alpha=0.8, seed=0)
    assert u.size == 5000 and len(set(zip(u.tolist(), i.tolist()))) == 5000
    assert u.min() >= 0 and u.max() < 300 and i.min() >= 0 and i.max() < 200
    u2, i2, _ = synth.powerlaw_bipartite(300, 200, 5000, alpha=0.8, seed=0)
    assert np.array_equal(u, u2) and np.array_equal(i, i2)
    deg = np.bincount(i, minlength=200)
    assert deg.max() > 5 * max(1.0, np.median(deg))           # skewed popularity
    b = synth.random_batch(300, 200, 64)
    assert set(b) == {"year", "u_id", "age", "sex", "month", "day", "dow", "pos_item", "neg_item"}


def test_greedy_tiles_cover_rows_within_caps():
    """Row tiling of the execution plan (plan.greedy_tiles): consecutive, complete, capped."""
    from seoul_tourism_recommendation_ngcf_b200.plan import greedy_tiles
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 129, size=5000)
    deg[rng.integers(0, 5000, 200)] = 0
    rp = np.concatenate([[0], np.cumsum(deg)])
    for max_rows, max_ent in ((16, 512), (64, 2048), (1, 128), (16, 128)):
        t = greedy_tiles(rp, max_rows, max_ent)
        assert t.dtype == np.int32 and t.shape[1] == 4
        assert t[0, 0] == 0 and t[-1, 1] == 5000 and np.array_equal(t[1:, 0], t[:-1, 1])
        assert np.all(t[:, 1] - t[:, 0] >= 1) and np.all(t[:, 1] - t[:, 0] <= max_rows)
        assert np.all(t[:, 3] - t[:, 2] <= max_ent)
        assert np.array_equal(t[:, 2], rp[t[:, 0]]) and np.array_equal(t[:, 3], rp[t[:, 1]])
    # a single over-long row gets a tile of its own instead of an endless loop
    t = greedy_tiles(np.array([0, 5, 1005, 1010]), 16, 512)
    assert t.tolist() == [[0, 1, 0, 5], [1, 2, 5, 1005], [2, 3, 1005, 1010]]
    assert greedy_tiles(np.array([0]), 16, 512).shape == (0, 4)


def test_matrix_dropin_builds_the_reference_lap_list(tmp_path):
    """Drop-in ``Matrix`` (matrix.py:12-83) over the sparse builder: same lap_list as the reference built from the same
    frame (golden fixture, generated by the reference's dense Matrix.create_matrix), and the pickle round trip."""
    import pickle
    import pandas as pd
    from seoul_tourism_recommendation_ngcf_b200.matrix import Matrix
    from tests._golden import Golden
    g = Golden("seoul_small")
    f = g.group("frame")
    df = pd.DataFrame({k: f[k] for k in ("year", "userid", "itemid", "visitor")})
    df["unused"] = 0
    m = Matrix(df, ["year", "userid", "itemid", "visitor"], "visitor", {"user": g.cfg["n_user"], "item": g.cfg["n_item"]},
               str(tmp_path), True, torch.device("cpu"))
    laps = m.create_matrix()
    ref = g.lap_list()
    assert len(laps) == len(ref)
    for a, b in zip(laps, ref):
        assert tuple(a.shape) == tuple(b.shape) and not a.is_coalesced()
        assert torch.equal(a._indices(), b._indices()) and torch.equal(a._values(), b._values())
    files = [p for p in os.listdir(tmp_path) if p.startswith("lap_list_implicit_") and p.endswith(".pkl")]
    assert len(files) == 1
    with open(os.path.join(tmp_path, files[0]), "rb") as fh:
        back = pickle.load(fh)
    assert torch.equal(back[1]._values(), ref[1]._values())
