"""Boundary proof: the reference's UNMODIFIED ``Experiment`` (model/experiment.py:9-130) driven
over the drop-in ``NGCF`` / ``BPR`` resolved through the two shim modules in ``dropin/`` — the imports a maintainer's
``main.py`` performs (``from NGCF import NGCF``, ``from bprloss import BPR``, main.py:9-10).

The reference modules come from ``baseline/_ref/model`` (a byte-for-byte copy made by ``baseline/install_ref.py`` at
build time; git-ignored, shipped to the GPU box with the snapshot).  /root/reference itself is never read here.
Tolerances: 1e-4 relative on BPR / RMSE; HR / NDCG exact up to one group flipping on a rank tie (as in
tests/test_eval_sampler_gpu.py)."""
import hashlib
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "model")
DROPIN = os.path.join(ROOT, "dropin")
HAVE_REF = os.path.exists(os.path.join(REF, "experiment.py"))


def _import_over_shims():
    """sys.path as a maintainer would have it: dropin/ ahead of the reference's model/ directory."""
    for p in (REF, DROPIN):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [DROPIN, REF]
    argv, sys.argv = sys.argv, [sys.argv[0]]                 # parsers.py:16 parses argv at import time
    try:
        for name in ("NGCF", "bprloss", "experiment"):
            sys.modules.pop(name, None)
        return (importlib.import_module("NGCF"), importlib.import_module("bprloss"),
                importlib.import_module("experiment"))
    finally:
        sys.argv = argv


def test_shims_resolve_to_the_b200_modules():
    import seoul_tourism_recommendation_ngcf_b200 as pkg
    for p in (DROPIN,):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, DROPIN)
    for name in ("NGCF", "bprloss"):
        sys.modules.pop(name, None)
    try:
        assert importlib.import_module("NGCF").NGCF is pkg.NGCF
        assert importlib.import_module("bprloss").BPR is pkg.BPR
    finally:
        sys.path.remove(DROPIN)
        for name in ("NGCF", "bprloss"):
            sys.modules.pop(name, None)


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed (run __graft_entry__.build() in the build container)")
def test_installed_reference_is_unmodified():
    sums = dict(line.split()[::-1] for line in open(os.path.join(REF, "SHA256SUMS")).read().splitlines())
    for f, want in sums.items():
        assert hashlib.sha256(open(os.path.join(REF, f), "rb").read()).hexdigest() == want, f
    if os.path.isdir("/root/reference/model"):               # build container only: equal to the reference tree itself
        for f in sums:
            assert open(os.path.join(REF, f), "rb").read() == open(os.path.join("/root/reference/model", f), "rb").read()


class _TestSet(torch.utils.data.Dataset):
    def __init__(self, users, items):
        self.users, self.items = users, items

    def __len__(self):
        return len(self.users)

    def __getitem__(self, k):
        u = self.users[k]
        return u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7], self.items[k]


class _TrainSet(torch.utils.data.Dataset):
    def __init__(self, users, items):
        self.users, self.items = users, items

    def __len__(self):
        return len(self.users)

    def __getitem__(self, k):
        u = self.users[k]
        return u[0], u[1], u[2], u[3], u[4], u[5], u[6], self.items[k][0], self.items[k][1]


def _close(got, want, n_groups, tol=1e-4):
    got, want = np.asarray([float(x) for x in got], np.float64), np.asarray(want, np.float64)
    assert abs(got[0] - want[0]) <= tol * abs(want[0]) and abs(got[3] - want[3]) <= tol * abs(want[3]), (got, want)
    assert abs(got[1] - want[1]) <= 1.0 / n_groups + 1e-6 and abs(got[2] - want[2]) <= 1.0 / n_groups + 1e-6, (got, want)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not installed")
def test_reference_experiment_runs_over_the_dropin_modules():
    from seoul_tourism_recommendation_ngcf_b200 import synth
    from tests._golden import Golden
    import seoul_tourism_recommendation_ngcf_b200 as pkg
    ngcf_mod, bpr_mod, exp_mod = _import_over_shims()
    assert ngcf_mod.NGCF is pkg.NGCF and bpr_mod.BPR is pkg.BPR
    assert os.path.samefile(exp_mod.__file__, os.path.join(REF, "experiment.py"))     # the reference's own loop
    g = Golden("eval_sampler")
    cfg = g.cfg
    dev = torch.device("cuda")
    nd = synth.num_dict_for(cfg["n_user"], cfg["n_item"])
    n_groups = len(g.raw["sampler/test_users"]) // cfg["test_batch"]
    test = _TestSet(torch.from_numpy(g.raw["sampler/test_users"]), torch.from_numpy(g.raw["sampler/test_items"]))
    te = torch.utils.data.DataLoader(test, batch_size=cfg["test_batch"], shuffle=False, drop_last=True)

    # ---- eval(): the reference's loop (experiment.py:66-119) over our forward / BPR vs the reference's own numbers
    m = ngcf_mod.NGCF(embed_size=cfg["emb"], layer_size=cfg["layers"], node_dropout=cfg["node_p"],
                      mess_dropout=cfg["mess_p"], emb_ratio=cfg["emb_ratio"], lap_list=g.lap_list(), num_dict=nd,
                      batch_size=cfg["B"], device=dev)
    m.load_state_dict(g.params())
    m = m.to(device=dev)
    exp = exp_mod.Experiment(model=m, optimizer=None, criterion=None,
                             test_criterion=bpr_mod.BPR(weight_decay=cfg["wd"], batch_size=cfg["test_batch"]).to(dev),
                             train_dataloader=None, test_dataloader=te, epochs=1, ks=cfg["ks"], device=dev)
    _close(exp.eval(), g.out("metrics"), n_groups)
    assert np.abs(m.user_embedding.weight.detach().cpu().numpy() - g.out("user_after")).max() <= 1e-6
    _close(exp.eval(), g.out("metrics_pass2"), n_groups)
    assert m.all_users_emb.shape == (cfg["n_user"], cfg["emb"] + sum(cfg["layers"]))   # demo.py:233

    # ---- train(): one epoch of experiment.py:32-64 (torch.optim.Adam, zero_grad / backward / step) + its eval();
    # dropout off so the run is deterministic, compared with the drop-in Experiment's eager loop on the same rows
    tr_u, tr_i = torch.from_numpy(g.raw["sampler/train_users"]), torch.from_numpy(g.raw["sampler/train_items"])
    short = _TestSet(test.users[:250], test.items[:250])
    finals = []
    for which in ("reference", "dropin"):
        mm = ngcf_mod.NGCF(cfg["emb"], cfg["layers"], 0.0, [0.0, 0.0], cfg["emb_ratio"], g.lap_list(), nd, 32, dev)
        mm.load_state_dict(g.params())
        mm = mm.to(device=dev)
        opt = torch.optim.Adam(mm.parameters(), lr=1e-2)
        trl = torch.utils.data.DataLoader(_TrainSet(tr_u, tr_i), batch_size=32, shuffle=False, drop_last=True)
        tel = torch.utils.data.DataLoader(short, batch_size=25, shuffle=False, drop_last=True)
        crit, tcrit = bpr_mod.BPR(0.025, 32).to(dev), bpr_mod.BPR(0.025, 25).to(dev)
        if which == "reference":
            exp_mod.Experiment(mm, opt, crit, tcrit, trl, tel, 1, cfg["ks"], dev).train()
        else:
            pkg.Experiment(mm, opt, crit, tcrit, trl, tel, 1, cfg["ks"], dev, verbose=False, graphed=False).train()
        torch.cuda.synchronize()
        finals.append({k: v.detach().cpu().numpy().copy() for k, v in mm.state_dict().items()})
    p0 = {k: v.numpy() for k, v in g.params().items()}
    moved = 0
    for k in finals[0]:
        den = max(1e-30, np.abs(finals[0][k]).max())
        assert np.abs(finals[0][k] - finals[1][k]).max() <= 2e-4 * den, k
        moved += int(np.abs(finals[0][k] - p0[k]).max() > 1e-4)
    assert moved >= 6                                                # the tables and every W/b were trained
