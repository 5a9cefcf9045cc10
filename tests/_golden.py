"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference)."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_CASES = ["seoul_small", "emb64_k3", "emb128_k4", "node_dropout", "train_mode"]
ALL_CASES = STEP_CASES + ["no_negatives"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.name = name
        self.raw = {k: z[k] for k in z.files}
        self.cfg = ast.literal_eval(str(self.raw["cfg"]))

    def group(self, prefix):
        p = prefix + "/"
        return {k[len(p):]: v for k, v in self.raw.items() if k.startswith(p)}

    def params(self):
        return {k: torch.from_numpy(v.copy()) for k, v in self.group("p").items()}

    def batch(self):
        return {k: torch.from_numpy(v.copy()) for k, v in self.group("batch").items()}

    def lap_list(self):
        out, j = [], 0
        while f"lap/{j}/indices" in self.raw:
            idx = torch.from_numpy(self.raw[f"lap/{j}/indices"].astype(np.int64))
            val = torch.from_numpy(self.raw[f"lap/{j}/values"].copy())
            shape = tuple(int(s) for s in self.raw[f"lap/{j}/shape"])
            out.append(torch.sparse_coo_tensor(idx, val, shape, is_coalesced=False))
            j += 1
        return out

    def out(self, key):
        return self.raw["out/" + key]

    def grads(self):
        return self.group("grad")

    def nograd_keys(self):
        return list(self.group("nograd").keys())


def rel_err(a, b):
    """max|a-b| / max|b|  — the parity metric SURVEY.md section 8(c) defines."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


EVAL_COLS = ("year", "userid", "age", "sex", "month", "day", "dayofweek")


def eval_frame_cols(g, which):
    """Columns of the eval_sampler golden's frames ('total' | 'test') named as oracle.negative_sampling expects."""
    f = g.group(which)
    cols = {c: f[c] for c in EVAL_COLS + ("itemid",)}
    cols["rating"] = f["visitor"]
    return cols


def eval_test_batches(g, users=None, items=None):
    """The reference's test DataLoader over TourDataset(train=False) (main.py:49-52: batch_size=test_batch,
    shuffle=False, drop_last=True) as a list of dict batches."""
    users = torch.from_numpy(g.raw["sampler/test_users"]) if users is None else users
    items = torch.from_numpy(g.raw["sampler/test_items"]) if items is None else items
    tb = g.cfg["test_batch"]
    out = []
    for s in range(0, (len(users) // tb) * tb, tb):
        u = users[s:s + tb]
        out.append(dict(year=u[:, 0], u_id=u[:, 1], age=u[:, 2], sex=u[:, 3], month=u[:, 4], day=u[:, 5], dow=u[:, 6],
                        rating=u[:, 7], pos_item=items[s:s + tb]))
    return out
