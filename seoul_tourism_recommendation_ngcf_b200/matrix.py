"""Drop-in ``Matrix`` (reference: model/matrix.py:12-83) — the ``lap_list`` builder, restricted to the non-zeros.

Same constructor and ``create_matrix()`` as the reference, same output (a list indexed by ``year % 18`` of uncoalesced
int64/fp32 ``torch.sparse_coo`` tensors on ``device``, optionally pickled as ``lap_list_*.pkl``), including its
quirks: R is never reset between years (matrix.py:33,45), the degree counts non-zeros while the values carry the
ratings (matrix.py:55-62).  The reference goes through dense N x N arrays (``toarray()`` + ``multi_dot``: O(N^2)
memory, O(N^3) time — 40 GB and hours at Gowalla size); this is O(nnz) and builds the Gowalla-shaped Laplacian in under
a second (laplacian.py).  Host-side and one-off, like the reference's; the conversion to the kernels' CSR happens once
per element on the device (plan.py).
"""
from __future__ import annotations

import os
import pickle
from datetime import datetime

import torch.nn as nn

from .laplacian import build_lap_list


class Matrix(nn.Module):
    def __init__(self, total_df, cols: list, rating_col: str, num_dict: dict, folder_path: str, save_data: bool, device,
                 args=None, builder: str = "host"):
        super().__init__()
        self.df = total_df[cols]
        self.rating_col = rating_col
        self.folder_path = folder_path
        self.save_data = save_data
        self.device = device
        self.n_user = num_dict['user']
        self.n_item = num_dict['item']
        self.args = args                  # the reference reads parsers.args for the pickle's file name (matrix.py:72)
        # "host": numpy, bit-identical to the reference's values; "device": ngcf_laplacian_entries on `device` (CUDA),
        # equal to 5e-7 relative (d^-1/2 correctly rounded instead of numpy's float32 power) — for graphs where even the O(nnz) host pass is the slow part
        # "csr": built on the device straight into the kernels' CSR (plgraph.CsrLaplacian elements: no COO at all; NGCF
        # accepts them in lap_list like the reference's sparse tensors)
        if builder not in ("host", "device", "csr"):
            raise ValueError("builder is 'host', 'device' or 'csr'")
        self.builder = builder
        self.lap_list = [[] for _ in self.df['year'].unique()]

    def create_matrix(self):
        df = self.df
        laps = build_lap_list(df['year'].to_numpy(), df['userid'].to_numpy(), df['itemid'].to_numpy(),
                              df[self.rating_col].to_numpy(), self.n_user, self.n_item,
                              device=self.device if self.builder != "host" else None,
                              fmt="csr" if self.builder == "csr" else "coo")
        self.lap_list = [L.to(self.device) if hasattr(L, "is_sparse") and L.is_sparse else L for L in laps]
        print('Laplacian Matrix Created!')
        if self.save_data:
            a, d1 = self.args, datetime.now()
            if a is None:
                try:
                    from parsers import args as a     # the reference's module, when this runs inside its tree
                except Exception:
                    a = None
            tag = (f'{a.epoch}_{a.batch_size}_{a.lr}_{a.emb_ratio}_{a.scaler}' if a is not None else 'ngcf_b200')
            path = os.path.join(self.folder_path, f'lap_list_implicit_{tag}_{d1.month}_{d1.day}_{d1.hour}_{d1.minute}.pkl')
            with open(path, 'wb') as f:
                pickle.dump(self.lap_list, f)
            print('Laplacian data Saved!')
        return self.lap_list
