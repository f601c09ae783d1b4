"""Data side of the drop-in surface: the dataset contract of the reference's ``datasets/dataset.py``
(``num_classes``, ``num_views``, ``dims``, ``__getitem__ -> [x_0 .. x_{V-1}, y]``; README.md:54-70,
datasets/dataset.py:164-268), its ``.mat`` loaders (:270-328), the synthetic two-modality generator (:331-471) and --
SURVEY §8f-4 -- a DEVICE-RESIDENT loader: every shipped set fits in a sliver of the 180 GB of HBM, so batches are
index-selected on the GPU from resident tensors instead of being collated on the host and copied per step.

Everything random replays the reference's draw ORDER on the same generators (numpy global RNG for the conflict /
noise post-processing, a seeded ``torch.Generator`` for the synthetic set), so equal seeds give bit-identical data
(``tests/test_cpu_host.py::test_synthetic_dataset_matches_reference`` / ``::test_multiview_postprocessing...``).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, random_split


def _data_root() -> str:
    return os.environ.get("DMF_DATA_ROOT", "data")


class MultiViewDataset(Dataset):
    """datasets/dataset.py:164-268.  ``data_X``: object array / sequence of V matrices [N, d_v]; ``data_Y``: labels
    (0- or 1-based).  Views are min-max scaled per feature to [0,1] (``norm_min=0``) or [-1,1]."""

    def __init__(self, data_name, data_X, data_Y, norm_min=0):
        super().__init__()
        self.data_name = data_name
        self.num_views = data_X.shape[0] if hasattr(data_X, "shape") else len(data_X)
        self.X = [self.normalize(data_X[v], min=norm_min) for v in range(self.num_views)]
        y = np.squeeze(data_Y)
        if np.min(y) == 1:
            y = y - 1
        self.Y = y.astype(dtype=np.int64)
        self.num_classes = len(np.unique(self.Y))
        self.dims = self.get_dims()

    def __getitem__(self, index):
        return [self.X[v][index].astype(np.float32) for v in range(self.num_views)] + [self.Y[index]]

    def __len__(self):
        return len(self.X[0])

    def get_dims(self):
        return np.array([[self.X[v].shape[1]] for v in range(self.num_views)])

    @staticmethod
    def normalize(x, min=0):
        from sklearn.preprocessing import MinMaxScaler
        return MinMaxScaler((0, 1) if min == 0 else (-1, 1)).fit_transform(x)

    # ---- test-set corruption (datasets/dataset.py:226-268); numpy global RNG, same call order as the reference
    def postprocessing(self, index, addNoise=False, sigma=0, ratio_noise=0.5, addConflict=False, ratio_conflict=0.5):
        if addNoise:
            self.addNoise(index, ratio_noise, sigma=sigma)
        if addConflict:
            self.addConflict(index, ratio_conflict)

    def addNoise(self, index, ratio, sigma):
        for i in np.random.choice(index, size=int(ratio * len(index)), replace=False):
            k = np.random.randint(1, self.num_views + 1)
            for v in np.random.choice(np.arange(self.num_views), size=k, replace=False):
                self.X[v][i] = np.random.normal(self.X[v][i], sigma)

    def addConflict(self, index, ratio):
        proto = {}
        for c in range(self.num_classes):
            members = np.where(self.Y == c)[0]
            if len(members):
                proto[c] = {v: self.X[v][members[0]].copy() for v in range(self.num_views)}
        for i in np.random.choice(index, size=int(ratio * len(index)), replace=False):
            v = np.random.randint(self.num_views)
            if proto:
                # one view is replaced by the prototype of the NEXT class; the label stays
                self.X[v][i] = proto[(self.Y[i] + 1) % self.num_classes][v]


def _mat(name):
    import scipy.io as sio
    return sio.loadmat(os.path.join(_data_root(), name))


def HandWritten():          # views 240 76 216 47 64 6
    m = _mat("handwritten.mat")
    return MultiViewDataset("HandWritten", m["X"][0], m["Y"])


def _transposed(m, key_y):
    X = m["X"][0]
    for v in range(len(X)):
        X[v] = X[v].T
    return X, m[key_y]


def Scene():                # views 20 59 40
    X, y = _transposed(_mat("scene15_mtv.mat"), "gt")
    return MultiViewDataset("Scene", X, y)


def PIE():                  # views 484 256 279
    X, y = _transposed(_mat("PIE_face_10.mat"), "gt")
    return MultiViewDataset("PIE", X, y)


def Caltech():              # views 48 40 254 1984 512 928
    m = _mat("Caltech101-20.mat")
    return MultiViewDataset("Caltech", m["X"].squeeze(), m["Y"])


def CUB():                  # views 1024 300
    m = _mat("cub_googlenet_doc2vec_c10.mat")
    return MultiViewDataset("CUB", m["X"][0], m["gt"] - 1)


# --------------------------------------------------------------------------------------------------------------
# synthetic two-modality set (datasets/dataset.py:331-458): draws in the reference order on one seeded generator
# --------------------------------------------------------------------------------------------------------------
def _rand_orthogonal(d, g):
    Q, R = torch.linalg.qr(torch.randn(d, d, generator=g))
    return Q @ torch.diag(torch.sign(torch.diag(R)))


class SimpleTwoModalPlus(Dataset):
    """Two modalities with tunable dependence ``rho`` between their signal blocks, class signal split between a
    shared channel (``shared_class_frac``) and per-modality private channels, optional cross-modal conflict
    (a rotation of the shared class mean in modality 2 for a random subset of classes), spurious dims and
    heteroscedastic observation noise.  ``__getitem__ -> (x1, x2, y)``."""

    def __init__(self, n_samples=1000, n_classes=3, d_signal=16, d_spurious=16, rho=0.5, shared_class_frac=1.0,
                 class_sep_shared=1.0, class_sep_private=1.0, alpha_shared=0.7, beta_specific=0.6, noise_std=0.8,
                 hetero_noise=True, hetero_scale=0.5, nonlinear_shared=True, nonlinear_specific=False,
                 conflict_frac=0.5, conflict_strength=0.8, seed=0):
        super().__init__()
        if not (0.0 <= rho <= 1.0 and 0.0 <= shared_class_frac <= 1.0):
            raise AssertionError("rho and shared_class_frac must lie in [0, 1]")
        g = torch.Generator().manual_seed(seed)
        n, d = n_samples, d_signal
        randn = lambda *s: torch.randn(*s, generator=g)          # noqa: E731
        # draw order: y, S0, E1, E2, mu_sh, mu_p1, mu_p2, conflict mask, rotations, U1, U2, spur1, spur2, m1, m2, eps1, eps2
        y = torch.randint(0, n_classes, (n,), generator=g)
        S0 = randn(n, d)
        a = math.sqrt(rho)
        E1, E2 = randn(n, d), randn(n, d)
        G1 = a * S0 + math.sqrt(1 - a * a) * E1
        G2 = a * S0 + math.sqrt(1 - a * a) * E2
        mu_sh = randn(n_classes, d) * class_sep_shared
        mu_p1 = randn(n_classes, d) * class_sep_private
        mu_p2 = randn(n_classes, d) * class_sep_private
        mu_sh_y, mu_p1_y, mu_p2_y = mu_sh[y], mu_p1[y], mu_p2[y]
        conflict = torch.rand(n_classes, generator=g) < conflict_frac
        rots = []
        for c in range(n_classes):
            if conflict[c]:
                rots.append((1.0 - conflict_strength) * torch.eye(d) + conflict_strength * _rand_orthogonal(d, g))
            else:
                rots.append(torch.eye(d))
        mu_sh_y_mod2 = torch.bmm(mu_sh_y.unsqueeze(1), torch.stack(rots)[y]).squeeze(1)
        U1, U2 = randn(n, d), randn(n, d)

        def channel(base, mean, frac, nonlinear, gain):
            z = base + frac * mean
            return gain * (torch.tanh(z) if nonlinear else z)
        sig1 = channel(G1, mu_sh_y, shared_class_frac, nonlinear_shared, alpha_shared) + \
            channel(U1, mu_p1_y, 1.0 - shared_class_frac, nonlinear_specific, beta_specific)
        sig2 = channel(G2, mu_sh_y_mod2, shared_class_frac, nonlinear_shared, alpha_shared) + \
            channel(U2, mu_p2_y, 1.0 - shared_class_frac, nonlinear_specific, beta_specific)
        if d_spurious > 0:
            sp1, sp2 = randn(n, d_spurious), randn(n, d_spurious)
            X1, X2 = torch.cat([sig1, sp1], dim=1), torch.cat([sig2, sp2], dim=1)
        else:
            X1, X2 = sig1, sig2
        if hetero_noise:
            m1 = 1.0 + hetero_scale * (2 * torch.rand(n, 1, generator=g) - 1.0)
            m2 = 1.0 + hetero_scale * (2 * torch.rand(n, 1, generator=g) - 1.0)
            n1 = randn(X1.shape) * noise_std * m1
            n2 = randn(X2.shape) * noise_std * m2
        else:
            n1 = randn(X1.shape) * noise_std
            n2 = randn(X2.shape) * noise_std
        self.X1, self.X2, self.y = X1 + n1, X2 + n2, y
        self.extras = {"G1": G1, "G2": G2, "mu_sh_y": mu_sh_y, "mu_p1_y": mu_p1_y, "mu_p2_y": mu_p2_y}

    def __len__(self):
        return self.X1.shape[0]

    def __getitem__(self, idx):
        return self.X1[idx], self.X2[idx], self.y[idx]


def make_loaders_simple_plus(batch_size=128, **kwargs):
    """datasets/dataset.py:460-471: (dataset, shuffled train loader with drop_last, validation loader)."""
    ds = SimpleTwoModalPlus(**{k: v for k, v in kwargs.items() if k != "val_split"}) if "val_split" in kwargs else \
        SimpleTwoModalPlus(**kwargs)
    n = len(ds)
    n_val = int(kwargs.get("val_split", 0.2) * n)
    tr, va = random_split(ds, [n - n_val, n_val], generator=torch.Generator().manual_seed(kwargs.get("seed", 0)))
    return ds, DataLoader(tr, batch_size=batch_size, shuffle=True, drop_last=True), \
        DataLoader(va, batch_size=batch_size, shuffle=False, drop_last=False)


# --------------------------------------------------------------------------------------------------------------
# device-resident loader
# --------------------------------------------------------------------------------------------------------------
class DeviceLoader:
    """Iterates ``[x_0, .., x_{V-1}, y]`` batches (the DataLoader contract of the reference) out of tensors that
    live on the GPU: the views are uploaded ONCE, every epoch draws a permutation on the device
    (``torch.randperm(generator=...)``) and each batch is one ``index_select`` per view -- no host collation, no
    per-step H2D copy, no worker processes.  ``indices`` restricts it to a subset (train / test split of
    ``run._split_indices``); ``rank`` / ``world_size`` hand every data-parallel rank its row shard of each global
    batch.  Shuffling is distribution-equal to ``DataLoader(shuffle=True)``, not stream-equal."""

    def __init__(self, views: Sequence, labels, batch_size: int, device="cuda", indices: Optional[Sequence[int]] = None,
                 shuffle: bool = False, drop_last: bool = False, seed: int = 0, rank: Optional[int] = None,
                 world_size: Optional[int] = None):
        dev = torch.device(device)
        as_t = lambda a, dt: torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(dt)   # noqa: E731
        self.views: List[torch.Tensor] = [as_t(v, torch.float32).to(dev).contiguous() for v in views]
        self.labels = as_t(labels, torch.int64).to(dev)
        n = self.labels.shape[0]
        self.index = torch.arange(n, device=dev) if indices is None else as_t(indices, torch.int64).to(dev)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        if rank is None or world_size is None:      # default: the torch.distributed process group, if there is one
            from .dp import world as _world
            r, w = _world()
            rank = r if rank is None else rank
            world_size = w if world_size is None else world_size
        self.rank, self.world_size = rank, world_size
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        if self.batch_size % world_size:
            raise ValueError(f"batch size {batch_size} is not divisible by world size {world_size}")

    @classmethod
    def from_dataset(cls, dataset, batch_size, **kw):
        """From a ``MultiViewDataset`` (``.X`` list + ``.Y``) or a ``SimpleTwoModalPlus`` (``.X1, .X2, .y``)."""
        if hasattr(dataset, "X1"):
            return cls([dataset.X1, dataset.X2], dataset.y, batch_size, **kw)
        return cls(dataset.X, dataset.Y, batch_size, **kw)

    def __len__(self):
        n = self.index.shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self.index.shape[0]
        order = self.index[torch.randperm(n, device=self.index.device, generator=self.gen)] if self.shuffle else self.index
        for b in range(len(self)):
            rows = order[b * self.batch_size:(b + 1) * self.batch_size]
            if self.world_size > 1:
                # equal shards on every rank (the all-gathers of the data-parallel step need them): a final partial
                # batch keeps len(rows) // world_size rows per rank and drops the remainder
                per = rows.shape[0] // self.world_size
                if per == 0:
                    continue
                rows = rows[self.rank * per:(self.rank + 1) * per]
            yield [v.index_select(0, rows) for v in self.views] + [self.labels.index_select(0, rows)]
