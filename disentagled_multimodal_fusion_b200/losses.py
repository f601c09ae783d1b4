"""Mirror of the hot-path part of the reference's ``models/losses.py``."""
from __future__ import annotations

from torch import nn

from . import ops


class SupConLoss(nn.Module):
    """models/losses.py:7-101, default path (contrast_mode='all', no labels / mask): symmetric
    cross-view InfoNCE + the two no-grad intra-view diagnostics, computed by the tiled K2 kernels
    without materialising the [2B,2B] logits."""

    def __init__(self, temperature=0.07, contrast_mode="all", base_temperature=0.07, precision="fp32"):
        super().__init__()
        self.temperature = temperature
        self.contrast_mode = contrast_mode
        self.base_temperature = base_temperature
        self.precision = precision

    def forward(self, features, labels=None, mask=None):
        if len(features.shape) < 3:
            raise ValueError('`features` needs to be [bsz, n_views, ...],at least 3 dimensions are required')
        if labels is not None and mask is not None:
            raise ValueError('Cannot define both `labels` and `mask`')
        if labels is not None or mask is not None:
            raise NotImplementedError("the fused kernel covers the unsupervised path the reference uses")
        if self.contrast_mode != "all":
            raise ValueError('Unknown mode: {}'.format(self.contrast_mode))
        if features.shape[1] != 2:
            raise NotImplementedError("the fused kernel covers n_views == 2 (as every reference call site)")
        if len(features.shape) > 3:
            features = features.view(features.shape[0], features.shape[1], -1)
        return self.pair(features[:, 0], features[:, 1])

    def pair(self, z0, z1, unit_norm=False, pre=None, reduce=True, diagnostics=True):
        loss, lx, ly = ops.infonce(z0, z1, self.temperature, self.precision, unit_norm=unit_norm, pre=pre, reduce=reduce,
                                   diagnostics=diagnostics)
        r = self.temperature / self.base_temperature
        if not diagnostics:
            return r * loss, None, None
        return r * loss, r * lx, r * ly


    def multi(self, pairs, pres=None, unit_norm=False, reduce=True, diagnostics=None, precision=None):
        """Several (z0, z1) critic calls in one op: [ncalls, 3] = (loss, loss_x, loss_y) rows (ops.infonce_multi)."""
        out = ops.infonce_multi(pairs, self.temperature, precision or self.precision, unit_norm=unit_norm, pres=pres,
                                reduce=reduce, diagnostics=diagnostics)
        return (self.temperature / self.base_temperature) * out


    def multi_stacked(self, stacks, layout, rows, pres=None, unit_norm=False, reduce=True, diagnostics=None, precision=None):
        """``multi`` on row-stacked inputs (ops.infonce_stacked): call c = rows [ra, ra+rows) of stacks[ia] vs rows
        [rb, rb+rows) of stacks[ib] for layout[c] = (ia, ra, ib, rb)."""
        out = ops.infonce_stacked(stacks, layout, rows, self.temperature, precision or self.precision, unit_norm=unit_norm,
                                  pres=pres, reduce=reduce, diagnostics=diagnostics)
        return (self.temperature / self.base_temperature) * out


def ortho_loss(z1, zs, norm=True, temp=0.1):
    if not norm:
        raise NotImplementedError('Please set norm=True')
    return ops.ortho_loss(z1, zs)


class AvgTrustedLoss(nn.Module):
    """models/losses.py:209-248.  ``forward`` keeps the reference signature; the aggregation rule is
    irrelevant to the value (the reference drops the fused-evidence term)."""

    def __init__(self, num_views: int, annealing_start=50, gamma=1):
        super().__init__()
        self.num_views = num_views
        self.annealing_step = 0
        self.annealing_start = annealing_start
        self.gamma = gamma

    def forward(self, evidences, target, evidence_a=None, fused=1, **kwargs):
        loss, _, _ = ops.edl_fused_loss(evidences, target, "cml", self.annealing_step, self.annealing_start,
                                        fused=fused, gamma=self.gamma, global_batch=kwargs.get("global_batch"))
        return loss

    def fused_forward(self, evidences, target, agg, fused=1, global_batch=None):
        """One kernel pass: (loss, fused evidence)."""
        loss, fe, _ = ops.edl_fused_loss(evidences, target, agg, self.annealing_step, self.annealing_start,
                                         fused=fused, gamma=self.gamma, global_batch=global_batch)
        return loss, fe
