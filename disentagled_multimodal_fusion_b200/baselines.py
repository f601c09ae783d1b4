"""Mirror of ``LateFusion`` from the reference's ``models/baselines.py`` (:12-150): per-view feature
encoder + evidential head, aggregation in {cml, avg, dbf}, same fused K3 loss kernel."""
from __future__ import annotations

import torch
from torch import nn

from .classifiers import EvidentialNN, grouped_forward
from .evidential_probe import _ProbeBase
from .losses import AvgTrustedLoss
from .optim import make_optimizer
from .utils import discounted_belief_fusion, get_avg_fusion, get_cml_fusion


class LateFusion(_ProbeBase):
    def __init__(self, feature_encoders, output_dims=[100, 100], num_classes=42, dropout=0.3, aggregation='cml',
                 lr=1e-4, annealing_start=20, hidden_dim=(32), optimizer=torch.optim.Adam, weight_decay=1e-5, fused=1):
        super().__init__()
        self.num_classes = num_classes
        self.lr = lr
        self.optimizer = optimizer
        self.fused = fused
        self.feature_encoders = nn.ModuleList([i[0](**i[1]) for i in feature_encoders])
        self.weight_decay = weight_decay
        self.heads = nn.ModuleList([
            EvidentialNN(dropout=dropout, output_dims=num_classes, layers=(int(output_dims[i]), *hidden_dim))
            for i in range(len(feature_encoders))])
        self.aggregation = {'cml': get_cml_fusion, 'avg': get_avg_fusion, 'dbf': discounted_belief_fusion}[aggregation]
        self.agg_name = aggregation
        self.num_views = len(feature_encoders)
        self._init_metrics(num_classes)
        self.criterion = AvgTrustedLoss(num_views=self.num_views, annealing_start=annealing_start)
        self.aleatoric_uncertainties = None
        self.epistemic_uncertainties = None

    def forward(self, inputs):
        feats = [fe(inputs[i].float()) for i, fe in enumerate(self.feature_encoders)]
        return grouped_forward(list(self.heads), feats)

    def shared_step(self, batch):
        return self._fused_step(self(batch), batch[-1], fused=self.fused)

    def on_train_epoch_end(self):
        self.log('train_acc', self.train_acc.compute(), prog_bar=True)
        self.criterion.annealing_step += 1

    def configure_optimizers(self):
        optimizer = make_optimizer(self.optimizer, self.parameters(), lr=self.lr)
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.1, patience=10)
        return {'optimizer': optimizer, 'lr_scheduler': scheduler, 'monitor': 'val_loss'}
