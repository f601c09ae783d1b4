"""Mirror of the reference's ``models/evidential_probe.py``: frozen backbone -> per-view evidential
heads (one grouped launch per layer over all V heads) -> fused K3 kernel (aggregation + AvgTrustedLoss
+ gradient in one pass).  ``shared_step`` returns the reference 4-tuple
``(loss, evidences_a, labels, evidences)`` consumed by analysis.py:63,262."""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from . import dp, ops
from .classifiers import EvidentialNN, grouped_forward
from .lightning import Accuracy, LightningModule
from .losses import AvgTrustedLoss
from .optim import make_optimizer
from .utils import get_avg_fusion, get_cml_fusion, get_disentangled_fusion, get_joint_fusion


def _strip_label(batch):
    xs = batch
    if isinstance(batch, (list, tuple)) and len(batch) > 0:
        last = batch[-1]
        if torch.is_tensor(last) and last.dtype in (torch.int64, torch.int32) and last.ndim >= 1:
            xs = batch[:-1]
    return xs


class _ProbeBase(LightningModule):
    def _init_metrics(self, num_classes):
        mk = lambda: Accuracy(task='multiclass', num_classes=num_classes)
        self.train_acc, self.val_acc, self.test_acc = mk(), mk(), mk()
        self.train_modality_accs = nn.ModuleList([mk() for _ in range(self.num_views)])
        self.val_modality_accs = nn.ModuleList([mk() for _ in range(self.num_views)])
        self.test_modality_accs = nn.ModuleList([mk() for _ in range(self.num_views)])
        self.validation_step_outputs = []
        self.test_step_outputs = []

    def freeze_backbone(self):
        for p in self.backbone.parameters():
            p.requires_grad = False
        self.backbone.eval()

    def _heads(self):
        raise NotImplementedError

    def _fused_step(self, evidences_list, labels, fused=1):
        evidences = torch.stack(evidences_list, dim=1)                  # (B, V, C)
        # data parallel: the kernel normalises by the GLOBAL batch (this rank's loss is then a partial sum and the
        # SUM all-reduce of the gradients gives the single-process gradient); the returned VALUE is the global loss
        _, ws = dp.world()
        gb = evidences.shape[0] * ws if ws > 1 else None
        loss, evidences_a = self.criterion.fused_forward(evidences, labels, self.agg_name, fused=fused, global_batch=gb)
        if ws > 1:
            loss = dp.globalize_sum_loss(loss)
        return loss, evidences_a, labels, evidences

    # ----------------- Training / eval steps (models/evidential_probe.py:106-203)
    def training_step(self, batch, batch_idx=None):
        loss, evidences_a, target, evidences = self.shared_step(batch)
        self.log('train_loss', loss, on_step=False, on_epoch=True, prog_bar=True)
        acc = self.train_acc(evidences_a, target)
        self.log('train_acc_step', acc, on_step=False, on_epoch=True, prog_bar=True)
        for i, modality_acc in enumerate(self.train_modality_accs):
            modality_acc.update(torch.argmax(evidences[:, i, :], dim=1), target)
            self.log(f'train_acc_modality_{i}_step', modality_acc.compute(), prog_bar=False)
        return loss

    def on_train_epoch_end(self):
        self.log('train_acc', self.train_acc.compute(), prog_bar=True)
        self.criterion.annealing_step += 1
        for i, modality_acc in enumerate(self.train_modality_accs):
            self.log(f'train_acc_modality_{i}', modality_acc.compute(), on_step=False, prog_bar=False)
            modality_acc.reset()

    def _eval_step(self, batch, acc, maccs, outputs):
        loss, evidences_a, target, evidences = self.shared_step(batch)
        # uncertainty summaries (models/evidential_probe.py:139-143) from the forward-only kernel pass
        _, u, ale, pred = ops.edl_summaries(evidences, target, self.agg_name)
        acc.update(pred[:, -1].long(), target)
        for i, modality_acc in enumerate(maccs):
            modality_acc.update(pred[:, i].long(), target)
        outputs.append({'loss': loss.detach(), 'entropy': u.unsqueeze(-1), 'aleatoric': ale})

    def validation_step(self, batch, batch_idx):
        self._eval_step(batch, self.val_acc, self.val_modality_accs, self.validation_step_outputs)

    def test_step(self, batch, batch_idx, dataloader_idx=0):
        self._eval_step(batch, self.test_acc, self.test_modality_accs, self.test_step_outputs)

    def on_validation_epoch_end(self):
        outs = self.validation_step_outputs
        self.log('val_loss', torch.stack([x['loss'] for x in outs]).mean(), on_step=False, prog_bar=True)
        self.log('val_entropy', torch.cat([x['entropy'] for x in outs]).mean(), on_step=False, prog_bar=True)
        self.log('val_sigma', torch.cat([x['aleatoric'] for x in outs]).mean(), on_step=False, prog_bar=True)
        self.log('val_acc', self.val_acc.compute(), on_step=False, prog_bar=True)
        for i, modality_acc in enumerate(self.val_modality_accs):
            self.log(f'val_acc_modality_{i}', modality_acc.compute(), on_step=False, prog_bar=False)
            modality_acc.reset()
        self.val_acc.reset()
        outs.clear()

    def on_test_epoch_end(self):
        outs = self.test_step_outputs
        self.log('test_acc', self.test_acc.compute(), prog_bar=True)
        self.log('test_entropy_epi', torch.cat([x['entropy'] for x in outs]).mean())
        self.log('test_ale', torch.cat([x['aleatoric'] for x in outs]).mean())
        for i, modality_acc in enumerate(self.test_modality_accs):
            self.log(f'test_acc_modality_{i}', modality_acc.compute(), on_step=False, prog_bar=False)
            modality_acc.reset()
        self.test_acc.reset()
        outs.clear()


class EvidentialProbeModule(_ProbeBase):
    """models/evidential_probe.py:11-212 (1 shared + N specific heads)."""

    def __init__(self, backbone, num_classes, input_dim, hidden_dim=(32), lr=1e-4, dropout=0.3, annealing_start=20,
                 optimizer=torch.optim.Adam, freeze_backbone=True, aggregation='cml', fused=1, precision="fp32"):
        super().__init__()
        self.backbone = copy.deepcopy(backbone)
        if not hasattr(self.backbone, 'N'):
            raise ValueError("backbone must expose attribute 'N' (number of modalities).")
        self.N = int(self.backbone.N)
        self.fused = fused
        self.num_views = 1 + self.N
        self.num_classes = num_classes
        self.lr = lr
        self.optimizer = optimizer
        self.annealing_start = annealing_start
        self.agg = {'cml': get_cml_fusion, 'avg': get_avg_fusion, 'joint': get_joint_fusion,
                    'disentangled': get_disentangled_fusion}[aggregation]
        self.agg_name = aggregation
        # the shared head is sized from the backbone's actual shared width
        # (DisentangledSSL.get_embedding returns cat([zsx1, zsx2]) of width 2*embed_dim)
        shared_dim = getattr(self.backbone, 'shared_embedding_dim', input_dim)
        self.x_shared = EvidentialNN(dropout=dropout, output_dims=num_classes, layers=(shared_dim, *hidden_dim))
        self.x_specs = nn.ModuleList([
            EvidentialNN(dropout=dropout, output_dims=num_classes, layers=(input_dim, *hidden_dim))
            for _ in range(self.N)])
        # "bf16": hidden layers of the heads on the tensor cores, evidence layer in fp32 (ops.grouped_mlp)
        for head in [self.x_shared, *self.x_specs]:
            head.precision = precision
        self._init_metrics(num_classes)
        self.criterion = AvgTrustedLoss(num_views=self.num_views, annealing_start=annealing_start)
        if freeze_backbone:
            self.freeze_backbone()

    @torch.no_grad()
    def get_embedding(self, batch):
        return self.backbone.get_embedding(_strip_label(batch))

    def forward(self, x):
        Zc, Zp_list = self.get_embedding(x)
        heads = [self.x_shared] + list(self.x_specs)
        return grouped_forward(heads, [Zc] + list(Zp_list))

    def shared_step(self, batch):
        return self._fused_step(self(batch), batch[-1], fused=self.fused)

    def configure_optimizers(self):
        # trainable parameters only (the frozen backbone copy stays out of the flat buffer); AdamW -> fused AdamW on CUDA
        optimizer = make_optimizer(torch.optim.AdamW, [p for p in self.parameters() if p.requires_grad], lr=self.lr, weight_decay=1e-4)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=self.trainer.max_epochs, eta_min=1e-6)
        return {'optimizer': optimizer, 'lr_scheduler': scheduler, 'monitor': 'val_loss'}


class DisentangledEvidentialProbeModule(_ProbeBase):
    """models/evidential_probe.py:214-408 (private heads only)."""

    def __init__(self, backbone, num_classes, input_dim, hidden_dim=(32), lr=1e-4, dropout=0.3,
                 annealing_start=20, optimizer=torch.optim.AdamW, freeze_backbone=True, aggregation='cml'):
        super().__init__()
        self.backbone = copy.deepcopy(backbone)
        if not hasattr(self.backbone, 'N'):
            raise ValueError("backbone must expose attribute 'N' (number of modalities).")
        self.N = int(self.backbone.N)
        self.num_classes = num_classes
        self.lr = lr
        self.optimizer = optimizer
        self.annealing_start = annealing_start
        agg_map = {'cml': get_cml_fusion, 'avg': get_avg_fusion}
        if aggregation not in agg_map:
            raise ValueError(f"aggregation must be one of {list(agg_map.keys())}")
        self.agg = agg_map[aggregation]
        self.agg_name = aggregation
        self.spec_heads = nn.ModuleList([
            EvidentialNN(dropout=dropout, output_dims=num_classes, layers=(input_dim, *hidden_dim))
            for _ in range(self.N)])
        self.num_views = self.N
        self._init_metrics(num_classes)
        self.criterion = AvgTrustedLoss(num_views=self.num_views, annealing_start=annealing_start)
        if freeze_backbone:
            self.freeze_backbone()

    @torch.no_grad()
    def get_embedding(self, batch):
        _, Zp_list = self.backbone.get_embedding(_strip_label(batch))
        return Zp_list

    def forward(self, x):
        return grouped_forward(list(self.spec_heads), list(self.get_embedding(x)))

    def shared_step(self, x):
        return self._fused_step(self(x), x[-1], fused=1)

    def configure_optimizers(self):
        optimizer = make_optimizer(self.optimizer, [p for p in self.parameters() if p.requires_grad], lr=self.lr)
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.1, patience=5)
        return {'optimizer': optimizer, 'lr_scheduler': scheduler, 'monitor': 'val_loss'}
