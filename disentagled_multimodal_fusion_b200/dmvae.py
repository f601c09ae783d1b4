"""Mirror of the reference's ``models/dmvae.py`` (N-view DMVAE) on the B200 kernels.

Same constructor, attributes (``N``, ``x_dims``, ``encoders``, ``decoders``, ``feature_encoders``),
state_dict keys and ``forward(x_list) -> (loss, logs)`` / ``get_embedding`` surface.  One forward is
3 grouped-GEMM launches for the N encoders, 1 head kernel (chunk + reparameterise + PoE + 3 KLs +
decoder-input packing), 3 grouped-GEMM launches for all N^2 decoder passes, and N MSE kernels --
instead of ~9k ATen ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import dp, ops
from ._lib import lib, check, ptr, stream, ptr_array, require_device
from .classifiers import IdentityEncoder, Linear, grouped_forward
from .lightning import LightningModule
from .optim import make_optimizer


class DMVAE(LightningModule):
    def __init__(self, feature_encoders=None, output_dim=[100, 100], dropout=0., a=1.0,
                 optimizer=torch.optim.Adam, hidden_dim=512, embed_dim=100, lr=1e-4, initialization='xavier',
                 num_epochs=50, poe_temperature=1.5, cross_weight=1.0, lambda_per_modality=None,
                 precision="fp32"):
        super().__init__()
        self.save_hyperparameters(ignore=['feature_encoders'])
        self.num_epochs = num_epochs
        self.optimizer_cls = optimizer
        self.lr = lr
        self.a = a
        assert isinstance(output_dim, (list, tuple)) and len(output_dim) >= 2, \
            "output_dim must be a list of per-modality input dims (N >= 2)."
        self.N = len(output_dim)
        self.x_dims = [int(d) for d in output_dim]
        self.hidden_dim = hidden_dim
        self.embed_dim = embed_dim
        self.poe_temperature = poe_temperature
        self.cross_weight = cross_weight
        self.lambda_per_modality = lambda_per_modality or [1.0] * self.N
        self.precision = precision
        if feature_encoders is not None:
            assert len(feature_encoders) == self.N, "feature_encoders length must equal number of modalities."
            self.feature_encoders = nn.ModuleList([ctor(**kwargs) for ctor, kwargs in feature_encoders])
        else:
            self.feature_encoders = nn.ModuleList([IdentityEncoder() for _ in range(self.N)])
        self.encoders = nn.ModuleList([
            Linear(layers=(self.x_dims[i], hidden_dim, hidden_dim), output_dims=4 * embed_dim,
                   initialization=initialization, dropout=dropout) for i in range(self.N)])
        self.decoders = nn.ModuleList([
            Linear(layers=(2 * embed_dim, hidden_dim, hidden_dim), output_dims=self.x_dims[i],
                   initialization=initialization, dropout=dropout) for i in range(self.N)])

    # ---------- helpers
    def _encode(self, x_list):
        feats = [self.feature_encoders[i](x_list[i]).float() for i in range(self.N)]
        stats = grouped_forward(list(self.encoders), feats, precision=self.precision)
        return feats, stats

    def draw_noise(self, B, device):
        """2N+1 sequential randn draws [B,e] in the reference's order (models/dmvae.py:147-150:
        z_p views, z_s_uni views, z_s) so the same seed gives the same noise as the reference on
        the same device."""
        return torch.stack([torch.randn(B, self.embed_dim, device=device) for _ in range(2 * self.N + 1)], dim=0)

    @torch.no_grad()
    def get_embedding(self, x_list, return_poe: bool = True):
        require_device()
        _, stats = self._encode(x_list)
        e = self.embed_dim
        mu_p_all = [s[:, 2 * e:3 * e] for s in stats]
        if return_poe:
            B = stats[0].shape[0]
            mu = torch.empty(B, e, dtype=torch.float32, device=stats[0].device)
            check(lib.dmf_dmvae_poe_mean(ptr_array(stats), self.N, B, e, float(self.poe_temperature), ptr(mu), stream()))
            return mu, mu_p_all
        return torch.cat([s[:, :e] for s in stats], dim=1), mu_p_all

    # ---------- core forward (models/dmvae.py:128-188)
    def forward(self, x_list, noise=None):
        N, e = self.N, self.embed_dim
        B = x_list[0].shape[0]
        feats, stats = self._encode(x_list)
        if noise is None:
            noise = self.draw_noise(B, stats[0].device)
        elif isinstance(noise, (list, tuple)):
            noise = torch.stack(list(noise), dim=0)
        # PoE temperature is the hard-coded 1.5 of models/dmvae.py:149 (not self.poe_temperature)
        outs = ops._DmvaeHead.apply((N, B, e, 1.5), noise, *stats)
        dec_in, kl3 = list(outs[:N]), outs[N]
        recon = grouped_forward(list(self.decoders), dec_in, precision=self.precision)
        lam = self.lambda_per_modality
        pairs = N * (N - 1)
        loss_recon_joint = 0.0
        loss_recon_cross = 0.0
        for i in range(N):
            o2 = ops._DmvaeMse.apply((N, B, self.x_dims[i], i, float(lam[i]),
                                      float(lam[i]) * self.cross_weight / max(pairs, 1)), recon[i], feats[i])
            loss_recon_joint = loss_recon_joint + o2[0]
            loss_recon_cross = loss_recon_cross + o2[1]
        kl_private, kl_shared_poe, kl_shared_uni = kl3[0], kl3[1], kl3[2]
        loss_joint = loss_recon_joint + self.a * (kl_private + N * kl_shared_poe)
        loss_cross = loss_recon_cross + self.a * kl_shared_uni
        loss = loss_joint + loss_cross
        # data parallel: every term is a mean over this rank's rows; value -> global mean, gradient -> local / world
        # (dp.FlatParams.allreduce_grads SUMs), logs -> global means
        if dp.world()[1] > 1:
            zero = torch.zeros((), device=loss.device)
            loss, (loss_recon_joint, loss_recon_cross, kl_private, kl_shared_poe, kl_shared_uni) = dp.globalize_mean_losses(
                loss, [loss_recon_joint, loss_recon_cross if pairs > 0 else zero, kl_private, kl_shared_poe, kl_shared_uni])
        # device scalars: one host sync when the caller reads them, not five float() calls per step
        logs = {'loss': loss.detach(), 'loss_joint_recon': loss_recon_joint.detach(),
                'loss_cross_recon': loss_recon_cross.detach() if pairs > 0 else 0.0,
                'kl_private': kl_private.detach(), 'kl_shared_poe': kl_shared_poe.detach(),
                'kl_shared_uni_sum': kl_shared_uni.detach(), 'a': float(self.a), 'N': self.N}
        return loss, logs

    # ---------- Lightning plumbing (models/dmvae.py:191-210)
    def training_step(self, batch, batch_idx):
        xs = [b.float() for b in batch[:-1]]
        loss, logs = self(xs)
        self.log('train/loss', logs['loss'], on_step=False, on_epoch=True, prog_bar=True)
        for k in ('loss_joint_recon', 'loss_cross_recon', 'kl_private', 'kl_shared_poe', 'kl_shared_uni_sum'):
            self.log('train/' + k, logs[k], on_epoch=True, prog_bar=True)
        return loss

    def configure_optimizers(self):
        opt = make_optimizer(self.optimizer_cls, self.parameters(), lr=self.lr)   # Adam -> fused flat-buffer Adam on CUDA
        sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=self.num_epochs, eta_min=0, last_epoch=-1)
        return {'optimizer': opt, 'lr_scheduler': {'scheduler': sch, 'interval': 'epoch', 'monitor': 'train/loss'}}
