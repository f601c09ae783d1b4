"""Mirror of the evaluation reducers of the reference's ``analysis.py``.

``evaluate_subjective_model`` (analysis.py:5-192) and ``evaluate_subjective_model_with_shared`` (:194-399) keep
their signatures and the exact layout of the returned dict; the per-batch work (≈ 40 small torch kernels and a
dozen ``.item()`` host syncs per batch in the reference) is ONE launch of ``dmf_eval_reduce`` into device
accumulators, read back once after the loader loop."""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import ops


def _collect(model, test_loader, device):
    model.eval()
    dev = device or next(model.parameters()).device
    K = getattr(model, "num_classes", None) or getattr(model, "num_labels", None)
    if K is None:
        raise ValueError("Could not infer num_classes from model; set model.num_classes.")
    K = int(K)
    acc = None
    for batch in test_loader:
        if not isinstance(batch, (list, tuple)):
            raise ValueError("Batch must be a (inputs..., labels) tuple/list.")
        inputs = [b.to(dev).float() for b in batch[:-1]]
        target = batch[-1].to(dev)
        if hasattr(model, "shared_step"):
            _, fused_ev, target_out, evidences = model.shared_step([*inputs, target])
        else:
            ev_list = model(inputs)
            evidences = torch.stack(ev_list, dim=1)
            aggregator = getattr(model, "agg", None) or getattr(model, "aggregation", None)
            if aggregator is None:
                raise ValueError("Model must expose an 'aggregation' function/attr.")
            fused_ev = aggregator(evidences)
            target_out = target
        acc = ops.eval_reduce(evidences, fused_ev, target_out, acc)
    if acc is None:
        raise ValueError("empty test_loader")
    out = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in acc.items()}     # the only device->host read
    out["K"] = K
    return out


def _block(st, N):
    """reduce_block of analysis.py:157-169 from one row of the stats accumulator."""
    correct, ev, epi, ale, inc_n, inc_ev, inc_epi, inc_ale = [float(x) for x in st]
    return {
        "accuracy": (correct / N) if N > 0 else 0.0,
        "evidence_mean": (ev / N) if N > 0 else 0.0,
        "epistemic_mean": (epi / N) if N > 0 else 0.0,
        "aleatoric_mean": (ale / N) if N > 0 else 0.0,
        "incorrect_only": {
            "evidence_mean": (inc_ev / inc_n) if inc_n > 0 else 0.0,
            "epistemic_mean": (inc_epi / inc_n) if inc_n > 0 else 0.0,
            "aleatoric_mean": (inc_ale / inc_n) if inc_n > 0 else 0.0,
        },
    }


@torch.no_grad()
def evaluate_subjective_model(model, test_loader, device: Optional[torch.device] = None) -> Dict[str, Any]:
    """analysis.py:5-192: per-view and fused accuracy / evidence / epistemic / aleatoric means, the same over the
    misclassified samples only, and per-class mean evidence (unconditional and for the true class)."""
    a = _collect(model, test_loader, device)
    N, V = a["N"], a["stats"].shape[0] - 1
    counts = torch.clamp(a["class_counts"], min=1e-12)
    uncond = a["class_sum"] / max(N, 1)
    true = a["true_sum"] / counts
    return {
        "per_view": [_block(a["stats"][v], N) for v in range(V)],
        "fused": _block(a["stats"][V], N),
        "per_class_evidence": {
            "unconditional": {"per_view": [uncond[v].tolist() for v in range(V)], "fused": uncond[V].tolist()},
            "true_class": {"per_view": [true[v].tolist() for v in range(V)], "fused": true[V].tolist()},
        },
    }


@torch.no_grad()
def evaluate_subjective_model_with_shared(model, test_loader, device: Optional[torch.device] = None) -> Dict[str, Any]:
    """analysis.py:194-399: evidences = [shared, view_0, ..., view_{N-1}]; the shared head is reported separately."""
    a = _collect(model, test_loader, device)
    N, V = a["N"], a["stats"].shape[0] - 1
    if V < 2:
        raise ValueError("Expected at least one shared and one specific view (V >= 2).")
    counts = torch.clamp(a["class_counts"], min=1e-12)
    uncond = a["class_sum"] / max(N, 1)
    true = a["true_sum"] / counts
    return {
        "shared": _block(a["stats"][0], N),
        "per_view": [_block(a["stats"][v], N) for v in range(1, V)],
        "fused": _block(a["stats"][V], N),
        "per_class_evidence": {
            "unconditional": {"shared": uncond[0].tolist(), "per_view": [uncond[v].tolist() for v in range(1, V)],
                              "fused": uncond[V].tolist()},
            "true_class": {"shared": true[0].tolist(), "per_view": [true[v].tolist() for v in range(1, V)],
                           "fused": true[V].tolist()},
        },
    }
