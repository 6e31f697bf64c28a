"""Early stopping on validation AUC (mirror of reference ``torch_rechub/basic/callback.py:4-33``)."""
import copy

import torch


class EarlyStopper(object):
    """Keeps the best weights seen so far and says when patience has run out.

    Args:
        patience (int): tolerated number of consecutive non-improving evaluations.

    Attributes read by the trainers: ``best_auc``, ``best_weights``, ``trial_counter``.

    The reference deep-copies the whole ``state_dict`` at every new best (``callback.py:27``) — 1.66 GB of tables at Criteo shape,
    re-allocated each time.  Here the snapshot lives in buffers allocated ONCE (first new best) and refreshed in place with
    asynchronous same-device copies on the current stream (CUDA tables: ~0.6 ms of HBM traffic for 1.66 GB, no host memory, no
    allocator churn); the snapshot is still an independent copy with the reference's keys, so ``load_state_dict(best_weights)``
    behaves the same.
    """

    def __init__(self, patience):
        self.patience, self.trial_counter = patience, 0
        self.best_auc, self.best_weights = 0, None

    def _snapshot(self, weights):
        if not isinstance(weights, dict) or not all(torch.is_tensor(v) for v in weights.values()):
            return copy.deepcopy(weights)
        old = self.best_weights
        reuse = isinstance(old, dict) and old.keys() == weights.keys() and all(
            torch.is_tensor(old[k]) and old[k].shape == v.shape and old[k].dtype == v.dtype and old[k].device == v.device for k, v in weights.items())
        if not reuse:
            return type(weights)((k, v.detach().clone()) for k, v in weights.items())
        with torch.no_grad():
            for k, v in weights.items():
                old[k].copy_(v, non_blocking=True)
        return old

    def stop_training(self, val_auc, weights):
        """True when ``val_auc`` failed to beat the best one ``patience`` times in a row; a new best snapshots ``weights``."""
        if val_auc > self.best_auc:
            self.best_auc, self.trial_counter, self.best_weights = val_auc, 0, self._snapshot(weights)
            return False
        exhausted = self.trial_counter + 1 >= self.patience
        if not exhausted:
            self.trial_counter += 1
        return exhausted
