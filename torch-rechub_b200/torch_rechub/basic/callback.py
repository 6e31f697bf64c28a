"""Early stopping on validation AUC (mirror of reference ``torch_rechub/basic/callback.py:4-33``)."""
import copy


class EarlyStopper(object):
    """Keeps the best weights seen so far and says when patience has run out.

    Args:
        patience (int): tolerated number of consecutive non-improving evaluations.

    Attributes read by the trainers: ``best_auc``, ``best_weights``, ``trial_counter``.
    """

    def __init__(self, patience):
        self.patience, self.trial_counter = patience, 0
        self.best_auc, self.best_weights = 0, None

    def stop_training(self, val_auc, weights):
        """True when ``val_auc`` failed to beat the best one ``patience`` times in a row; a new best snapshots ``weights``."""
        if val_auc > self.best_auc:
            self.best_auc, self.trial_counter, self.best_weights = val_auc, 0, copy.deepcopy(weights)
            return False
        exhausted = self.trial_counter + 1 >= self.patience
        if not exhausted:
            self.trial_counter += 1
        return exhausted
