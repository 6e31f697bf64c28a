"""Early stopping on validation AUC (mirror of reference ``torch_rechub/basic/callback.py:4-33``)."""
import copy


class EarlyStopper(object):
    """Keeps the best weights seen so far and says when patience has run out.

    Args:
        patience (int): tolerated number of consecutive non-improving evaluations.
    """

    def __init__(self, patience):
        self.patience = patience
        self.trial_counter = 0
        self.best_auc = 0
        self.best_weights = None

    def stop_training(self, val_auc, weights):
        """Return True when training should stop (reference ``callback.py:17-33``)."""
        improved = val_auc > self.best_auc
        if improved:
            self.best_auc, self.trial_counter = val_auc, 0
            self.best_weights = copy.deepcopy(weights)
            return False
        if self.trial_counter + 1 < self.patience:
            self.trial_counter += 1
            return False
        return True
