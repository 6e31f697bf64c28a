"""Trainer for matching / retrieval models (mirror of reference ``torch_rechub/trainers/match_trainer.py:12-258``).

Same constructor, defaults, loss selection and methods.  The in-batch-negative branch (reference ``:118-140``) builds the
``B x B`` score matrix of the two towers, samples negatives for every row (``utils.match.inbatch_negative_sampling``: one
batched draw on CUDA instead of a Python loop over the batch) and trains with cross entropy (or BPR) on the gathered logits.
CUDA tensors take the fused gather / tensor-core tower route of the engine inside the model; the trainer itself adds no
device-specific code besides checking the engine's out-of-range-id flag where the reference already syncs (``loss.item()``).
"""
import os

import torch
import tqdm
from sklearn.metrics import roc_auc_score

from ..basic.callback import EarlyStopper
from ..basic.loss_func import BPRLoss, RegularizationLoss
from ..utils.match import gather_inbatch_logits, inbatch_negative_sampling


def _unwrap(model):
    return model.module if isinstance(model, torch.nn.DataParallel) else model


class MatchTrainer(object):
    """A general trainer for matching / retrieval.

    Args:
        model (nn.Module): any matching model.
        mode (int): ``0`` point-wise (BCE), ``1`` pair-wise (BPR), ``2`` list-wise (softmax cross entropy).
        in_batch_neg (bool): draw the negatives from the other samples of the batch (needs ``user_tower`` / ``item_tower``).
        in_batch_neg_ratio (int): negatives per positive in that mode (default: all ``B - 1``).
        hard_negative (bool): take the highest-scoring in-batch negatives instead of uniform ones.
        sampler_seed (int): seed of the in-batch sampler's private generator.
        optimizer_fn, optimizer_params, regularization_params, scheduler_fn, scheduler_params, n_epoch,
        earlystop_patience, device, gpus, model_path, model_logger: as ``CTRTrainer``.
    """

    def __init__(
        self,
        model,
        mode=0,
        in_batch_neg=False,
        in_batch_neg_ratio=None,
        hard_negative=False,
        sampler_seed=None,
        optimizer_fn=torch.optim.Adam,
        optimizer_params=None,
        regularization_params=None,
        scheduler_fn=None,
        scheduler_params=None,
        n_epoch=10,
        earlystop_patience=10,
        device="cpu",
        gpus=None,
        model_path="./",
        model_logger=None,
    ):
        self.model = model
        self.gpus = [] if gpus is None else gpus
        self.device = torch.device(device)
        if len(self.gpus) > 1:
            if self.device.type == "cuda":
                # nn.DataParallel replicas see the tables as non-leaf Broadcast outputs: the engine's scatter-add lands in the
                # replica's gradient buffer and is never reduced back, so the tables would silently stop training while the
                # towers keep going (the same configuration CTRTrainer rejects)
                raise RuntimeError("multi-GPU training of the two-tower path is not wired to the field-sharded engine yet: "
                                   "run MatchTrainer on one CUDA device (gpus=[i]) — nn.DataParallel cannot carry the engine's table gradients")
            print('parallel running on these gpus:', self.gpus)
            self.model = torch.nn.DataParallel(self.model, device_ids=self.gpus)
        self.model.to(self.device)
        self.in_batch_neg = in_batch_neg
        self.in_batch_neg_ratio = in_batch_neg_ratio
        self.hard_negative = hard_negative
        self._sampler_generator = None
        if sampler_seed is not None:
            self._sampler_generator = torch.Generator(device=self.device)
            self._sampler_generator.manual_seed(sampler_seed)
        if in_batch_neg:
            base = _unwrap(model)
            if not hasattr(base, 'user_tower') or not hasattr(base, 'item_tower'):
                raise ValueError(f"Model {type(base).__name__} does not support in-batch negative sampling. "
                                 "Only two-tower models with user_tower() and item_tower() methods are supported, "
                                 "such as DSSM, YoutubeDNN, MIND, GRU4Rec, SINE, ComiRec, SASRec, NARM, STAMP, etc.")
        if optimizer_params is None:
            optimizer_params = {"lr": 1e-3, "weight_decay": 1e-5}
        if regularization_params is None:
            regularization_params = {"embedding_l1": 0.0, "embedding_l2": 0.0, "dense_l1": 0.0, "dense_l2": 0.0}
        self.mode = mode
        if mode == 0:  # with in-batch negatives the point-wise mode becomes a classification over the sampled columns
            self.criterion = torch.nn.CrossEntropyLoss() if in_batch_neg else torch.nn.BCELoss()
        elif mode == 1:
            self.criterion = BPRLoss()
        elif mode == 2:
            self.criterion = torch.nn.CrossEntropyLoss()
        else:
            raise ValueError("mode only contain value in %s, but got %s" % ([0, 1, 2], mode))
        self.optimizer = optimizer_fn(self.model.parameters(), **optimizer_params)
        self.scheduler = scheduler_fn(self.optimizer, **scheduler_params) if scheduler_fn is not None else None
        self.evaluate_fn = roc_auc_score
        self.n_epoch = n_epoch
        self.early_stopper = EarlyStopper(patience=earlystop_patience)
        self.model_path = model_path
        self.reg_loss_fn = RegularizationLoss(**regularization_params)
        self.model_logger = model_logger

    # ---------------------------------------------------------------------------------------------
    def _to_device(self, x_dict):
        return {k: v.to(self.device) for k, v in x_dict.items()}

    def _inbatch_loss(self, x_dict):
        """Two towers -> (B, B) scores -> sampled negatives -> loss on ``[positive | negatives]`` logits (reference ``:118-140``)."""
        base = _unwrap(self.model)
        towers = []
        for emb in (base.user_tower(x_dict), base.item_tower(x_dict)):
            if emb is None:
                raise ValueError("Model must return user/item embeddings when in_batch_neg is True.")
            if emb.dim() > 2 and emb.size(1) == 1:
                emb = emb.squeeze(1)
            towers.append(emb)
        user, item = towers
        if user.dim() != 2 or item.dim() != 2:
            raise ValueError(f"In-batch negative sampling requires 2D embeddings, got shapes {user.shape} and {item.shape}")
        if user.is_cuda and self.mode != 1:
            fused = self._inbatch_loss_cuda(user, item)
            if fused is not None:
                return fused
        scores = torch.matmul(user, item.t())
        picks = inbatch_negative_sampling(scores, neg_ratio=self.in_batch_neg_ratio, hard_negative=self.hard_negative, generator=self._sampler_generator)
        logits = gather_inbatch_logits(scores, picks)
        if self.mode == 1:
            return self.criterion(logits[:, 0], logits[:, 1:], in_batch_neg=True)
        return self.criterion(logits, torch.zeros(logits.size(0), dtype=torch.long, device=self.device))  # the positive is column 0

    def _inbatch_loss_cuda(self, user, item):
        """The same loss on the engine's kernels: the (B, B) scores only when hard negatives need them (tensor-core GEMM), one
        sampling launch for all rows, cross entropy + its backward straight from the tower outputs (no gathered logits, no dense
        score gradient).  None when a shape is outside the kernels (the composition above runs then)."""
        from ..b200 import ops
        n = user.size(0)
        if n <= 1 or item.size(0) != n:
            return None
        k = self.in_batch_neg_ratio if (self.in_batch_neg_ratio is not None and 0 < self.in_batch_neg_ratio <= n - 1) else n - 1
        seed = None
        scores = user.detach()
        if self.hard_negative:
            scores = ops.scores_nt(user.detach(), item.detach())
            if scores is None:
                scores = torch.matmul(user.detach(), item.detach().t())
        else:
            seed = torch.randint(0, 2**62, (1,), device=user.device, generator=self._sampler_generator)  # advances the sampler's stream
        picks = ops.inbatch_sample(scores if self.hard_negative else user, k, self.hard_negative, seed)
        if picks is None:
            return None
        self.last_picks = picks
        return ops.inbatch_cross_entropy(user, item, picks)

    def _loss(self, x_dict, y):
        if self.in_batch_neg:
            loss = self._inbatch_loss(x_dict)
        elif self.mode == 1:
            pos_score, neg_score = self.model(x_dict)
            loss = self.criterion(pos_score, neg_score)
        else:
            loss = self.criterion(self.model(x_dict), y)
        return loss + self.reg_loss_fn(self.model)

    def train_one_epoch(self, data_loader, log_interval=10):
        self.model.train()
        window, total, count = 0.0, 0.0, 0
        tk0 = tqdm.tqdm(data_loader, desc="train", smoothing=0, mininterval=1.0)
        for x_dict, y in tk0:
            x_dict = self._to_device(x_dict)
            y = y.to(self.device)
            y = y.float() if self.mode == 0 else y.long()
            loss = self._loss(x_dict, y)
            self.model.zero_grad()
            loss.backward()
            self.optimizer.step()
            value = loss.item()
            if self.device.type == "cuda":
                from ..b200 import _lib
                _lib.check_errors(self.device)
            window += value
            total += value
            count += 1
            if count % log_interval == 0:
                tk0.set_postfix(loss=window / log_interval)
                window = 0.0
        return total / count if count > 0 else 0

    def fit(self, train_dataloader, val_dataloader=None):
        for logger in self._iter_loggers():
            logger.log_hyperparams({'n_epoch': self.n_epoch, 'learning_rate': self.optimizer.param_groups[0]['lr'], 'loss_mode': self.mode})
        for epoch_i in range(self.n_epoch):
            print('epoch:', epoch_i)
            train_loss = self.train_one_epoch(train_dataloader)
            for logger in self._iter_loggers():
                logger.log_metrics({'train/loss': train_loss, 'learning_rate': self.optimizer.param_groups[0]['lr']}, step=epoch_i)
            if self.scheduler is not None:
                if epoch_i % self.scheduler.step_size == 0:
                    print("Current lr : {}".format(self.optimizer.state_dict()['param_groups'][0]['lr']))
                self.scheduler.step()
            if val_dataloader:
                auc = self.evaluate(self.model, val_dataloader)
                print('epoch:', epoch_i, 'validation: auc:', auc)
                for logger in self._iter_loggers():
                    logger.log_metrics({'val/auc': auc}, step=epoch_i)
                if self.early_stopper.stop_training(auc, self.model.state_dict()):
                    print(f'validation: best auc: {self.early_stopper.best_auc}')
                    self.model.load_state_dict(self.early_stopper.best_weights)
                    break
        torch.save(self.model.state_dict(), os.path.join(self.model_path, "model.pth"))
        for logger in self._iter_loggers():
            logger.finish()

    def _iter_loggers(self):
        if self.model_logger is None:
            return []
        if isinstance(self.model_logger, (list, tuple)):
            return list(self.model_logger)
        return [self.model_logger]

    def _predict_batches(self, model, data_loader, desc, labelled=True):
        model.eval()
        with torch.no_grad():
            for batch in tqdm.tqdm(data_loader, desc=desc, smoothing=0, mininterval=1.0):
                x_dict, y = batch if labelled else (batch, None)
                yield y, model(self._to_device(x_dict))
        if self.device.type == "cuda":
            from ..b200 import _lib
            _lib.check_errors(self.device)

    def evaluate(self, model, data_loader):
        targets, predicts = [], []
        for y, y_pred in self._predict_batches(model, data_loader, "validation"):
            targets.extend(y.tolist())
            predicts.extend(y_pred.tolist())
        return self.evaluate_fn(targets, predicts)

    def predict(self, model, data_loader):
        predicts = []
        for _, y_pred in self._predict_batches(model, data_loader, "predict"):
            predicts.extend(y_pred.tolist())
        return predicts

    def inference_embedding(self, model, mode, data_loader, model_path):
        """Embeddings of one tower for every batch of a label-free loader, from the saved ``model.pth`` (reference ``:243-258``)."""
        assert mode in ["user", "item"], "Invalid mode={}.".format(mode)
        model.mode = mode
        model.load_state_dict(torch.load(os.path.join(model_path, "model.pth"), map_location=self.device, weights_only=True))
        model = model.to(self.device)
        chunks = [emb.data for _, emb in self._predict_batches(model, data_loader, "%s inference" % mode, labelled=False)]
        return torch.cat(chunks, dim=0)

    def export_onnx(self, output_path, mode=None, dummy_input=None, batch_size=2, seq_length=10, opset_version=14, dynamic_batch=True, device=None, verbose=False, onnx_export_kwargs=None):
        """ONNX export is a deployment utility of the reference (``utils/onnx_export.py``) outside this engine's scope."""
        raise NotImplementedError("ONNX export is out of scope for the B200 hot-path engine: load model.pth into upstream "
                                  "torch-rechub (the state_dict layout is identical) and export there")
