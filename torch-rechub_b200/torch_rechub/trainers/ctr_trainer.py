"""Single-task CTR trainer (mirror of reference ``torch_rechub/trainers/ctr_trainer.py:11-322``).

Same constructor, defaults and methods.  What changes underneath for CUDA devices:

* table gradients arrive as ONE persistent dense buffer per table filled by scatter-add kernels (no per-lookup
  ``vocab x dim`` zero-fill), so ``model.zero_grad(); loss.backward()`` costs HBM traffic proportional to the batch;
* with ``torch_rechub.b200.config.rowwise_optimizer`` (env ``RECHUB_B200_ROWWISE_OPT=1``) the tables are updated by the
  fused row-wise optimiser (``b200.optim``) that mirrors ``optimizer_fn`` on touched rows only; default keeps the
  reference's dense optimiser semantics exactly;
* with ``config.cuda_graph`` the step is captured once and replayed (``b200.graph``);
* under ``torchrun`` (WORLD_SIZE > 1) tables are sharded by field over the ranks (``b200.dist``) instead of the
  reference's single-process ``nn.DataParallel`` (ctr_trainer.py:53-55).
"""
import os

import torch
import tqdm
from sklearn.metrics import roc_auc_score

from ..basic.callback import EarlyStopper
from ..basic.loss_func import RegularizationLoss


class CTRTrainer(object):
    """A general trainer for single task learning.

    Args:
        model (nn.Module): the ranking model.
        optimizer_fn (torch.optim): optimizer class (default ``torch.optim.Adam``).
        optimizer_params (dict): optimizer kwargs (default ``{"lr": 1e-3, "weight_decay": 1e-5}``).
        regularization_params (dict): ``embedding_l1/embedding_l2/dense_l1/dense_l2`` coefficients.
        scheduler_fn (torch.optim.lr_scheduler): scheduler class, stepped once per epoch.
        scheduler_params (dict): scheduler kwargs.
        n_epoch (int): number of epochs.
        earlystop_patience (int): epochs without validation-AUC improvement before stopping (default 10).
        device (str): ``"cpu"`` or ``"cuda:0"``.
        gpus (list): device ids; more than one asks for multi-GPU training.
        loss_mode (bool): True: ``model(x) -> y_pred``; False: ``model(x) -> (y_pred, other_loss)``.
        model_path (str): directory where ``model.pth`` (best validation weights) is written.
        model_logger: a logger or list of loggers with ``log_hyperparams / log_metrics / finish``.
    """

    def __init__(
        self,
        model,
        optimizer_fn=torch.optim.Adam,
        optimizer_params=None,
        regularization_params=None,
        scheduler_fn=None,
        scheduler_params=None,
        n_epoch=10,
        earlystop_patience=10,
        device="cpu",
        gpus=None,
        loss_mode=True,
        model_path="./",
        model_logger=None,
    ):
        self.model = model
        if gpus is None:
            gpus = []
        self.gpus = gpus
        self.device = torch.device(device)
        if len(gpus) > 1:
            if self.device.type == "cuda":
                # The reference wraps the model in single-process nn.DataParallel, which re-broadcasts every table each
                # forward (1.66 GB/step at Criteo shape).  The B200 design is one process per GPU with field-sharded
                # tables: launch with torchrun and the trainer picks the ranks up from the environment.
                if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
                    raise RuntimeError("multi-GPU training runs one process per GPU: launch with "
                                       "`python -m torch.distributed.run --nproc-per-node %d your_script.py` "
                                       "(tables are sharded by field across ranks; see torch_rechub.b200.dist)" % len(gpus))
            else:
                print('parallel running on these gpus:', gpus)
                self.model = torch.nn.DataParallel(self.model, device_ids=gpus)
        self.model.to(self.device)
        self._dist = None
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and (self.device.type == "cuda" or os.environ.get("RECHUB_B200_SHARD_ON_CPU") == "1"):
            from ..b200 import dist as _dist
            self._dist = _dist.attach(self.model, self.device)
        if optimizer_params is None:
            optimizer_params = {"lr": 1e-3, "weight_decay": 1e-5}
        self.optimizer = self._make_optimizer(optimizer_fn, optimizer_params)
        if regularization_params is None:
            regularization_params = {"embedding_l1": 0.0, "embedding_l2": 0.0, "dense_l1": 0.0, "dense_l2": 0.0}
        self.scheduler = None
        if scheduler_fn is not None:
            self.scheduler = scheduler_fn(getattr(self.optimizer, "scheduler_target", self.optimizer), **scheduler_params)
        self.loss_mode = loss_mode
        self.criterion = torch.nn.BCELoss()
        if self.device.type == "cuda":
            from ..b200 import ops
            self.criterion = ops.EngineBCELoss()  # an nn.BCELoss: same arithmetic, one launch each way on CUDA
        self.evaluate_fn = roc_auc_score
        self.n_epoch = n_epoch
        self.early_stopper = EarlyStopper(patience=earlystop_patience)
        self.model_path = model_path
        self.reg_loss_fn = RegularizationLoss(**regularization_params)
        self.model_logger = model_logger
        self._graph_step = None
        self._lagged = None

    # ---------------------------------------------------------------------------------------------
    def _make_optimizer(self, optimizer_fn, optimizer_params):
        if self.device.type == "cuda":
            from ..b200 import config, optim
            if config.rowwise_optimizer:
                exclude = self._dist.replicated_ids if self._dist is not None else ()
                hybrid = optim.HybridOptimizer.build(self.model, optimizer_fn, optimizer_params, exclude=exclude)
                if hybrid is not None:
                    return hybrid
        return optimizer_fn([p for p in self.model.parameters() if p.numel() > 0], **optimizer_params)

    def _to_device(self, x_dict):
        if hasattr(x_dict, "copy_into"):  # b200.data.PackedColumns: <= 3 copies instead of one per column
            return x_dict.to(self.device)
        return {k: v.to(self.device) for k, v in x_dict.items()}

    def _loss(self, x_dict, y):
        if self.loss_mode:
            y_pred = self.model(x_dict)
            loss = self.criterion(y_pred, y)
        else:
            y_pred, other_loss = self.model(x_dict)
            loss = self.criterion(y_pred, y) + other_loss
        reg = self.reg_loss_fn(self.model)
        if isinstance(reg, float) and reg == 0.0:  # no regularisation configured: `loss + 0.0` would still be a launch per step
            return loss
        return loss + reg

    def _train_step(self, x_dict, y):
        """zero_grad -> forward -> loss -> backward -> optimizer step; returns the loss tensor (reference ``:87-99``)."""
        if self._dist is not None:
            return self._dist.train_step(self, x_dict, y)
        rw = getattr(self.optimizer, "rowwise", None)
        if rw is not None:
            rw.advance_early()  # step counter / bias corrections: off the critical path (joined by optimizer.step())
        loss = self._loss(x_dict, y)
        self.model.zero_grad()
        if loss.is_cuda and loss.dim() == 0 and loss.dtype == torch.float32:
            # a cached unit gradient: `loss.backward()` would fill a fresh ones-tensor every step (one more launch on the chain)
            one = getattr(self, "_unit_grad", None)
            if one is None or one.device != loss.device:
                one = self._unit_grad = torch.ones((), dtype=torch.float32, device=loss.device)
            loss.backward(one)
        else:
            loss.backward()
        self.optimizer.step()
        return loss.detach()  # callers only read the value; holding the graph alive would pin AccumulateGrad nodes

    def train_one_epoch(self, data_loader, log_interval=10):
        self.model.train()
        stats = {"total": 0.0, "epoch": 0.0, "count": 0}
        on_cuda = self.device.type == "cuda"
        lagged = None
        if on_cuda:
            from ..b200 import _lib, config, graph
            if config.lagged_loss:
                if self._lagged is None:
                    self._lagged = graph.LaggedReader(self.device)  # pinned slots + events, kept across epochs
                lagged = self._lagged
                lagged.reset()
        tk0 = tqdm.tqdm(data_loader, desc="train", smoothing=0, mininterval=1.0)

        def account(loss_value):
            stats["total"] += loss_value
            stats["epoch"] += loss_value
            stats["count"] += 1
            if stats["count"] % log_interval == 0:
                tk0.set_postfix(loss=stats["total"] / log_interval)
                stats["total"] = 0.0

        for x_dict, y in tk0:
            if on_cuda and config.cuda_graph and (self._dist is None or config.dist_cuda_graph):
                if self._graph_step is None:
                    self._graph_step = graph.GraphedStep(self)
                loss = self._graph_step(x_dict, y)  # copies the (host) batch into the graph's static inputs
            else:
                x_dict = self._to_device(x_dict)
                y = y.to(self.device).float()
                loss = self._train_step(x_dict, y)
            if lagged is not None:
                # the reference syncs on loss.item() every step (ctr_trainer.py:100); here step t's loss is read while step
                # t+1 is already queued, so the device never waits for the host
                lagged.push(loss)
                if lagged.pending() > 1:
                    account(lagged.pop())
                continue
            loss_value = loss.item()
            if on_cuda:
                _lib.check_errors(self.device)  # out-of-range ids -> IndexError, at the reference's own sync point
            account(loss_value)
        while lagged is not None and lagged.pending():
            account(lagged.pop())
        return stats["epoch"] / stats["count"] if stats["count"] > 0 else 0

    def fit(self, train_dataloader, val_dataloader=None):
        for logger in self._iter_loggers():
            logger.log_hyperparams({'n_epoch': self.n_epoch, 'learning_rate': self.optimizer.param_groups[0]['lr'], 'loss_mode': self.loss_mode})

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

        if self._dist is None:
            torch.save(self.model.state_dict(), os.path.join(self.model_path, "model.pth"))
        else:  # every rank takes part in gathering the sharded tables; rank 0 writes the reference-layout checkpoint
            full = self._dist.full_state_dict()
            if self._dist.rank == 0:
                torch.save(full, os.path.join(self.model_path, "model.pth"))

        for logger in self._iter_loggers():
            logger.finish()

    def _iter_loggers(self):
        if self.model_logger is None:
            return []
        if isinstance(self.model_logger, (list, tuple)):
            return list(self.model_logger)
        return [self.model_logger]

    def _predict_batches(self, model, data_loader, desc):
        model.eval()
        with torch.no_grad():
            for x_dict, y in tqdm.tqdm(data_loader, desc=desc, smoothing=0, mininterval=1.0):
                x_dict = self._to_device(x_dict)
                y_pred = model(x_dict) if self.loss_mode else model(x_dict)[0]
                yield y, y_pred
        if self.device.type == "cuda":
            from ..b200 import _lib
            _lib.check_errors(self.device)

    def evaluate(self, model, data_loader):
        targets, predicts = list(), list()
        for y, y_pred in self._predict_batches(model, data_loader, "validation"):
            targets.extend(y.float().view(-1, 1).tolist())
            predicts.extend(y_pred.tolist())
        return self.evaluate_fn(targets, predicts)

    def predict(self, model, data_loader):
        predicts = list()
        for _, y_pred in self._predict_batches(model, data_loader, "predict"):
            predicts.extend(y_pred.tolist())
        return predicts

    def export_onnx(self, output_path, dummy_input=None, batch_size=2, seq_length=10, opset_version=14, dynamic_batch=True, device=None, verbose=False, onnx_export_kwargs=None):
        """ONNX export is a deployment utility of the reference (``utils/onnx_export.py``) outside this engine's scope."""
        raise NotImplementedError("ONNX export is out of scope for the B200 hot-path engine: load model.pth into upstream "
                                  "torch-rechub (the state_dict layout is identical) and export there")

    def visualization(self, *args, **kwargs):
        """torchview visualisation is outside this engine's scope (reference ``ctr_trainer.py:246``)."""
        raise NotImplementedError("model visualisation is out of scope for the B200 hot-path engine; use upstream torch-rechub")
