"""Frozen-encoder classifier training: what the reference's ``AugClassificationTrainer`` does per batch
(trainer/online_aug_trainer.py:6-45) --

    self.encoder.eval(); embedding = self.encoder(data).detach(); output = self.model(embedding)
    loss = criterion(output, target); optimizer.zero_grad(); loss.backward(); optimizer.step()

-- with the encoder forward on this package's CUDA path (InceptionResnetV1.forward: tcgen05 convolutions + fused tail, no
autograd graph, so ``.detach()`` is inherent) and the classifier's forward / backward / optimiser step left to torch (see
MLPModel.forward).  ``validate_epoch`` is the eval-mode pass of trainer/online_aug_trainer.py:55-98 through the fused
tail kernel.  Loss = ``neg_log_llhood`` (losses/__init__.py:3 = nn.NLLLoss), metric = ``accuracy`` (losses/metrics.py:3-7).
Logging, checkpoint rotation, early stopping and the LR schedulers of trainer/base_trainer.py are the caller's."""
import torch
from torch import nn


def accuracy(output, target):
    """losses/metrics.py:3-7"""
    return (torch.argmax(output, dim=1) == target).sum().item() / output.size(0)


class FrozenEncoderTrainer:
    def __init__(self, model, encoder, optimizer=None, criterion=None, device=None, lr=1e-4, weight_decay=1e-4):
        """model: MLPModel (trained); encoder: InceptionResnetV1 (frozen, eval).  Default optimiser = the reference's
        config: Adam(lr 1e-4, weight_decay 1e-4) (cfg/train_cfg_aug_emb_classify.json:104-110)."""
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.model, self.encoder = model.to(self.device), encoder.to(self.device)
        for p in self.encoder.parameters():
            p.requires_grad = False                     # online_aug_trainer.py:16-17
        self.encoder.eval()
        self.criterion = criterion or nn.NLLLoss()
        self.optimizer = optimizer or torch.optim.Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay)

    def embed(self, data):
        """encoder(data).detach() (online_aug_trainer.py:28): (B,3,S,S) standardised fp32 -> (B,512) unit-norm fp32."""
        self.encoder.eval()
        return self.encoder(data.to(self.device, non_blocking=True)).detach()

    def train_step(self, data, target):
        """One optimiser step on a batch; returns (loss, log-probabilities)."""
        self.model.train()
        target = target.to(self.device, non_blocking=True)
        output = self.model(self.embed(data))
        loss = self.criterion(output, target)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item(), output.detach()

    def train_epoch(self, loader):
        """loader yields (data, target[, id]) batches; returns the epoch's mean loss / accuracy (per-batch means averaged
        with the batch size as weight, as MetricTracker does)."""
        tot = {"loss": 0.0, "acc": 0.0, "n": 0}
        for batch in loader:
            data, target = batch[0], batch[1]
            loss, output = self.train_step(data, target)
            n = output.size(0)
            tot["loss"] += loss * n
            tot["acc"] += accuracy(output, target.to(self.device)) * n
            tot["n"] += n
        n = max(tot["n"], 1)
        return {"neg_log_llhood": tot["loss"] / n, "accuracy": tot["acc"] / n}

    def validate_epoch(self, loader):
        """Eval-mode pass (no dropout): encoder + fused tail kernel; returns val_neg_log_llhood / val_accuracy."""
        self.model.eval()
        tot = {"loss": 0.0, "acc": 0.0, "n": 0}
        with torch.no_grad():
            for batch in loader:
                data, target = batch[0], batch[1].to(self.device)
                output = self.model(self.embed(data))
                n = output.size(0)
                tot["loss"] += self.criterion(output, target).item() * n
                tot["acc"] += accuracy(output, target) * n
                tot["n"] += n
        n = max(tot["n"], 1)
        return {"val_neg_log_llhood": tot["loss"] / n, "val_accuracy": tot["acc"] / n}
