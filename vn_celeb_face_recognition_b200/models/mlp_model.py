"""Drop-in for the reference's ``models.MLPModel`` (models/mlp_model.py:4-15): Linear(input_dim,2048)+ReLU ->
(dropout, training only) -> Linear(2048,C) -> log_softmax.  ``state_dict`` keys dense_1.*, dense_2.* as written by the
reference's trainer checkpoints (trainer/base_trainer.py:83-105; loaded at demo_image.py:16-21).  Forward = ONE launch of the
fused tail kernel (csrc/tail_fused.cu: both contractions on the tensor cores in split precision, bias / ReLU / log-softmax /
argmax in its row phases); no CPU fallback."""
import torch
from torch import nn

from .. import _lib, tail


class MLPModel(nn.Module):
    #: rows per launch of the fused tail kernel (scratch is allocated per 128-row bucket up to this size)
    chunk = 4096
    #: cached TailPlans (one per row-capacity bucket), least recently used first
    max_plans = 4

    def __init__(self, input_dim, num_classes):
        super().__init__()
        self.dense_1 = nn.Linear(input_dim, 2048)
        self.dense_2 = nn.Linear(2048, num_classes)
        self.input_dim, self.num_classes = input_dim, num_classes
        self._packed = None
        self._plans = {}

    def _invalidate(self):
        self._packed = None
        self._plans = {}

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    def split_layers(self, dev):
        """[(SplitLinear dense_1, "relu"), (SplitLinear dense_2, "logsoftmax")] on ``dev`` (packed once per weight set)."""
        # in-place updates (optimizer.step() of the frozen-encoder trainer) bump the parameters' version counters: the
        # packed copies follow them
        key = (dev,) + tuple(p._version for p in self.parameters())
        if self._packed is None or self._packed[0] != key:
            sd = self.state_dict()
            self._packed = (key, [(tail.SplitLinear(sd["dense_1.weight"], sd["dense_1.bias"], dev), "relu"),
                                  (tail.SplitLinear(sd["dense_2.weight"], sd["dense_2.bias"], dev), "logsoftmax")])
            self._plans = {}
        return self._packed[1]

    def _plan(self, n, dev):
        layers = self.split_layers(dev)
        bucket = -(-n // 128) * 128
        plan = self._plans.pop(bucket, None)
        if plan is None:
            plan = tail.TailPlan(layers, bucket, dev, in_mode=1)
            while len(self._plans) >= self.max_plans:
                self._plans.pop(next(iter(self._plans)))
        self._plans[bucket] = plan                      # most recently used last
        return plan

    def classify(self, emb, logp=None, threshold=0.0, thr_class=None):
        """Device fast path: emb fp32 (n, input_dim) -> (label int64 (n,), prob fp32 (n,)) [+ log-probs into ``logp``]:
        argmax / exp(max log-prob) / threshold of identify_person (demo_image.py:113-137) fused with the log-softmax."""
        n = emb.shape[0]
        dev = emb.device
        emb = emb.float()
        if emb.stride(1) != 1:
            emb = emb.contiguous()
        label = torch.empty(n, dtype=torch.int64, device=dev)
        prob = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            for s in range(0, n, self.chunk):
                m = min(self.chunk, n - s)
                self._plan(m, dev).run(m, x_f32=emb[s:s + m], out_vecs=[None, None if logp is None else logp[s:s + m]],
                                       label=label[s:s + m], prob=prob[s:s + m], thr=threshold, thr_class=thr_class,
                                       n_classes=self.num_classes)
        return label, prob

    def classify_half(self, emb16, logp=None):
        """Kept for callers that hold 16-bit embeddings: same as ``classify`` on their fp32 values."""
        return self.classify(emb16.float(), logp)

    def forward(self, input):
        if not (isinstance(input, torch.Tensor) and input.is_cuda):
            raise _lib.VnfrError("MLPModel.forward needs a CUDA tensor: this package has no CPU path")
        if self.training:
            # mlp_model.py:10-15 with dropout active: the differentiable path of the frozen-encoder trainer (trainer.py).  The
            # classifier is 0.4 % of a training step's FLOPs (the frozen encoder is this package's CUDA path); its forward /
            # backward here are torch's own CUDA kernels under autograd -- a library path, not a CPU fallback.
            x = torch.nn.functional.relu(self.dense_1(input))
            x = torch.nn.functional.dropout(x, p=0.5, training=True)
            return torch.nn.functional.log_softmax(self.dense_2(x), dim=1)
        with torch.no_grad():
            logp = torch.empty(input.shape[0], self.num_classes, dtype=torch.float32, device=input.device)
            self.classify(input.detach(), logp)
        return logp
