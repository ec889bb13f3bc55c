"""Drop-in for the reference's ``models.MLPModel`` (models/mlp_model.py:4-15): Linear(input_dim,2048)+ReLU ->
(dropout, training only) -> Linear(2048,C) -> log_softmax.  ``state_dict`` keys dense_1.*, dense_2.* as written by the
reference's trainer checkpoints (trainer/base_trainer.py:83-105; loaded at demo_image.py:16-21).  Forward = two
tcgen05 GEMMs with fused bias/ReLU epilogues + one log-softmax kernel; no CPU fallback."""
import torch
from torch import nn

from .. import _lib, encoder_plan


class MLPModel(nn.Module):
    chunk = 4096
    #: 16-bit compute type (torch.float16 default / torch.bfloat16); None = encoder_plan.HALF
    half_dtype = None

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

    def _plan(self, n, dev):
        if self._packed is None:
            assert self.input_dim % 8 == 0, "input_dim must be a multiple of 8"
            self._packed = encoder_plan.MlpWeights(self.state_dict(), dev, self.half_dtype)
            self._plans = {}
        if n not in self._plans:
            self._plans[n] = encoder_plan.MlpPlan(self._packed, n, dev)
        return self._plans[n]

    def classify_half(self, emb16, logp=None):
        """Device fast path: emb 16-bit (n, input_dim) -> (label int64 (n,), prob fp32 (n,)) [+ log-probs into ``logp``]:
        argmax / exp(max log-prob) of identify_person (demo_image.py:126-130) fused with the log-softmax."""
        n = emb16.shape[0]
        dev = emb16.device
        label = torch.empty(n, dtype=torch.int64, device=dev)
        prob = torch.empty(n, dtype=torch.float32, device=dev)
        for s in range(0, n, self.chunk):
            m = min(self.chunk, n - s)
            plan = self._plan(m, dev)
            plan.x.view(m, self.input_dim).copy_(emb16[s:s + m])
            plan.run()
            _lib.call("vnfr_logsoftmax_argmax", _lib.ptr(plan.logits), m, self.num_classes, plan.logits.shape[1],
                      _lib.ptr(None if logp is None else logp[s:s + m]), _lib.ptr(label[s:s + m]), _lib.ptr(prob[s:s + m]),
                      _lib.stream_ptr())
        return label, prob

    def forward(self, input):
        if not (isinstance(input, torch.Tensor) and input.is_cuda):
            raise _lib.VnfrError("MLPModel.forward needs a CUDA tensor: this package has no CPU path")
        if self.training:
            raise _lib.VnfrError("training-mode forward (dropout p=0.5) is out of scope; call .eval()")
        with torch.no_grad():
            x16 = input.detach().to(self.half_dtype or encoder_plan.HALF).contiguous()
            logp = torch.empty(input.shape[0], self.num_classes, dtype=torch.float32, device=input.device)
            self.classify_half(x16, logp)
        return logp
