"""`FlatAdamW` — the optimizer of the training loop: torch.optim.AdamW semantics (model_fqandtoyo.py:1599-1616) over flat
buffers, one libmobgt kernel per step (csrc/k9_optim.cu).

On construction every parameter's storage is moved into ONE flat fp32 buffer (`p.data` become views of it) and every
gradient into a second one (`p.grad` are views: the buffer the data-parallel all-reduce exchanges, and the one the Linear
backward kernels add into).  It is a `torch.optim.Optimizer`, so the reference's `PolynomialDecayLR` drives its learning rate
unchanged.  CUDA only: there is no CPU fallback (a model on the CPU gets `torch.optim.AdamW` from `configure_optimizers`)."""
import torch

from . import _C


class FlatAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        params = [p for p in params]
        if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise _C.MobgtError("FlatAdamW needs fp32 CUDA parameters (there is no CPU fallback)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        dev = params[0].device
        # every parameter starts on a 16-byte boundary of the flat buffers (the kernels read tables with 128-bit loads); the
        # padding elements in between stay zero (zero gradient, zero moments -> AdamW leaves them at zero)
        n = sum((p.numel() + 3) // 4 * 4 for p in params)
        self.n = n
        self.flat_param = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.steps = 0
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                view = self.flat_param[off:off + k].view_as(p)
                view.copy_(p.data)
                p.data = view
                old = p.grad
                p.grad = self.flat_grad[off:off + k].view_as(p)
                if old is not None:
                    p.grad.copy_(old)
                off += (k + 3) // 4 * 4

    @torch.no_grad()
    def step(self, closure=None):
        from . import ops
        g = self.param_groups[0]
        self.steps += 1
        _C.call("mobgt_adamw_step", _C.ptr(self.flat_param), _C.ptr(self.flat_grad), _C.ptr(self.exp_avg), _C.ptr(self.exp_avg_sq),
                self.n, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                self.steps, _C.stream_ptr())
        ops.weights_changed()           # the kernel wrote through the flat buffer: tensor version counters did not move

    def zero_grad(self, set_to_none=False):
        from . import ops
        self.flat_grad.zero_()          # gradients stay views of the flat buffer
        ops.grads_zeroed(self.flat_grad)

    def state_dict(self):
        return {"flat": True, "steps": self.steps, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
