"""Fused AdamW for the fcwdm training path (replaces torch.optim.AdamW of train_util.py:75-82,391).

The master weights stay fp32 ``nn.Parameter``s (state-dict compatible); they are re-laid as views of ONE flat fp32
buffer, in the same order and with the same 16-byte-aligned offsets as the training engine's flat gradient, so an
optimizer step is a single ``fcwdm_adamw`` launch over ~54 M elements (HBM-bound: 4 reads + 3 writes of 4 bytes per
element) instead of several launches per parameter tensor.  ``grad_scale`` folds the 1/world_size of a summed
data-parallel all-reduce into the same pass.
"""
import torch

from . import ops


class FusedAdamW:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.param_groups = [{"lr": lr}]            # TrainLoop._anneal_lr writes param_group["lr"] (train_util.py:464-470)
        params = list(model.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW: move the model to the CUDA device first (no CPU path)")
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append((off, off + p.numel()))
            off += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, (lo, hi) in zip(params, self.offsets):
            self.flat[lo:hi].copy_(p.data.reshape(-1))
            p.data = self.flat[lo:hi].view(p.shape)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self._gather = torch.zeros_like(self.flat)

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    def _flat_grad(self):
        """The gradients as one flat tensor: the engine's buffer when .grad still aliases it, else gathered."""
        params = list(self.model.parameters())
        eng = getattr(self.model, "_train_engine", None)
        flat = getattr(eng, "last_flat", None) if eng is not None else None
        if flat is not None and flat.numel() == self.flat.numel() and all(
                p.grad is not None and p.grad.data_ptr() == flat.data_ptr() + 4 * lo
                for p, (lo, _) in zip(params, self.offsets)):
            return flat
        views = [self._gather[lo:hi].view(p.shape) for p, (lo, hi) in zip(params, self.offsets)]
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        torch._foreach_copy_(views, grads)
        return self._gather

    def step(self, grad_scale=1.0):
        self.step_count += 1
        lr = self.param_groups[0]["lr"]
        ops.adamw(self.flat, self._flat_grad(), self.m, self.v, lr, self.betas[0], self.betas[1], self.eps,
                  self.weight_decay, self.step_count, grad_scale)
        for eng_name in ("_engine", "_train_engine"):       # the raw-pointer update does not bump parameter versions
            eng = getattr(self.model, eng_name, None)
            if eng is not None:
                eng.invalidate()

    def state_dict(self):
        """The layout ``torch.optim.AdamW(model.parameters()).state_dict()`` has (what the reference's TrainLoop writes to
        ``opt_best_<contr>.pt`` / ``optNNNNNN.pt``, train_util.py:75-82,507-523): per-parameter ``step`` / ``exp_avg`` /
        ``exp_avg_sq`` under integer ids in ``model.parameters()`` order, one param group.  The tensors are views of the
        flat moment buffers (torch.save serialises them individually)."""
        params = list(self.model.parameters())
        state = {}
        if self.step_count > 0:
            for i, (p, (lo, hi)) in enumerate(zip(params, self.offsets)):
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.m[lo:hi].view(p.shape),
                            "exp_avg_sq": self.v[lo:hi].view(p.shape)}
        group = {"lr": self.param_groups[0]["lr"], "betas": tuple(self.betas), "eps": self.eps,
                 "weight_decay": self.weight_decay, "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": True,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts a ``torch.optim.AdamW`` state dict (written by the reference or by ``state_dict`` above) or the flat
        ``{step, m, v, lr}`` form earlier builds of this package wrote.  Sizes are validated before anything is copied."""
        params = list(self.model.parameters())
        if "state" in sd and "param_groups" in sd:
            ids = [i for g in sd["param_groups"] for i in g["params"]]
            if len(ids) != len(params):
                raise ValueError(f"optimizer state has {len(ids)} parameters, the model has {len(params)}")
            for pos, pid in enumerate(ids):
                st = sd["state"].get(pid)
                if st is not None and tuple(st["exp_avg"].shape) != tuple(params[pos].shape):
                    raise ValueError(f"optimizer state of parameter {pos}: shape {tuple(st['exp_avg'].shape)} does not "
                                     f"match the model's {tuple(params[pos].shape)}")
            steps = set()
            self.m.zero_()
            self.v.zero_()
            for pos, pid in enumerate(ids):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                lo, hi = self.offsets[pos]
                self.m[lo:hi].copy_(st["exp_avg"].reshape(-1))
                self.v[lo:hi].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): FusedAdamW keeps one step counter")
            self.step_count = steps.pop() if steps else 0
            self.param_groups[0]["lr"] = sd["param_groups"][0].get("lr", self.lr)
            return
        if not {"step", "m", "v"} <= set(sd):
            raise ValueError("unrecognised optimizer state: expected a torch.optim.AdamW state_dict ('state', "
                             "'param_groups') or the flat fcwdm form ('step', 'm', 'v')")
        if sd["m"].numel() != self.m.numel() or sd["v"].numel() != self.v.numel():
            raise ValueError(f"flat optimizer state has {sd['m'].numel()} elements, this model needs {self.m.numel()}")
        self.step_count = int(sd["step"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.param_groups[0]["lr"] = sd.get("lr", self.lr)
