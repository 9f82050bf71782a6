"""TEST INFRASTRUCTURE — the data-parallel restatement of the oracle: `ShardedOracle` runs `OracleSequential`'s step on ONE
RANK'S ROWS of the global batch and sums over the ranks exactly the quantities the CUDA data-parallel path sums
(neuralnetlib_b200/parallel.py): BatchNorm moments and backward reductions, the squared norms behind every error clip, the
loss numerator, and the weight / bias gradients. `allreduce(np.ndarray) -> np.ndarray` is injected (gloo in the CPU tests).
If the schedule is right, every rank ends the step with the weights the single-process oracle gets on the global batch.
"""
import numpy as np

import nnl_oracle as O


class ShardedOracle(O.OracleSequential):
    def __init__(self, specs, loss, optimizer, seed, allreduce, world):
        super().__init__(specs, loss, optimizer, seed)
        self.allreduce, self.world = allreduce, world

    # -- BatchNorm with rank-summed moments (layers.py:1230-1256 on the global batch)
    def _bn_fwd(self, x, L, training):
        eps, mom = L.get("eps", 1e-5), L.get("momentum", 0.9)
        xc = np.clip(x, -10, 10)
        if not training:
            return O.bn_fwd(x, L["gamma"], L["beta"], L["running_mean"], L["running_var"], mom, eps, False)[0]
        n = x.shape[0] * self.world
        mom12 = self.allreduce(np.stack([xc.sum(axis=0), (xc * xc).sum(axis=0)]))
        mean = mom12[0] / n
        batch_var = mom12[1] / n - mean * mean + eps
        L["running_mean"] = mom * L["running_mean"] + (1 - mom) * mean
        L["running_var"] = mom * L["running_var"] + (1 - mom) * batch_var
        std = np.sqrt(batch_var + eps)
        L["dp_cache"] = (xc - mean, (xc - mean) / std, std, n)
        return L["gamma"] * L["dp_cache"][1] + L["beta"]

    def _bn_bwd(self, err, L):
        xc, xh, std, n = L["dp_cache"]
        sums = self.allreduce(np.stack([err.sum(axis=0), (err * xh).sum(axis=0)]))
        L["dbeta"], L["dgamma"] = sums[0], sums[1]
        return L["gamma"] / std * (err - xh * sums[1] / n - sums[0] / n)   # layers.py:1264-1274 in closed form

    def forward(self, x, training=True):
        bn_fwd = O.bn_fwd
        this = self

        def patched(x_, gamma, beta, rm, rv, mom, eps, training_):
            L = next(l for l in this.layers if l.get("gamma") is gamma)
            y = this._bn_fwd(x_, L, training_)
            return y, L["running_mean"], L["running_var"], None
        O.bn_fwd = patched
        try:
            return super().forward(x, training)
        finally:
            O.bn_fwd = bn_fwd

    def _clip_global(self, e):
        n2 = float(self.allreduce(np.array([np.sum(e * e)]))[0])
        return e / np.sqrt(n2) if n2 > 1.0 else e

    def backward(self, error, gan=False, compute_only=False, trace=None):
        n = len(self.layers)
        for i, L in enumerate(reversed(self.layers)):
            if i == 0 and L["type"] == "act":
                if not gan:
                    if (L["fn"] == "softmax" and self.loss == "cce") or (L["fn"] == "sigmoid" and self.loss == "bce"):
                        error = self.predictions - self.y_true
                else:
                    error = self._layer_bwd(L, error)
            else:
                error = self._clip_global(error)                      # models.py:214 on the GLOBAL error tensor
                error = self._bn_bwd(error, L) if L["type"] == "bn" else self._layer_bwd(L, error)
            if compute_only:
                continue
            if "w" in L:
                L["dw"], L["db"] = self.allreduce(L["dw"]), self.allreduce(L["db"])   # SUM over ranks (Q4)
                self.optimizer.update(n - 1 - i, L["w"], O.clip_l2(L["dw"]), L["b"], O.clip_l2(L["db"]))
        return error

    def train_on_batch(self, x, y, trace=None):
        y = np.asarray(y, dtype=np.float64)
        self.y_true = y
        p = self.forward(x, training=True)
        rows = x.shape[0]
        local = O.loss_value(self.loss, y, p.copy()) * (rows if self.loss == "cce" else rows * p.shape[1])
        total = float(self.allreduce(np.array([local]))[0])
        denom = rows * self.world if self.loss == "cce" else rows * self.world * p.shape[1]
        err = O.loss_derivative(self.loss, y, p.copy())
        if self.loss == "mse":
            err = err / self.world   # losses.py:84 divides by y_true.size of the WHOLE batch, not of this rank's shard
        self.backward(err)
        return total / denom
