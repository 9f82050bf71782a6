"""Resident execution plan of a `Sequential` for one (batch shape, training flag).

Replaces the per-layer NumPy loop of models.py:159-264 by a static program over pre-allocated HBM buffers:
  * forward: [Dense|Conv2D|Conv2DTranspose]+Activation pairs run as one kernel (bias + activation epilogue);
    Input/Flatten/Reshape are views;
  * backward: the reference's clip of every error tensor to L2 norm <= 1 (models.py:214) is carried lazily as a
    chain of float64 sum-of-squares slots (include/b200nn.h), so no extra pass over any tensor is needed; the
    backward of an Activation is folded into the dx epilogue of the layer that consumes its output; the dx of the
    first parametrised layer is skipped inside fit (nothing observes it);
  * update: per-tensor gradient clip (models.py:240-242) + Adam/SGD for every parameter tensor in one fused
    multi-tensor launch, Adam's per-layer `t` advanced on the device;
  * the whole step (step_begin, forward, loss, backward, update) is captured into one CUDA graph on its second
    execution and replayed afterwards: one host call per step, no allocation, no host<->device traffic except the
    batch upload and (when asked for) the loss read-back.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _backend as B
from . import _ops as ops
from .layers import Activation, BatchNormalization, Conv2D, Dense, Input, _Err
from .losses import BinaryCrossentropy, CategoricalCrossentropy

MAX_SLOTS = 256


class _Program:
    """A list of kernel launches that is run eagerly once and replayed as a CUDA graph afterwards."""

    def __init__(self, name):
        self.name = name
        self.ops = []
        self.meta = []      # (kernel family name, algorithmic bytes) per op, parallel to ops
        self.graph = None
        self.runs = 0

    def run(self, use_graph=True):
        if self.graph is not None:
            ops.graph_launch(self.graph)
        elif use_graph and self.runs >= 1:
            ops.graph_begin()
            try:
                for fn in self.ops:
                    fn()
            finally:
                self.graph = ops.graph_end()
            ops.graph_launch(self.graph)
        else:
            for fn in self.ops:
                fn()
        self.runs += 1

    def destroy(self):
        if self.graph is not None:
            ops.graph_destroy(self.graph)
            self.graph = None


class Plan:
    def __init__(self, model, batch_shape, training: bool):
        self.model = model
        self.layers = list(model.layers)
        self.batch_shape = tuple(int(s) for s in batch_shape)
        self.B = self.batch_shape[0]
        self.training = bool(training)
        # accumulators (loss numerator, correct count) directly in front of the clip slots: under data parallelism the
        # loss kernel's outputs {acc, slot 0} then travel in ONE exchange
        self._ssacc = B.zeros((4 + MAX_SLOTS,), dtype=torch.float64)
        self.acc = self._ssacc[0:4]
        self.ss = self._ssacc[4:]
        self.n_slots = 0
        self.dp = getattr(model, "_dp", None)
        self.world = self.dp.world if self.dp is not None else 1
        # Deferred clips (default under data parallelism; B200NN_DEFER_CLIPS=1 / 0 forces it on / off): every backward op is linear
        # in the error, so the backward kernels run WITHOUT the clip multipliers, record the raw sums of squares, and the
        # whole pending chain is applied once to each weight gradient inside the optimizer launch. The clip norms then need
        # ONE exchange per step instead of one after every backward kernel.
        mode = os.environ.get("B200NN_DEFER_CLIPS", "auto")
        self.defer_clips = mode == "1" or (mode != "0" and self.dp is not None)
        self._raw_begin = 0
        self._acc_pending = False
        self._upd_chain = {}
        self._ws = None
        self._ws_bytes = 0
        self._emit_to = None
        self.x_in = B.empty(self.batch_shape)
        self.x_stage = None
        self.y_in = None
        self.y_stage = None
        self.state = [dict() for _ in self.layers]
        self.outs = [None] * len(self.layers)
        self.fwd_fused = [False] * len(self.layers)   # Activation i applied inside layer i-1's kernel
        self.programs = {}
        # B200NN_FUSE_BLOCKS: 0 = layer by layer, 1 = BatchNorm+MaxPool kernels, 2 (default) = also the first-block kernels
        level = int(os.environ.get("B200NN_FUSE_BLOCKS", "2"))
        self.fuse_blocks, self.fuse_first_block = level >= 1, level >= 2
        self._head = None
        self._build_forward()

    # ---- allocation context used by the layers' plan hooks ---------------------------------------------------
    def buf(self, shape):
        return B.empty(tuple(int(s) for s in shape))

    def slot(self):
        if self.n_slots >= MAX_SLOTS:
            raise RuntimeError("too many clip slots in one plan")
        s = self.n_slots
        self.n_slots += 1
        return s

    def need_ws(self, nbytes):
        self._ws_bytes = max(self._ws_bytes, int(nbytes))

    def buf64(self, shape):
        return B.empty(tuple(int(s) for s in shape), dtype=torch.float64)

    # ---- data parallelism (parallel.py, csrc/comm.cu) ------------------------------------------------------------------
    def allreduce(self, t, name="allreduce"):
        """Emit an in-place sum of `t` over the ranks (no-op on one GPU). Small payloads get a one-shot exchange site."""
        if self.dp is None:
            return
        comm = self.dp.comm
        site = comm.site(t.numel() * t.element_size())
        self.emit(lambda: comm.allreduce(t, site), name, ())

    def peer_region(self, nbytes):
        """(comm handle, site, offset) of a fresh exchange region for a peer-fused kernel, or None (no peer memory / disabled)."""
        if self.dp is None or os.environ.get("B200NN_PEER_FUSED", "1") == "0":
            return None
        r = self.dp.comm.region(nbytes)
        return None if r is None else (self.dp.comm.handle, r[0], r[1])

    def pending_chain(self):
        """With deferred clips: the range chain over every slot recorded since the backward began; else None."""
        if not self.defer_clips:
            return None
        return ("range", self._raw_begin, self.n_slots - self._raw_begin)

    def _sync_slots(self, first, final=False, mergeable=False):
        """Sum the clip slots [first, n_slots) produced by the op(s) just emitted over the ranks: the reference's clips
        act on whole-batch tensors (models.py:214), so their norms are global. Slots already exchanged (sync_slots_now
        called from inside a layer's hook) are skipped. With deferred clips nothing is exchanged between the backward
        kernels: one exchange (loss accumulators included) happens when `final`."""
        if self.defer_clips and not final:
            return
        if self.defer_clips:
            first = getattr(self, "_slots_synced", 0)   # nothing was exchanged since the backward began
        first = max(first, getattr(self, "_slots_synced", 0))
        if self.dp is not None:
            if self._acc_pending and first == 0 and final and mergeable and getattr(self, "_merge_final", False):
                # train step with deferred clips: this exchange can share ONE NVLink round trip with the last small gradient
                # bucket (_grad_exchange emits it)
                self._final_slots = self._ssacc[0:4 + self.n_slots]
            elif self._acc_pending and first == 0:
                self.allreduce(self._ssacc[0:4 + self.n_slots], "slot_exchange")   # accumulators sit right in front of slot 0
            else:
                if self._acc_pending:
                    self.allreduce(self.acc, "loss_exchange")
                if self.n_slots > first:
                    self.allreduce(self.ss[first:self.n_slots], "slot_exchange")
        self._acc_pending = False
        self._slots_synced = self.n_slots

    def sync_slots_now(self):
        self._sync_slots(0, final=True)

    @property
    def ws(self):
        if self._ws is None or self._ws.numel() * 4 < self._ws_bytes:
            self._ws_keep = getattr(self, "_ws_keep", []) + [self._ws]   # captured graphs may still point at it
            self._ws = B.empty((max(self._ws_bytes // 4, 1),))
        return self._ws

    def emit(self, fn, name=None, io=()):
        """Record one launch; `io` lists the tensors it must read or write once each (algorithmic bytes)."""
        self._emit_to.ops.append(fn)
        seen, nbytes = set(), 0
        for t in io:
            if t is not None and t.data_ptr() not in seen:
                seen.add(t.data_ptr())
                nbytes += t.numel() * t.element_size()
        self._emit_to.meta.append((name or "op", nbytes))

    # ---- forward ------------------------------------------------------------------------------------------------
    def _build_forward(self):
        prog = _Program("forward")
        self._emit_to = prog
        layers = self.layers
        x = self.x_in
        i, n = 0, len(layers)
        while i < n:
            L = layers[i]
            if L._is_view:
                shp = L._out_shape(tuple(x.shape))
                if isinstance(L, Input):
                    if tuple(L.input_shape) != tuple(x.shape[1:]) and int(np.prod(L.input_shape)) != int(np.prod(x.shape[1:])):
                        raise ValueError(f"Input layer expects {tuple(L.input_shape)}, got {tuple(x.shape[1:])}")
                else:
                    L.input_shape = tuple(x.shape)
                y = x.reshape(shp)
                self.state[i].update(in_shape=tuple(x.shape))
                self.outs[i] = y
                x = y
                i += 1
                continue
            nxt = layers[i + 1] if i + 1 < n else None
            fuse = L._fuses_act and isinstance(nxt, Activation) and nxt._fusable
            self.state[i].update(in_shape=tuple(x.shape))
            if self.fuse_first_block and isinstance(L, Conv2D) and self._try_conv_block(i, x, fuse):
                i, x = self._block_next
                continue
            if isinstance(L, BatchNormalization) and self.fuse_blocks and L._can_fuse_pool(x, nxt, self.training):
                y = L._plan_forward(self, self.state[i], x, None, self.training, pool=nxt)   # BN + MaxPool: one kernel
                self.outs[i] = None                      # the BatchNorm output is never materialised
                self.outs[i + 1] = y
                self.state[i + 1].update(in_shape=tuple(x.shape), fused_into_bn=True)
                i += 2
                x = y
                continue
            # classifier head Dense(N <= 32) + Softmax as the LAST two layers: remember its two launches so that a train
            # step can replace them (and the loss kernel) by the fused head kernel
            is_head = (i == n - 2 and isinstance(L, Dense) and isinstance(nxt, Activation) and not fuse and x.dim() == 2
                       and type(nxt.activation_function).__name__ == "Softmax" and L.units <= 32)
            op0 = len(prog.ops)
            y = L._plan_forward(self, self.state[i], x, nxt if fuse else None, self.training)
            if is_head:
                self._head = dict(layer=L, x=x, op0=op0)
            self.outs[i] = y
            if fuse:
                self.outs[i + 1] = y
                self.state[i + 1].update(y=y, in_shape=tuple(y.shape))
                self.fwd_fused[i + 1] = True
                i += 2
            else:
                i += 1
            x = y
        self.pred = x
        if getattr(self, "_head", None) is not None and len(prog.ops) != self._head["op0"] + 2:
            self._head = None
        self.programs["forward"] = prog
        self._emit_to = None

    def _try_conv_block(self, i, x, has_act):
        """Conv2D [+ Activation] -> BatchNormalization -> MaxPooling2D(2) as the fused first-block kernels (conv output never
        materialised). Returns True and sets self._block_next = (next layer index, pooled output) when the pattern applies."""
        layers, n = self.layers, len(self.layers)
        j = i + (2 if has_act else 1)
        if j + 1 >= n:
            return False
        L, act, bn, pool = layers[i], (layers[i + 1] if has_act else None), layers[j], layers[j + 1]
        if not L._can_fuse_block(x, act, bn, pool, self.training):
            return False
        y = L._plan_forward_block(self, self.state[i], self.state[j], x, act, bn)
        shp = tuple(x.shape[:3]) + (L.filters,)
        for k in range(i, j + 1):
            self.outs[k] = None
            self.state[k].update(in_shape=shp if k > i else tuple(x.shape))
        if has_act:
            self.fwd_fused[i + 1] = True
        self.outs[j + 1] = y
        self.state[j + 1].update(in_shape=shp, conv_block=(i, j, has_act))
        self._block_next = (j + 2, y)
        return True

    # ---- backward -----------------------------------------------------------------------------------------------
    def _has_weights(self, L):
        return hasattr(L, "weights")

    def _build_backward(self, err: _Err, gan: bool, compute_only: bool, want_input_error: bool):
        """Emits the backward kernels into self._emit_to; returns (updates, input_error or None)."""
        layers = self.layers
        n = len(layers)
        updates = []
        cur = err
        i = n - 1
        self._grad_begin()
        raw = self.defer_clips
        self._upd_chain = {}
        if raw:
            self._raw_begin = min(cur.chain) if cur.chain else self.n_slots
            cur = _Err(cur.t, ())
        if isinstance(layers[i], Activation):
            if gan:   # models.py:211-212: the real derivative of the last activation, no clip before it
                first = self.n_slots
                cur = layers[i]._plan_backward(self, self.state[i], cur, None, True, y=self.outs[i])
                cur = _Err(cur.t, ()) if raw else cur
                self._sync_slots(first)
            # not gan: p - y (recognised pairs) or the raw loss derivative with the activation's derivative
            # skipped (models.py:198-210, Q7) -- both already sit in `err`
            i -= 1
        first = self.n_slots
        while i >= 0 and cur is not None:
            self._sync_slots(first)   # slots of the previous op (no-op on one GPU)
            first = self.n_slots
            L = layers[i]
            if L._is_view:
                cur = _Err(cur.t.reshape(self.state[i]["in_shape"]), cur.chain)
                i -= 1
                continue
            need_dx = want_input_error or any(self._has_weights(layers[k]) for k in range(i))
            if not need_dx and not self._has_weights(L):
                cur = None   # nothing upstream is trainable and nobody observes the input error
                break
            n_before = self.n_slots   # the error entering this op still owes slots [_raw_begin, n_before) when clips are deferred
            pend = ("range", self._raw_begin, n_before - self._raw_begin) if raw else None
            if isinstance(L, Activation):
                cur = L._plan_backward(self, self.state[i], cur, None, True, y=self.outs[i])
                cur = _Err(cur.t, ()) if raw else cur
                i -= 1
                continue
            if self.state[i].get("conv_block") is not None:
                ci, bi, has_act = self.state[i]["conv_block"]
                need_dx_c = want_input_error or any(self._has_weights(layers[k]) for k in range(ci))
                et = cur.t.reshape(self.outs[i].shape)
                cur = layers[ci]._plan_backward_block(self, self.state[ci], self.state[bi], layers[bi], _Err(et, cur.chain), has_act,
                                                      need_dx_c)
                if raw and cur is not None:
                    cur = _Err(cur.t, ())
                # the block's kernels apply their pending chain to dw / db themselves (the reduce launch), also when deferred --
                # unless the block left it to the optimizer (last op of a data-parallel train step, see _plan_backward_block)
                dchain = self.state[ci].pop("wgrad_chain_deferred", None)
                if dchain is not None:
                    layers[ci]._grad_pending = (self.ss, dchain[1], dchain[2])
                    self._upd_chain[ci] = dchain
                else:
                    layers[ci]._grad_pending = None
                if not compute_only:
                    updates.append(ci)
                    self._grad_bucket_ready(layers[ci])
                    self._maybe_early_update(updates)
                i = ci - 1
                continue
            if self.state[i].get("fused_into_bn"):
                # MaxPool whose forward ran inside the preceding BatchNorm kernel: pool-bwd + BN-bwd (+ the backward of
                # the Activation feeding the BatchNorm) as one kernel
                bn_i = i - 1
                mask, j = None, bn_i - 1
                if j >= 1 and isinstance(layers[j], Activation) and layers[j]._fusable:
                    mask = layers[j]
                et = cur.t.reshape(self.outs[i].shape)
                cur = layers[bn_i]._plan_backward_fused_pool(self, self.state[bn_i], _Err(et, cur.chain), mask)
                cur = _Err(cur.t, ()) if raw else cur
                i = j - 1 if mask is not None else bn_i - 1
                continue
            mask, j = None, i - 1
            if L._supports_mask and need_dx:
                while j >= 0 and layers[j]._is_view and not isinstance(layers[j], Input):
                    j -= 1
                if j >= 1 and isinstance(layers[j], Activation) and layers[j]._fusable:
                    mask = layers[j]
            et = cur.t.reshape(self.outs[i].shape)
            cur = L._plan_backward(self, self.state[i], _Err(et, cur.chain), mask, need_dx)
            if raw and cur is not None:
                cur = _Err(cur.t, ())
            if self._has_weights(L):
                L._grad_pending = (self.ss, pend[1], pend[2]) if raw else None
                self._upd_chain[i] = pend if raw else ()
            if self._has_weights(L) and not compute_only:
                updates.append(i)
                self._grad_bucket_ready(L)
                self._maybe_early_update(updates)
            if mask is not None and cur is not None:
                cur = _Err(cur.t.reshape(self.outs[j].shape), cur.chain)
                i = j - 1
            else:
                i -= 1
        self._sync_slots(first, final=True, mergeable=True)
        if raw and cur is not None:
            cur = _Err(cur.t, self.pending_chain())   # whoever materialises the input error applies the whole chain
        return updates, cur

    # Gradient buckets (parallel.py): a bucket is summed over the ranks as soon as the backward of the layer that completes
    # it has been emitted -- large buckets with NCCL on a SIDE stream (forked / joined with events, so the captured graph
    # overlaps the exchange with the rest of backward), small ones with the one-shot exchange on the main stream.
    def _grad_begin(self):
        self._buckets_done = 0
        self._side_used = False

    def _grad_bucket_ready(self, layer, final=False):
        dp = self.dp
        if dp is None or dp.buckets is None:
            return
        while self._buckets_done < len(dp.buckets) and (final or dp.bucket_layers[self._buckets_done] is layer):
            lo, hi, _ = dp.buckets[self._buckets_done]
            if (not final and getattr(self, "_merge_final", False) and self._buckets_done == len(dp.buckets) - 1
                    and (hi - lo) * 4 <= 3072 and dp.comm.peer):
                return    # the last small bucket of a train step travels together with the clip slots (_grad_exchange)
            self._buckets_done += 1
            t = dp.flat[lo:hi]
            site = dp.comm.site(t.numel() * 4) if t.numel() * 4 <= (64 << 10) else -1
            if site >= 0 or final:
                self.emit(lambda t=t, site=site: dp.comm.allreduce(t, site), "grad_allreduce", (t,))
                continue
            evt, side, main = torch.cuda.Event(), dp.side, B.stream_obj()

            def fork(t=t, evt=evt):
                evt.record(main)
                side.wait_event(evt)
                dp.comm.allreduce(t, -1, side.cuda_stream)
            self.emit(fork, "grad_allreduce_async", (t,))
            self._side_used = True

    # Early optimizer launch (single GPU, train step): once every parametrised layer but the FIRST has its gradients, their
    # Adam / SGD update runs on a side stream while the first layer's backward (the largest kernel of a CNN) still computes;
    # only the first layer's update remains on the critical path. The per-layer `t` and clip make the split exact.
    def _early_setup(self, enabled):
        params = [i for i, L in enumerate(self.layers) if self._has_weights(L)]
        self._early = None
        if enabled and len(params) >= 2 and self.dp is None and not self.defer_clips and os.environ.get("B200NN_EARLY_OPT", "1") != "0":
            self._early = dict(fire_after=params[1], n_layers=len(params), done=False, side=torch.cuda.Stream(device=B.device()))

    def _maybe_early_update(self, updates):
        e = getattr(self, "_early", None)
        if e is None or e["done"] or updates[-1] != e["fire_after"]:
            return
        e["done"] = True
        opt = self.model.optimizer
        ups = [(li, self.layers[li]._weights, self.layers[li]._d_weights, self.layers[li]._bias, self.layers[li]._d_bias, ())
               for li in updates]
        oplan = opt._make_plan(ups, self.ss, 1, e["n_layers"])
        e["count"] = len(ups)
        evt, side, main = torch.cuda.Event(), e["side"], B.stream_obj()

        def fork():
            evt.record(main)
            side.wait_event(evt)
            opt._run_plan(oplan, side.cuda_stream)
        self.emit(fork, "optimizer_multi_early", opt._plan_io(oplan))

    def _early_join(self):
        e = getattr(self, "_early", None)
        if e is None or not e["done"]:
            return 0
        evt, side, main = torch.cuda.Event(), e["side"], B.stream_obj()

        def join():
            evt.record(side)
            main.wait_event(evt)
        self.emit(join, "optimizer_early_join", ())
        return e["count"]

    def _grad_exchange(self, updates):
        """Before the optimizer: flush the buckets not yet issued and join the side stream."""
        if self.dp is None:
            return
        pend, self._final_slots = getattr(self, "_final_slots", None), None
        if pend is not None:
            dp, merged = self.dp, False
            if updates and dp.buckets is not None and self._buckets_done == len(dp.buckets) - 1 and dp.comm.peer:
                lo, hi, _ = dp.buckets[-1]
                t = dp.flat[lo:hi]
                words = ((t.numel() + 1) & ~1) + 2 * pend.numel()
                if words <= 1024:
                    site = dp.comm.site(4 * words)
                    if site >= 0:
                        self.emit(lambda: dp.comm.allreduce2(t, pend, site), "grad_slot_exchange", (t, pend))
                        self._buckets_done += 1
                        merged = True
            if not merged:
                self.allreduce(pend, "slot_exchange")
        if not updates:
            return
        self._grad_bucket_ready(None, final=True)
        if self._side_used:
            evt, side, main = torch.cuda.Event(), self.dp.side, B.stream_obj()

            def join():
                evt.record(side)
                main.wait_event(evt)
            self.emit(join, "grad_allreduce_join", ())

    def _try_mlp_step(self, prog):
        """The whole train step of a small MLP as ONE cooperative kernel (csrc/mlp_step.cu; include/b200nn.h b200nn_mlp_train_step):
        [Input] (Dense [+ fusable Activation])+ Dense + Softmax, categorical cross-entropy, plain Adam, batch <= 64, one GPU.
        Same buffers, slots, stored gradients and optimizer state as the layer-by-layer program. B200NN_MLP_STEP=0 disables it."""
        from .optimizers import Adam
        if os.environ.get("B200NN_MLP_STEP", "1") == "0" or self.dp is not None or self.defer_clips or not self.training:
            return False
        opt, lf = self.model.optimizer, self.model.loss_function
        if type(opt) is not Adam or opt.clip_norm is not None or opt.clip_value is not None:
            return False
        if not isinstance(lf, CategoricalCrossentropy) or getattr(lf, "_sparse", False) or lf._kind != ops.LOSS_CCE:
            return False
        if self.B > B.MLP_MAX_BATCH or len(self.batch_shape) != 2:
            return False
        layers, n = self.layers, len(self.layers)
        dense = []          # (layer index, Dense, activation layer or None)
        i = 0
        while i < n:
            L = layers[i]
            if isinstance(L, Input):
                if i != 0:
                    return False
                i += 1
                continue
            if type(L) is not Dense:
                return False
            nxt = layers[i + 1] if i + 1 < n else None
            if isinstance(nxt, Activation):
                dense.append((i, L, nxt))
                i += 2
            else:
                dense.append((i, L, None))
                i += 1
        if not 2 <= len(dense) <= B.MLP_MAX_LAYERS:
            return False
        hi, head, hact = dense[-1]
        if hact is None or type(hact.activation_function).__name__ != "Softmax" or head.units > 32 or hi != n - 2:
            return False
        for li, L, act in dense[:-1]:
            if act is not None and (not act._fusable or not self.fwd_fused[li + 1]):
                return False
            if L.units > 4096:
                return False
        if any(L._weights is None or L._bias is None or L._d_weights is None for _, L, _ in dense):
            return False
        self._ensure_targets()
        nl = len(dense)
        desc = B.MlpStep()
        keep = []
        s0 = self.slot()
        chain = (s0,)
        # optimizer table in update order (last layer first), as the layer-by-layer program builds it
        ups = [(li, L._weights, L._d_weights, L._bias, L._d_bias, ()) for li, L, _ in reversed(dense)]
        oplan = opt._make_plan(ups, self.ss, 1, nl)
        sync = torch.zeros(64, dtype=torch.int32, device=B.device())   # barrier words + phase time stamps
        errs = [None] * nl
        errs[nl - 1] = self.buf(self.pred.shape)
        for l in range(nl - 1, -1, -1):
            li, L, act = dense[l]
            d = desc.layer[l]
            out = self.pred if l == nl - 1 else self.outs[li + 1 if act is not None else li]
            K, N = int(L._weights.shape[0]), int(L._weights.shape[1])
            kind, alpha = (ops.ACT_NONE, 0.0) if (l == nl - 1 or act is None) else (act.activation_function._kind,
                                                                                 float(act.activation_function._alpha))
            d.w, d.b, d.dw, d.db = L._weights.data_ptr(), L._bias.data_ptr(), L._d_weights.data_ptr(), L._d_bias.data_ptr()
            d.out, d.err = out.data_ptr(), errs[l].data_ptr()
            d.K, d.N, d.act, d.alpha = K, N, int(kind), alpha
            d.chain = ops.pack_chain(chain)
            j = nl - 1 - l                      # position in the optimizer table: (w, b) pairs, last layer first
            d.gi_w, d.gi_b = 2 * j, 2 * j + 1
            d.s_out0 = d.s_out1 = -1
            if l > 0:
                pli, pL, pact = dense[l - 1]
                errs[l - 1] = self.buf((self.B, K))
                d.s_out0 = self.slot()
                chain = (d.s_out0,)
                if pact is not None:
                    d.s_out1 = self.slot()
                    chain = (d.s_out0, d.s_out1)
            L._grad_pending = None
            keep.append(out)
        desc.L, desc.B, desc.n_slots = nl, self.B, self.n_slots
        desc.n_params, desc.n_layers_opt, desc.total_blocks = oplan.n_params, nl, oplan.total_blocks
        desc.x, desc.y = self.x_in.data_ptr(), self.y_in.data_ptr()
        desc.ss, desc.acc, desc.gss = self.ss.data_ptr(), self.acc.data_ptr(), oplan.gss.data_ptr()
        desc.hyper, desc.t_counter = opt._hyper.data_ptr(), opt._t_counter().data_ptr()
        desc.params, desc.block_map, desc.sync = oplan.desc.data_ptr(), oplan.block_map.data_ptr(), sync.data_ptr()
        desc.s0 = s0
        self._mlp_keep = (desc, oplan, sync, errs, keep)
        io = (self.x_in, self.y_in) + tuple(t for _, L, _ in dense for t in (L._weights, L._d_weights)) + opt._plan_io(oplan)
        self.emit(lambda: ops.mlp_train_step(desc), "mlp_train_step", io)
        self.train_err0 = errs[nl - 1]
        return True

    def _loss_kind(self):
        lf = self.model.loss_function
        if lf is None or lf._kind is None:
            raise ValueError("the model must be compiled with a supported loss before training")
        last = self.layers[-1]
        fused = False
        if isinstance(last, Activation):
            name = type(last.activation_function).__name__
            # models.py:200-210: the recognised pairs, and SparseCategoricalCrossentropy behind ANY last activation
            fused = (name == "Softmax" and isinstance(lf, CategoricalCrossentropy)) or \
                    (name == "Sigmoid" and isinstance(lf, BinaryCrossentropy)) or bool(getattr(lf, "_sparse", False))
        return lf._kind, fused

    def _ensure_targets(self):
        if self.y_in is None:
            self.y_in = B.empty(tuple(self.pred.shape))

    def program(self, kind: str):
        """kind: 'forward' | 'train' (fwd+loss+bwd+update) | 'loss' (loss value only, evaluate) |
        'backward' / 'backward_gan' / 'backward_compute_only' (public backward_pass on an external error)."""
        if kind in self.programs:
            return self.programs[kind]
        prog = _Program(kind)
        self._emit_to = prog
        model = self.model
        if kind == "train" and self._try_mlp_step(prog):
            self._emit_to = None
            self.programs[kind] = prog
            return prog
        if kind == "train":
            self._ensure_targets()
            opt = model.optimizer
            lk, fused = self._loss_kind()
            counter = opt._t_counter()
            begin_at = len(prog.ops)
            prog.ops.append(None)  # placeholder for step_begin (needs the number of updated layers)
            prog.meta.append(("step_begin", 0))
            head = getattr(self, "_head", None) if (fused and lk == ops.LOSS_CCE) else None
            n_fwd = head["op0"] if head is not None else len(self.programs["forward"].ops)
            prog.ops.extend(self.programs["forward"].ops[:n_fwd])
            prog.meta.extend(self.programs["forward"].meta[:n_fwd])
            err0 = self.buf(self.pred.shape)
            s0 = self.slot()
            pred, y_in, acc, ss = self.pred, self.y_in, self.acc, self.ss
            head_state = None
            if head is not None:   # Dense + Softmax + CCE + (p - y) in one launch (+ the head's weight-gradient partials)
                hx, hw, hb = head["x"], head["layer"]._weights, head["layer"]._bias
                n_part = ops.dense_head_partials(hx.shape[0], hw.shape[0])
                dwp = self.buf((n_part, (hw.shape[0] + 1) * hw.shape[1])) if n_part > 0 else None
                self.emit(lambda: ops.dense_softmax_cce(hx, hw, hb, y_in, pred, err0, acc, ss, s0, dwp), "dense_softmax_cce",
                          (hx, hw, y_in, pred, err0))
                if dwp is not None:
                    head_state = self.state[self.layers.index(head["layer"])]
                    head_state["dw_partial"] = dwp
            else:
                world, lparam = self.world, float(getattr(model.loss_function, "_param", 1.0))
                self.emit(lambda: ops.loss_fwd_bwd(lk, fused, pred, y_in, err0, acc, ss, s0, world, lparam), "loss_fwd_bwd",
                          (pred, y_in, err0))
            if self.dp is not None:
                self.dp.flatten_grads(self.layers)
            if self.defer_clips:
                self._acc_pending, self._slots_synced = True, s0   # loss accumulators + every slot: ONE exchange after backward
            elif self.dp is not None:
                if s0 == 0:
                    self.allreduce(self._ssacc[0:5], "loss_exchange")      # loss numerator, correct count, ||err||^2
                else:
                    self.allreduce(self.acc, "loss_exchange")
                    self.allreduce(self.ss[s0:s0 + 1], "slot_exchange")
                self._slots_synced = self.n_slots
            self._early_setup(True)
            self._merge_final = self.defer_clips and self.dp is not None
            updates, _ = self._build_backward(_Err(err0, (s0,)), False, False, False)
            self._merge_final = False
            if head_state is not None:
                del head_state["dw_partial"]   # other programs of this plan (public backward_pass) run the plain Dense backward
            self._grad_exchange(updates)
            n_early = self._early_join()
            ups = []
            for li in updates[n_early:]:
                L = self.layers[li]
                ups.append((li, L._weights, L._d_weights, L._bias, L._d_bias, self._upd_chain.get(li, ())))
            n_upd = len(updates)
            if ups:
                oplan = opt._make_plan(ups, self.ss, n_early + 1, n_upd)
                self.emit(lambda: opt._run_plan(oplan), "optimizer_multi", opt._plan_io(oplan))
            self._early = None
            prog.ops[begin_at] = lambda: ops.step_begin(ss, acc, counter, n_upd)
            self.train_err0 = err0
        elif kind == "loss":
            self._ensure_targets()
            lk, fused = self._loss_kind()
            pred, y_in, acc = self.pred, self.y_in, self.acc
            self.emit(lambda: ops.step_begin(None, acc, None, 0), "step_begin")
            lparam = float(getattr(model.loss_function, "_param", 1.0))
            self.emit(lambda: ops.loss_fwd_bwd(lk, fused, pred, y_in, None, acc, None, None, 1, lparam), "loss_fwd_bwd", (pred, y_in))
        elif kind in ("backward", "backward_gan", "backward_compute_only"):
            gan, compute_only = kind == "backward_gan", kind == "backward_compute_only"
            opt = model.optimizer
            counter = opt._t_counter() if (opt is not None and not compute_only) else None
            ss, acc = self.ss, self.acc
            begin_at = len(prog.ops)
            prog.ops.append(None)
            prog.meta.append(("step_begin", 0))
            self.err_in = self.buf(self.pred.shape)
            if self.dp is not None:
                self.dp.flatten_grads(self.layers)
            last_is_act = isinstance(self.layers[-1], Activation)
            first = self.n_slots
            if last_is_act and not gan:
                # models.py:198-210: for the recognised pairs the incoming error is REPLACED by predictions - y_true
                lk, fused = self._loss_kind() if model.loss_function is not None else (None, False)
                if fused:
                    self._ensure_targets()
                    pred, y_in, err_in = self.pred, self.y_in, self.err_in
                    self.emit(lambda: ops.loss_fwd_bwd(lk, 1, pred, y_in, err_in, None, None, None), "loss_fwd_bwd", (pred, y_in, err_in))
                s0 = self.slot()
                err_in = self.err_in
                self.emit(lambda: ops.sumsq(err_in, ss, s0), "sumsq", (err_in,))
                start = _Err(self.err_in, (s0,))
            elif last_is_act and gan:
                start = _Err(self.err_in, ())
            else:
                s0 = self.slot()
                err_in = self.err_in
                self.emit(lambda: ops.sumsq(err_in, ss, s0), "sumsq", (err_in,))
                start = _Err(self.err_in, (s0,))
            self._sync_slots(first)
            self._early = None
            updates, in_err = self._build_backward(start, gan, compute_only, True)
            self._grad_exchange(updates)
            # materialise the returned input error: clip folded in (the reference clips before Input too)
            self.in_err_out = self.buf(in_err.t.shape)
            iet, chain, out = in_err.t, in_err.chain, self.in_err_out
            self.emit(lambda: ops.scale_apply(iet, ss, chain, out), "scale_apply", (iet, out))
            ups = [(li, self.layers[li]._weights, self.layers[li]._d_weights, self.layers[li]._bias, self.layers[li]._d_bias,
                    self._upd_chain.get(li, ())) for li in updates]
            n_upd = len(ups)
            if n_upd:
                oplan = opt._make_plan(ups, self.ss)
                self.emit(lambda: opt._run_plan(oplan), "optimizer_multi", opt._plan_io(oplan))
            prog.ops[begin_at] = lambda: ops.step_begin(ss, None, counter, n_upd)
        else:
            raise KeyError(kind)
        self._emit_to = None
        _ = self.ws  # allocate the workspace before any capture
        self.programs[kind] = prog
        return prog

    # ---- host <-> device staging -----------------------------------------------------------------------------------
    @staticmethod
    def _upload(dst: torch.Tensor, stage_attr, owner, src):
        """Copy a host array (any float dtype) or a device tensor into the static input buffer `dst`.
        float32 C-contiguous arrays are copied straight from their own memory (asynchronously when it is pinned);
        anything else is converted into a pinned staging buffer first."""
        if isinstance(src, torch.Tensor):
            if src.is_cuda:
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src.reshape(dst.shape))
                return
            src_t = src
        else:
            a = src if type(src) is np.ndarray else np.asarray(src)
            if a.dtype == np.float32 and a.flags.c_contiguous and a.size == dst.numel():
                # straight from the caller's buffer (asynchronous when it is pinned): one cudaMemcpyAsync through the C ABI
                B.check(B.lib().b200nn_upload(dst.data_ptr(), a.__array_interface__["data"][0], a.nbytes, B.stream()))
                return
            if a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable:
                src_t = torch.from_numpy(a)
            else:
                stage = getattr(owner, stage_attr, None)
                if stage is None:
                    stage = torch.empty(tuple(dst.shape), dtype=torch.float32, pin_memory=True)
                    setattr(owner, stage_attr, stage)
                evt = getattr(owner, stage_attr + "_evt", None)
                if evt is not None:
                    evt.synchronize()   # the previous async copy out of this staging buffer must have finished
                np.copyto(stage.numpy(), a.reshape(tuple(stage.shape)), casting="unsafe")
                dst.copy_(stage, non_blocking=True)
                evt = torch.cuda.Event()
                evt.record(B.stream_obj())
                setattr(owner, stage_attr + "_evt", evt)
                return
        dst.copy_(src_t.reshape(dst.shape), non_blocking=True)

    def read_acc(self):
        """(loss numerator, correct count) of the step just run: one D2H copy into pinned memory + stream sync."""
        if getattr(self, "_acc_host", None) is None:
            self._acc_host = torch.zeros(2, dtype=torch.float64).pin_memory()
            self._acc_host_np = self._acc_host.numpy()
        B.check(B.lib().b200nn_read_sync(self._acc_host.data_ptr(), self.acc.data_ptr(), 16, B.stream()))
        if self.dp is not None and self.dp.comm.peer and self.dp.comm.timeouts():
            # a peer exchange gave up after 10 s (a rank never arrived): the sums of this step are garbage, replicas have diverged
            raise B.B200NNError("data-parallel peer exchange timed out: a rank did not arrive; results of this step are invalid")
        return float(self._acc_host_np[0]), float(self._acc_host_np[1])

    def set_input(self, x):
        self._upload(self.x_in, "x_stage", self, x)

    def set_target(self, y):
        self._ensure_targets()
        self._upload(self.y_in, "y_stage", self, y)

    def destroy(self):
        for p in self.programs.values():
            p.destroy()
