"""One VO optimisation step entirely on the device, without autograd or torch.optim in the loop:

    forward program -> loss + d(loss)/d(pred) -> backward program -> [NCCL all-reduce of the flat gradient
    bucket] -> Adam over the flat parameter bucket

Semantics follow the reference's engine: per-delta mean((gt - pred)^2 * w) summed over dx/dz/dyaw
(vo/engine/vo_cnn_engine.py:135-198) -- one such mean PER DATA TYPE when the batch carries geometric-invariance rows
(vo_cnn_regression_geo_invariance_engine.py:676-750), plus loss_inv_weight times the inversion loss over the TURN rows
(:781-792) -- optimiser Adam(lr=2.5e-4, eps=1e-8, weight_decay=0)
(vo_cnn_regression_geo_invariance_engine.py:122-133, configs/vo/vo_pointnav.yaml:36-40).
Data-parallel training (new capability; the reference trains VO on one GPU): every rank
holds a replica, the flat gradient bucket is summed with ONE all-reduce and scaled by 1/world inside the
loss-gradient kernel, which equals the single-GPU gradient of the mean loss over the concatenated batch
when every rank has the same batch size.
"""
import os

import torch

from ... import lib as L


# where the side-stream input pipeline of the next batch starts: "start" = with the step (under the stem / layer1 forward),
# "bwd" = after the forward program (under the backward pass).  Measured: see DESIGN.md section 3.
_PREFETCH_AT = os.environ.get("PNVO_PREFETCH_AT", "start")


class PrefetchedBatches:
    """Double-buffered host -> device staging of training batches (the device-side replacement of the reference's
    `_transfer_batch`, vo_cnn_regression_geo_invariance_engine.py:283-353, which copies a list of unpinned CPU
    tensors synchronously in front of every forward).  `submit()` enqueues the copy of the NEXT batch from pinned
    host memory on a side stream; `acquire()` makes the compute stream wait for the oldest submitted batch and
    returns its device tensors; `release()` lets the copy stream reuse the slot.  With two slots the PCIe copy of
    batch i+1 overlaps the kernels of step i."""

    def __init__(self, example, device, slots=2):
        self.dev = torch.device(device)
        self.stream = torch.cuda.Stream(self.dev)
        self.bufs = [{k: torch.empty(v.shape, dtype=v.dtype, device=self.dev) for k, v in example.items()}
                     for _ in range(slots)]
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.free = [torch.cuda.Event() for _ in range(slots)]
        self.head = self.tail = 0  # next slot to fill / next slot to hand out
        self.pending = 0
        self.bytes_per_batch = sum(v.numel() * v.element_size() for v in example.values())

    def submit(self, host_batch):
        assert self.pending < len(self.bufs), "all staging slots are in flight: acquire()/release() one first"
        s = self.head
        self.stream.wait_event(self.free[s])
        with torch.cuda.stream(self.stream):
            for k, v in host_batch.items():
                if not v.is_pinned():
                    raise L.PnvoError(f"host batch tensor {k!r} must live in pinned memory")
                self.bufs[s][k].copy_(v, non_blocking=True)
            self.ready[s].record(self.stream)
        self.head = (s + 1) % len(self.bufs)
        self.pending += 1

    def acquire(self):
        assert self.pending > 0, "acquire() without a submitted batch"
        s = self.tail
        torch.cuda.current_stream(self.dev).wait_event(self.ready[s])
        self._held = s
        return self.bufs[s]

    def peek_next(self):
        """(device tensors, ready event) of the batch AFTER the one currently held, if it has been submitted: lets the
        trainer start that batch's input pipeline on a side stream as soon as its copy has landed."""
        if self.pending < 2:
            return None
        s = (self._held + 1) % len(self.bufs)
        return self.bufs[s], self.ready[s]

    def release(self):
        s = self._held
        self.free[s].record(torch.cuda.current_stream(self.dev))
        self.tail = (s + 1) % len(self.bufs)
        self.pending -= 1


class FusedVOTrainStep:
    def __init__(self, model, lr=2.5e-4, betas=(0.9, 0.999), eps=1e-8, loss_weights=(1.0, 1.0, 1.0),
                 process_group=None, loss_inv_weight=0.0, move_forward_id=1):
        """loss_inv_weight > 0 adds the geometric-inversion loss (VO.GEOMETRY.loss_inv_weight,
        vo_cnn_regression_geo_invariance_engine.py:367-449,781-792); step() then needs the per-row action ids and
        data types (CUR_REL_TO_PREV / PREV_REL_TO_CUR, vo/dataset/geo_invariance.make_pair_map emits both)."""
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.loss_weights = tuple(float(w) for w in loss_weights)
        self.loss_inv_weight, self.move_forward_id = float(loss_inv_weight), int(move_forward_id)
        self.group = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.step_count = 0
        self.use_peer_exchange = os.environ.get("PNVO_PEER_ADAM", "1") != "0"   # read when the flat buckets are created
        self._flat = None
        self._plan = None
        # double-buffered input staging (step(..., prefetch=next_obs)): side stream, per-buffer events
        self._side = None
        self._parity = 0
        self._prefetched = None   # (obs identity, parity) prepared ahead on the side stream
        self._staged = [None, None]   # event: input buffer `parity` holds a prepared batch

    # parameters are re-pointed into one flat fp32 buffer ordered like the plan's gradient bucket, so the
    # optimiser and the all-reduce each touch a single contiguous range
    def _flatten(self, plan):
        names = plan.param_names()
        self._names = list(names)
        P = dict(self.model.named_parameters())
        missing = set(P) - set(names)
        assert not missing, f"parameters without a gradient slot: {sorted(missing)}"
        n = sum(P[k].numel() for k in names)
        dev = plan.dev
        self._peer = None
        if 2 <= self.world <= 8 and self.use_peer_exchange \
                and torch.distributed.get_backend(self.group) == "nccl":
            # gradient exchange fused with Adam over NVLink peer memory (csrc/peer_reduce.cu): parameters and gradients live
            # in a peer-mapped region; the NCCL all-reduce + adam_kernel pair stays as the path for other world sizes
            from ...parallel_utils import PeerBuckets

            try:
                self._peer = PeerBuckets(n, dev, self.group)
            except L.PnvoError as e:
                import warnings

                warnings.warn(f"peer-memory gradient exchange unavailable ({e}); using the NCCL all-reduce")
        if self._peer is not None:
            flat = self._peer.params[:n]
            self.model._grad_bucket = self._peer.grads
            n_state = self._peer.n_pad
        else:
            flat = torch.empty(n, dtype=torch.float32, device=dev)
            n_state = n
        off = 0
        for k in names:
            m = P[k].numel()
            flat[off:off + m].copy_(P[k].data.reshape(-1))
            P[k].data = flat[off:off + m].view(P[k].shape)
            off += m
        self._flat = flat
        self._m = torch.zeros(n_state, dtype=torch.float32, device=dev)
        self._v = torch.zeros(n_state, dtype=torch.float32, device=dev)
        self._loss = torch.zeros(3, dtype=torch.float32, device=dev)  # [total, inversion rot, inversion pos]
        if self.world > 1:  # replicas start from rank 0's weights (as DDP does, ddppo.py:55-58)
            torch.distributed.broadcast(flat, 0, group=self.group)

    # ---- optimiser state in torch.optim.Adam's format (the reference checkpoints `optim_states`,
    # vo_cnn_regression_geo_invariance_engine.py:1425-1434): a run trained on the fused path resumes with its moments and
    # bias-correction step, and the file loads into a plain torch.optim.Adam over model.parameters() as well
    def _full_moments(self):
        m, v = self._m, self._v
        if getattr(self, "_peer", None) is not None:
            # every rank holds the moments of its own slice only (zeros elsewhere): a sum gives the whole vectors
            m, v = m.clone(), v.clone()
            torch.distributed.all_reduce(m, group=self.group)
            torch.distributed.all_reduce(v, group=self.group)
        return m, v

    def state_dict(self):
        if self._flat is None:
            raise L.PnvoError("state_dict() before the first step: there is no optimiser state yet")
        m, v = self._full_moments()
        order = [k for k, _ in self.model.named_parameters()]
        state, off = {}, 0
        sizes = dict(self.model.named_parameters())
        spans = {}
        for k in self._names:
            n = sizes[k].numel()
            spans[k] = (off, off + n)
            off += n
        for i, k in enumerate(order):
            lo, hi = spans[k]
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": m[lo:hi].view(sizes[k].shape).clone(),
                        "exp_avg_sq": v[lo:hi].view(sizes[k].shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(order)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd, example_obs=None):
        """sd: state_dict() of this class or of a torch.optim.Adam over model.parameters().  The flat buckets exist after the
        first plan is built: pass `example_obs` (a batch of the training shape) when loading before the first step."""
        if self._flat is None:
            if example_obs is None:
                raise L.PnvoError("load_state_dict() before the first step needs example_obs to build the flat buckets")
            self._get_plan(example_obs)
        order = [k for k, _ in self.model.named_parameters()]
        sizes = dict(self.model.named_parameters())
        off, spans = 0, {}
        for k in self._names:
            spans[k] = (off, off + sizes[k].numel())
            off += sizes[k].numel()
        self._m.zero_()
        self._v.zero_()
        lo_own, hi_own = (0, self._m.numel()) if getattr(self, "_peer", None) is None else self._peer.slice_range()
        steps = set()
        for i, k in enumerate(order):
            st = sd["state"].get(i)
            if st is None:
                continue
            lo, hi = spans[k]
            self._m[lo:hi].copy_(st["exp_avg"].reshape(-1))
            self._v[lo:hi].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if getattr(self, "_peer", None) is not None:  # keep the owned slice only
            self._m[:lo_own].zero_(); self._m[hi_own:].zero_()
            self._v[:lo_own].zero_(); self._v[hi_own:].zero_()
        if len(steps) > 1:
            raise L.PnvoError(f"optimizer state with differing step counts {sorted(steps)}: not a single-group Adam")
        self.step_count = steps.pop() if steps else 0
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])

    def _get_plan(self, obs):
        model = self.model
        plan = model._plan_for(obs, True, model.training)
        if self._flat is None:
            self._flatten(plan)
            model._plans.clear()
            plan = model._plan_for(obs, True, model.training)  # rebuilt on the flat storage
        if plan is not self._plan:
            B, O = plan.B, plan.head["out_dim"]
            dev = plan.dev
            self._target = torch.zeros(B, O, dtype=torch.float32, device=dev)
            self._actions = torch.zeros(B, dtype=torch.int64, device=dev)
            self._data_types = torch.zeros(B, dtype=torch.int64, device=dev)
            self._dz_mask = torch.ones(B, dtype=torch.float32, device=dev)
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
            self._loss_progs = {}
            self._plan = plan
        return plan

    def _loss_prog_for(self, plan, typed, masked):
        """Loss + d(loss)/d(pred) program: `typed` = per-data-type means (+ the inversion loss over the turn rows when
        loss_inv_weight > 0); untyped batches with loss_inv_weight > 0 are taken as interleaved pairs (every row)."""
        key = (typed, masked)
        prog = self._loss_progs.get(key)
        if prog is None:
            B, O = plan.B, plan.head["out_dim"]
            types = self._data_types if typed else None
            ops = [L.op_mse_loss(plan.out, self._target, self._dz_mask if masked else None, plan.dout, self._loss, B, O,
                                 self.loss_weights, 1.0 / self.world, data_types=types)]
            if self.loss_inv_weight > 0:
                ops.append(L.op_geo_inv_loss(plan.out, self._actions, plan.dout, self._loss, B, O, self.loss_inv_weight,
                                             1.0 / self.world, self.move_forward_id, data_types=types, err=self._err))
            prog = self._loss_progs[key] = L.Program(ops)
        return prog

    @staticmethod
    def _check_pairing(actions, data_types, turn_ids=(2, 3)):
        """The reference's assertion (vo_cnn_regression_geo_invariance_engine.py:373-374) on host-side metadata: the TURN
        rows, in batch order, alternate [cur-rel-to-prev, prev-rel-to-cur]."""
        import numpy as np

        a = np.asarray(actions).reshape(-1)
        t = np.asarray(data_types).reshape(-1)
        v = t[(a == turn_ids[0]) | (a == turn_ids[1])]
        if v.size % 2 or (v[0::2] != 0).any() or (v[1::2] != 1).any():
            raise L.PnvoError("geometric-inversion loss: the TURN rows of the batch must alternate "
                              "[cur_rel_to_prev, prev_rel_to_cur, ...] (vo_cnn_regression_geo_invariance_engine.py:373-374)")

    def check(self):
        """Raises if a device-side pairing check of an earlier step failed (only CUDA-resident metadata is checked on the
        device; numpy / CPU metadata is validated on the host inside step())."""
        if self._plan is not None and int(self._err.item()) != 0:
            raise L.PnvoError("geometric-inversion loss: TURN rows did not alternate [cur_rel_to_prev, prev_rel_to_cur]")

    def _prefetch(self, plan, obs, parity, ready=None):
        """Input pipeline (top-down, statistics, assembly) of a FUTURE batch into staging buffer `parity`, on the side
        stream: it only depends on the data, not on the weights, so it overlaps the current step's kernels."""
        dev = plan.dev
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        # Everything the compute stream has enqueued so far comes first: the step that last read staging buffer `parity`
        # AND -- when this step fell back to the non-prefetched path -- the current batch's own input pipeline, which
        # shares plan.in_stats / in_scale / in_shift / td_pair and the RunningMeanAndVar buffers with this one.
        self._side.wait_stream(torch.cuda.current_stream(dev))
        if ready is not None:
            self._side.wait_event(ready)  # e.g. the host->device copy of that batch
        with torch.cuda.stream(self._side):
            self.model._run_forward_raw(plan, obs, self.model.training, parity=parity, prepare_only=True)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._staged[parity] = ev
        self._prefetched = (tuple(id(v) for v in obs.values()), parity)

    def step(self, obs, target, actions=None, prefetch=None, prefetch_ready=None, data_types=None, dz_regress_masks=None):
        """obs: dict of NHWC fp32 CUDA tensors (the model's forward input), or the raw pairs
        {"rgb": uint8 [B,H,W,6], "depth": fp32 [B,H,W,2]} (derived channels computed on the device);
        target: [B, 3] fp32 CUDA.  prefetch: the NEXT step's raw pairs -- their input pipeline is started on a side
        stream into the alternate staging buffer and overlaps this step (the next call must pass the same tensors as
        `obs`); prefetch_ready: optional event the side stream waits for first (the batch's host->device copy).
        actions / data_types: per-row action ids and CUR_REL_TO_PREV / PREV_REL_TO_CUR flags (numpy, CPU or CUDA);
        dz_regress_masks: optional [B] weights of the dz term (vo_cnn_engine.py:160-166).
        Returns the (device) loss tensor of this rank's batch."""
        model = self.model
        plan = self._get_plan(obs)
        dev = plan.dev
        main = torch.cuda.current_stream(dev)
        self._target.copy_(target, non_blocking=True)

        def put(dst, src):
            if not isinstance(src, torch.Tensor):
                src = torch.as_tensor(src)
            dst.copy_(src.reshape(-1), non_blocking=True)

        typed = data_types is not None
        if self.loss_inv_weight > 0:
            if actions is None:
                raise L.PnvoError("the geometric-inversion loss needs the per-row action ids")
            if typed and not (isinstance(actions, torch.Tensor) and actions.is_cuda):
                dt = data_types.cpu() if isinstance(data_types, torch.Tensor) else data_types
                self._check_pairing(actions, dt)
            put(self._actions, actions)
        if typed:
            put(self._data_types, data_types)
        if dz_regress_masks is not None:
            put(self._dz_mask, dz_regress_masks)
        key = tuple(id(v) for v in obs.values())
        if self._prefetched is not None and self._prefetched[0] == key and model._is_raw(obs):
            parity = self._prefetched[1]
            main.wait_event(self._staged[parity])   # inputs were assembled ahead of time
            self._prefetched = None
            late = _PREFETCH_AT == "bwd"   # start the next batch's input pipeline under the backward pass instead
            if prefetch is not None and not late:
                self._prefetch(plan, prefetch, 1 - parity, prefetch_ready)
            model._run_backbone(plan, parity)
            if prefetch is not None and late:
                self._prefetch(plan, prefetch, 1 - parity, prefetch_ready)
        else:
            parity = self._parity
            self._prefetched = None
            if model._is_raw(obs):
                model._run_forward_raw(plan, obs, model.training, parity=parity, prepare_only=True)
                if prefetch is not None:
                    self._prefetch(plan, prefetch, 1 - parity, prefetch_ready)
                model._run_backbone(plan, parity)
            else:
                parity = 0
                model._run_forward(plan, obs, model.training)
        self._parity = 1 - parity if prefetch is not None else parity
        self._loss_prog_for(plan, typed, dz_regress_masks is not None).run(dev)
        plan.programs_for(parity)[1].run(dev)
        self._exchange_and_update(plan)
        # the optimiser wrote through raw pointers: tell the module its packed fp16 weights are stale
        model._packed_version = None
        return self._loss[:1]

    def _exchange_and_update(self, plan):
        """Gradient exchange between the replicas + the Adam step (everything after the backward program)."""
        dev = plan.dev
        self.step_count += 1
        if os.environ.get("PNVO_DIAG_NO_EXCHANGE") == "1":   # timing diagnosis only: replicas diverge
            L.run_ops([L.op_adam(self._flat, plan.grad_flat, self._m, self._v, self._flat.numel(), self.step_count, self.lr,
                                 self.betas[0], self.betas[1], self.eps)], dev)
            return
        if self._peer is not None:
            # ONE kernel: reduce-scatter of the gradients over NVLink + Adam on this rank's slice + all-gather of the
            # updated parameters into every replica
            self._peer.reduce_adam(self._m, self._v, self.step_count, self.lr, self.betas[0], self.betas[1], self.eps)
        else:
            if self.world > 1:
                torch.distributed.all_reduce(plan.grad_flat, group=self.group)
            L.run_ops([L.op_adam(self._flat, plan.grad_flat, self._m, self._v, self._flat.numel(), self.step_count, self.lr,
                                 self.betas[0], self.betas[1], self.eps)], dev)
