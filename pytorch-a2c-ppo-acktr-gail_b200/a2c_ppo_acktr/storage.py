"""Rollout buffer with the reference's public surface, backed by sm_100a kernels.

Drop-in for `a2c_ppo_acktr.storage.RolloutStorage` (reference storage.py:9-202):
same constructor, same public tensor attributes ([T(+1), N, ...] layout), same
`insert / after_update / compute_returns / feed_forward_generator /
recurrent_generator` signatures, same assertion messages, same consumption of the
global CPU RNG (one `randperm` per epoch, drawn when the generator is first
advanced).  `compute_returns` is one reverse-scan kernel instead of a Python loop
over time; the generators are one fused gather launch per minibatch instead of
eight `aten::index` calls.

Two opt-in extensions: `obs_dtype=torch.uint8` stores Atari frames as bytes (a
quarter of the HBM footprint and gather traffic); the policy kernels normalise
bytes to fp32/255 while loading.  The default (fp32) keeps the reference contract.
`stage_host(...)` ingests a whole rollout collected on the host (pinned memory): the
narrow tensors are copied at once, the frames are pulled on a side stream --
per minibatch environment set by the recurrent PPO update, so that the upload of
later minibatches overlaps the compute of earlier ones.
`frame_ring=True` (uint8 Atari stacks only) stores every 84 x 84 frame ONCE: the reference's
`VecPyTorchFrameStack` (envs.py:218-259) rolls a 4-frame stack, so consecutive observations
share three of their four frames.  `frames[t + c]` is channel c of observation slot t and
`stack_valid[t]` (1..4) says how many of the newest frames of slot t are real (the stack is
cleared when an episode ends) -- a quarter of the HBM footprint and of the host upload; the
update's frame re-layout kernel reads the ring directly (csrc/tcx_aux.cu).  `obs` stays
readable / assignable per slot for the reference's call sites.
"""
import torch

from . import _native


def _space_is_discrete(action_space):
    # duck-typed by class NAME, exactly as the reference does (storage.py:19,24)
    return action_space.__class__.__name__ == 'Discrete'


class _RingSlot(torch.Tensor):
    """Materialised stack of ONE observation slot of a frame-ring rollout: a plain tensor for every reader, whose
    in-place `copy_` writes through to the ring (`rollouts.obs[0].copy_(obs)`, reference main.py:97)."""

    @staticmethod
    def make(data, storage, slot):
        t = torch.Tensor._make_subclass(_RingSlot, data)
        t._ring, t._slot = storage, slot
        return t

    def copy_(self, src, non_blocking=False):
        self._ring._write_stack(self._slot, src)
        return super(_RingSlot, self).copy_(src, non_blocking)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        with torch._C.DisableTorchFunctionSubclass():
            out = func(*args, **(kwargs or {}))
        return out


class _RingObs(object):
    """`rollouts.obs` of a frame-ring rollout: indexable like the [T + 1, N, C, H, W] tensor it stands for."""

    def __init__(self, storage):
        self._s = storage

    def __getitem__(self, key):
        s = self._s
        n_slots = s.stack_valid.shape[0]
        if isinstance(key, int):
            t = key + n_slots if key < 0 else key
            return _RingSlot.make(s._stacks(t, t + 1)[0], s, t)
        if isinstance(key, slice):
            lo, hi, step = key.indices(n_slots)
            assert step == 1, "frame ring: only contiguous slot ranges"
            return s._stacks(lo, hi)
        raise TypeError("frame ring: index observation slots with an int or a slice")

    def __len__(self):
        return self._s.stack_valid.shape[0]

    def size(self, *a):
        return self.shape if not a else self.shape[a[0]]

    @property
    def shape(self):
        s = self._s
        return torch.Size((s.stack_valid.shape[0], s.frames.shape[1], s._ring_c) + tuple(s.frames.shape[2:]))

    @property
    def dtype(self):
        return torch.uint8

    @property
    def device(self):
        return self._s.frames.device

    def data_ptr(self):
        return self._s.frames.data_ptr()

    def fill_(self, v):
        self._s.frames.fill_(v)
        self._s.stack_valid.fill_(self._s._ring_c)
        return self


class RolloutStorage(object):
    _MOVABLE = ('obs', 'recurrent_hidden_states', 'rewards', 'value_preds', 'returns',
                'action_log_probs', 'actions', 'masks', 'bad_masks')

    def __init__(self, num_steps, num_processes, obs_shape, action_space,
                 recurrent_hidden_state_size, obs_dtype=torch.float32, frame_ring=False):
        T, N = num_steps, num_processes
        self.frame_ring = bool(frame_ring)
        if self.frame_ring:
            if len(obs_shape) != 3 or obs_dtype != torch.uint8:
                raise ValueError("frame_ring stores uint8 image stacks [C, H, W]")
            self._ring_c = int(obs_shape[0])
            self.frames = torch.zeros(T + self._ring_c, N, *obs_shape[1:], dtype=torch.uint8)
            self.stack_valid = torch.ones(T + 1, N, dtype=torch.uint8)
            self._MOVABLE = ('frames', 'stack_valid') + tuple(k for k in RolloutStorage._MOVABLE if k != 'obs')
        else:
            self.obs = torch.zeros(T + 1, N, *obs_shape, dtype=obs_dtype)
        self.recurrent_hidden_states = torch.zeros(T + 1, N, recurrent_hidden_state_size)
        self.rewards = torch.zeros(T, N, 1)
        self.value_preds = torch.zeros(T + 1, N, 1)
        self.returns = torch.zeros(T + 1, N, 1)
        self.action_log_probs = torch.zeros(T, N, 1)
        if _space_is_discrete(action_space):
            self.actions = torch.zeros(T, N, 1, dtype=torch.int64)
        else:
            self.actions = torch.zeros(T, N, action_space.shape[0])
        self.masks = torch.ones(T + 1, N, 1)
        # 0 where an episode ended because of a time limit rather than a true terminal
        self.bad_masks = torch.ones(T + 1, N, 1)

        self.num_steps = T
        self.step = 0
        self._idx_cache = None
        self._staged_obs = None          # pinned host frames whose upload has not been started yet
        self._copy_stream = None
        self._staged_tail = None         # (device, pinned host) tensors whose rows [1:] follow the frames on the copy stream
        self._tail_event = None

    # -- frame ring --------------------------------------------------------
    def __getattr__(self, name):
        # `obs` of a frame-ring rollout is a view object (the attribute only exists in stacked mode)
        if name == "obs" and self.__dict__.get("frame_ring"):
            return _RingObs(self)
        raise AttributeError(name)

    def _stacks(self, lo, hi):
        """Materialised observation stacks of slots [lo, hi): uint8 [hi - lo, N, C, H, W] (a copy)."""
        C = self._ring_c
        out = torch.stack([self.frames[lo + c:hi + c] for c in range(C)], dim=2)
        keep = torch.arange(C, device=out.device).view(1, 1, C) >= (C - self.stack_valid[lo:hi].long().unsqueeze(-1))
        return out * keep.view(hi - lo, -1, C, 1, 1).to(out.dtype)

    def _write_stack(self, slot, src):
        """Stores a full stack [N, C, H, W] at `slot` (all C frames as given: valid = C)."""
        C = self._ring_c
        src = src.to(self.frames.device)
        for c in range(C):
            self.frames[slot + c].copy_(src[:, c])
        self.stack_valid[slot].fill_(C)

    def ring_stack(self, slot):
        """Stack of `slot` in a persistent staging tensor (stable address: Policy.act_into replays CUDA graphs)."""
        st = self.__dict__.get("_ring_stage")
        if st is None or st.device != self.frames.device:
            st = self._ring_stage = torch.zeros((self.frames.shape[1], self._ring_c) + tuple(self.frames.shape[2:]),
                                                dtype=torch.uint8, device=self.frames.device)
        st.copy_(self._stacks(slot, slot + 1)[0])
        return st

    # -- placement ---------------------------------------------------------
    def to(self, device):
        for name in self._MOVABLE:
            setattr(self, name, getattr(self, name).to(device))

    # -- host ingest (extension; the device-side half of envs.py:167-190 `VecPyTorch`) ------------
    def stage_host(self, host):
        """`host`: dict attribute name -> PINNED host tensor of the attribute's shape and dtype (any subset of the
        buffer's tensors).  Everything except `obs` is copied now, asynchronously on the current stream (of a wide
        `recurrent_hidden_states` only row 0, the one an update reads; rows [1:] follow the frames).  The frames
        are only registered: the next `update()` (or `wait_staged()`, which every other reader of `obs` in this
        package calls) uploads them on a side stream.  The host tensors must stay untouched until then."""
        dev = self.rewards.device
        if dev.type != "cuda":
            raise _native.NativeError("stage_host: the rollout buffer must live on a CUDA device")
        for name, src in host.items():
            if name not in self._MOVABLE:
                raise KeyError("stage_host: unknown rollout tensor %r" % name)
            dst = getattr(self, name)
            if tuple(src.shape) != tuple(dst.shape) or src.dtype != dst.dtype:
                raise ValueError("stage_host: %s must be %s %s" % (name, tuple(dst.shape), dst.dtype))
            if name == ("frames" if self.frame_ring else "obs"):
                if not src.is_pinned() or not src.is_contiguous():
                    raise _native.NativeError("stage_host: %s must be a contiguous pinned host tensor" % name)
                self._staged_obs = src
            elif (name == "recurrent_hidden_states" and src.is_pinned() and dst.shape[0] > 1
                  and dst[1:].numel() * dst.element_size() >= (1 << 20)):
                # an update reads only the initial state (reference storage.py:160-161); the other T rows (tens of MB for
                # a GRU) follow the frames on the copy stream instead of holding up the launching stream
                dst[0].copy_(src[0], non_blocking=True)
                self._staged_tail = (dst, src)
            else:
                dst.copy_(src, non_blocking=True)

    def _staged_upload(self):
        """Begins the upload of the registered frames on the copy stream.  Returns None when no frames are staged, else an
        object with `columns(envs)` -- enqueues the strided copies of those environments (host int64 array, may be
        empty) and returns an event -- and `finish()` -- uploads every environment not asked for yet, then the deferred
        rows of the hidden states, and returns the closing event."""
        host, self._staged_obs = self._staged_obs, None
        if host is None:
            return None
        tail, self._staged_tail = self._staged_tail, None
        frames = self.frames if self.frame_ring else self.obs
        dev = frames.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream(dev))      # earlier readers / writers of obs are done
        storage = self
        seen = torch.zeros(frames.shape[1], dtype=torch.bool)

        class Upload(object):
            def columns(self, envs):
                envs = torch.as_tensor(envs, dtype=torch.int64)
                if envs.numel():
                    _native.upload_env_columns(frames, host, envs, cs)
                    seen[envs] = True
                ev = torch.cuda.Event()
                ev.record(cs)
                return ev

            def finish(self):
                ev = self.columns((~seen).nonzero().view(-1))
                storage._staged_keepalive = host            # until the next stage / upload replaces it
                if tail is not None:
                    storage._upload_tail(tail, cs)
                return ev
        return Upload()

    def _upload_staged(self, env_chunks=None):
        """Starts the upload of the registered frames on the copy stream, one strided copy set per chunk of
        environments (default: everything at once).  Returns one event per chunk (+ one for the environments no chunk
        named), or None if nothing is staged."""
        up = self._staged_upload()
        if up is None:
            tail, self._staged_tail = self._staged_tail, None
            if tail is not None:
                dev = self.rewards.device
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(dev)
                self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
                self._upload_tail(tail, self._copy_stream)
            return None
        events = [up.columns(c) for c in (env_chunks if env_chunks is not None else [])]
        events.append(up.finish())
        return events

    def _upload_tail(self, tail, cs):
        if tail is None:
            return
        dst, src = tail
        with torch.cuda.stream(cs):
            dst[1:].copy_(src[1:], non_blocking=True)
        self._tail_event = torch.cuda.Event()
        self._tail_event.record(cs)
        self._tail_keepalive = src

    def _wait_tail(self):
        ev, self._tail_event = self._tail_event, None
        if ev is not None:
            torch.cuda.current_stream(self.rewards.device).wait_event(ev)

    def wait_staged(self):
        """Makes the current stream wait for the staged frames (no-op when nothing is staged)."""
        events = self._upload_staged()
        if events:
            cur = torch.cuda.current_stream(self.rewards.device)
            for ev in events:
                cur.wait_event(ev)
        self._wait_tail()

    # -- actor-side writes (reference storage.py:46-64) --------------------------
    def insert(self, obs, recurrent_hidden_states, actions, action_log_probs,
               value_preds, rewards, masks, bad_masks):
        t = self.step
        if self.frame_ring:
            # the newest frame of the rolling stack; the stack restarts where the episode ended (masks == 0: the
            # wrapper cleared it before writing this frame, reference envs.py:246-249, main.py:129-134)
            C = self._ring_c
            self.frames[t + C].copy_(obs[:, -1])
            grown = torch.clamp(self.stack_valid[t].to(torch.int64) + 1, max=C)
            fresh = masks.to(self.stack_valid.device).view(-1) == 0
            self.stack_valid[t + 1].copy_(torch.where(fresh, torch.ones_like(grown), grown).to(torch.uint8))
        else:
            self.obs[t + 1].copy_(obs)
        self.recurrent_hidden_states[t + 1].copy_(recurrent_hidden_states)
        self.masks[t + 1].copy_(masks)
        self.bad_masks[t + 1].copy_(bad_masks)
        self.actions[t].copy_(actions)
        self.action_log_probs[t].copy_(action_log_probs)
        self.value_preds[t].copy_(value_preds)
        self.rewards[t].copy_(rewards)
        self.step = (t + 1) % self.num_steps

    def after_update(self):
        self.wait_staged()
        if self.frame_ring:
            C = self._ring_c
            self.frames[:C].copy_(self.frames[-C:].clone())
            self.stack_valid[0].copy_(self.stack_valid[-1])
        for buf in (() if self.frame_ring else (self.obs,)) + (self.recurrent_hidden_states, self.masks, self.bad_masks):
            buf[0].copy_(buf[-1])

    # -- (1) returns / GAE: one reverse-scan kernel (reference storage.py:66-105) --
    def compute_returns(self, next_value, use_gae, gamma, gae_lambda,
                        use_proper_time_limits=True, schedule=1):
        """`schedule` (extension): 1 = env-parallel scan, every fp32 operation in the reference's order -- BIT-IDENTICAL
        to the reference loop (the default: the drop-in contract); 2 = time-parallel warp scan, re-associated carry,
        <= 1e-5 relative (faster when N is too small to fill the machine); 0 = pick by shape (2 when N < 16384 and
        T >= 16)."""
        nv = next_value.detach().to(dtype=torch.float32).reshape(-1).contiguous()
        if nv.numel() != self.rewards.size(1):
            raise ValueError("next_value must hold one value per process")
        _native.returns_scan(self.rewards, self.value_preds, self.returns, self.masks,
                             self.bad_masks, nv, gamma, gae_lambda, use_gae,
                             use_proper_time_limits, schedule)

    # -- (2) minibatch generators -------------------------------------------------
    def _flat_sources(self, advantages):
        """[T*N, ...] views in the order of the yielded 8-tuple."""
        T, N = self.rewards.size()[0:2]
        srcs = [
            self.obs[:-1].view(T * N, *self.obs.size()[2:]),
            self.recurrent_hidden_states[:-1].view(T * N, self.recurrent_hidden_states.size(-1)),
            self.actions.view(T * N, self.actions.size(-1)),
            self.value_preds[:-1].view(T * N, 1),
            self.returns[:-1].view(T * N, 1),
            self.masks[:-1].view(T * N, 1),
            self.action_log_probs.view(T * N, 1),
        ]
        if advantages is not None:
            adv = advantages
            if adv.dtype != torch.float32 or not adv.is_contiguous():
                adv = adv.to(torch.float32).contiguous()
            srcs.append(adv.view(T * N, 1))
        return srcs

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None):
        num_steps, num_processes = self.rewards.size()[0:2]
        batch_size = num_processes * num_steps

        if mini_batch_size is None:
            assert batch_size >= num_mini_batch, (
                "PPO requires the number of processes ({}) "
                "* number of steps ({}) = {} "
                "to be greater than or equal to the number of PPO mini batches ({})."
                "".format(num_processes, num_steps, num_processes * num_steps,
                          num_mini_batch))
            mini_batch_size = batch_size // num_mini_batch

        # SubsetRandomSampler(range(B)) draws exactly one randperm(B) from the global
        # CPU generator when iteration starts; BatchSampler(drop_last=True) cuts it into
        # consecutive chunks (reference storage.py:122-125).
        perm = torch.randperm(batch_size)
        dev = self.rewards.device
        self.wait_staged()
        perm_dev = perm.to(dev, non_blocking=False)
        srcs = self._flat_sources(advantages)
        for k in range(batch_size // mini_batch_size):
            idx = perm_dev[k * mini_batch_size:(k + 1) * mini_batch_size]
            outs = [torch.empty((mini_batch_size,) + tuple(s.shape[1:]), dtype=s.dtype, device=dev)
                    for s in srcs]
            _native.gather_rows(list(zip(srcs, outs)), idx, mini_batch_size, mode=0)
            if advantages is None:
                outs.append(None)
            yield tuple(outs)

    def recurrent_generator(self, advantages, num_mini_batch):
        num_processes = self.rewards.size(1)
        assert num_processes >= num_mini_batch, (
            "PPO requires the number of processes ({}) "
            "to be greater than or equal to the number of "
            "PPO mini batches ({}).".format(num_processes, num_mini_batch))
        E = num_processes // num_mini_batch
        T, N = self.num_steps, num_processes
        perm = torch.randperm(num_processes)
        dev = self.rewards.device
        self.wait_staged()
        perm_dev = perm.to(dev)
        srcs = self._flat_sources(advantages)
        hidden_src = self.recurrent_hidden_states[0]            # [N, H]: state at t = 0 only
        for start in range(0, N, E):
            if start + E > N:
                # the reference walks perm[start + offset] past its end here
                # (storage.py:154,165 with N % num_mini_batch != 0)
                raise IndexError("index {} is out of bounds for dimension 0 with size {}".format(N, N))
            env = perm_dev[start:start + E]
            outs = []
            pairs = []
            for i, s in enumerate(srcs):
                if i == 1:
                    outs.append(None)
                    continue
                o = torch.empty((T * E,) + tuple(s.shape[1:]), dtype=s.dtype, device=dev)
                outs.append(o)
                pairs.append((s, o))
            _native.gather_rows(pairs, env, T * E, mode=1, E=E, N=N)
            h = torch.empty((E, hidden_src.size(-1)), dtype=hidden_src.dtype, device=dev)
            _native.gather_rows([(hidden_src, h)], env, E, mode=0)
            outs[1] = h
            if advantages is None:
                outs.append(None)
            yield tuple(outs)
