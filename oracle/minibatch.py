"""Oracle: minibatch index generation and row gathers of the rollout buffer.

feed-forward: reference storage.py:107-143 (BatchSampler(SubsetRandomSampler(range(B)),
mbs, drop_last=True) == consecutive chunks of ONE torch.randperm(B) drawn from the
global CPU generator; torch 2.11 torch/utils/data/sampler.py).
recurrent:    reference storage.py:145-202.
"""
import torch

FIELDS = ("obs", "recurrent_hidden_states", "actions", "value_preds", "returns", "masks",
          "action_log_probs")


def feed_forward_index_chunks(T, N, num_mini_batch=None, mini_batch_size=None):
    batch = T * N
    if mini_batch_size is None:
        assert batch >= num_mini_batch
        mini_batch_size = batch // num_mini_batch
    perm = torch.randperm(batch)
    return [perm[k * mini_batch_size:(k + 1) * mini_batch_size] for k in range(batch // mini_batch_size)]


def _rows(x, T):
    """[T(+1), N, ...] -> first T steps flattened to [T*N, ...]."""
    return x[:T].reshape(T * x.shape[1], *x.shape[2:])


def gather_feed_forward(ro, idx, advantages=None):
    """ro: dict with the storage tensors. Returns the reference's 8-tuple (storage.py:126-143)."""
    T = ro["rewards"].shape[0]
    out = [_rows(ro[name], T)[idx] for name in FIELDS]
    out.append(None if advantages is None else advantages.reshape(-1, 1)[idx])
    return tuple(out)


def recurrent_env_chunks(N, num_mini_batch):
    assert N >= num_mini_batch
    E = N // num_mini_batch
    perm = torch.randperm(N)
    chunks = []
    for start in range(0, N, E):
        if start + E > N:
            raise IndexError("recurrent_generator walks past the permutation (N %% num_mini_batch != 0)")
        chunks.append(perm[start:start + E])
    return chunks


def gather_recurrent(ro, env, advantages):
    """Rows ordered t*E + e; hidden state taken at t = 0 only (storage.py:165-199)."""
    T = ro["rewards"].shape[0]
    out = []
    for name in FIELDS:
        x = ro[name]
        if name == "recurrent_hidden_states":
            out.append(x[0, env].reshape(len(env), -1))
        else:
            sel = x[:T][:, env]                     # [T, E, ...]
            out.append(sel.reshape(T * len(env), *x.shape[2:]))
    out.append(advantages[:, env].reshape(T * len(env), 1))
    return tuple(out)
