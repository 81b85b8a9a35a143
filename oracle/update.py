"""Oracle: whole-update restatements of PPO.update and A2C_ACKTR.update (non-KFAC).

reference algo/ppo.py:34-96 and algo/a2c_acktr.py:33-80.  Autograd and torch.optim
are used where the reference uses them (they ARE the reference's arithmetic); the
explicit optimiser math lives in oracle/optim.py.  `ro` is a dict holding the
RolloutStorage tensors under the reference's attribute names.
"""
import torch

from . import losses, minibatch, policy


def make_leaf_params(p):
    return {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}


def ppo_update(p, spec, ro, hp, optimizer=None, trace=None):
    """p: dict of leaf tensors (updated in place). Returns (value_loss, action_loss, entropy, optimizer)."""
    plist = list(p.values())
    if optimizer is None:
        optimizer = torch.optim.Adam(plist, lr=hp["lr"], eps=hp["eps"])
    T, N = ro["rewards"].shape[:2]
    adv = losses.normalized_advantages(ro["returns"], ro["value_preds"])
    vsum = asum = esum = 0.0
    for _ in range(hp["ppo_epoch"]):
        if spec.recurrent:
            chunks = minibatch.recurrent_env_chunks(N, hp["num_mini_batch"])
            batches = (minibatch.gather_recurrent(ro, env, adv) for env in chunks)
        else:
            chunks = minibatch.feed_forward_index_chunks(T, N, hp["num_mini_batch"])
            batches = (minibatch.gather_feed_forward(ro, idx, adv) for idx in chunks)
        for k, (obs_b, hxs_b, act_b, val_b, ret_b, mask_b, oldlp_b, adv_b) in enumerate(batches):
            values, logp, ent, _ = policy.evaluate_actions(p, spec, obs_b, hxs_b, mask_b, act_b)
            vl, al = losses.ppo_losses(values, logp, val_b, ret_b, oldlp_b, adv_b, hp["clip_param"],
                                       hp.get("use_clipped_value_loss", True))
            optimizer.zero_grad()
            (vl * hp["value_loss_coef"] + al - ent * hp["entropy_coef"]).backward()
            gn = torch.nn.utils.clip_grad_norm_(plist, hp["max_grad_norm"])
            if trace is not None:
                trace.append(dict(idx=chunks[k].clone(), value_loss=vl.item(), action_loss=al.item(),
                                  entropy=ent.item(), grad_norm=float(gn),
                                  grads={n: t.grad.detach().clone() for n, t in p.items()}))
            optimizer.step()
            if trace is not None:
                trace[-1]["params"] = {n: t.detach().clone() for n, t in p.items()}
            vsum += vl.item()
            asum += al.item()
            esum += ent.item()
    n_upd = hp["ppo_epoch"] * hp["num_mini_batch"]
    return vsum / n_upd, asum / n_upd, esum / n_upd, optimizer


def a2c_update(p, spec, ro, hp, optimizer=None, trace=None):
    plist = list(p.values())
    if optimizer is None:
        optimizer = torch.optim.RMSprop(plist, hp["lr"], eps=hp["eps"], alpha=hp["alpha"])
    T, N = ro["rewards"].shape[:2]
    obs = ro["obs"][:-1].reshape(T * N, *ro["obs"].shape[2:])
    values, logp, ent, _ = policy.evaluate_actions(
        p, spec, obs, ro["recurrent_hidden_states"][0].reshape(N, -1),
        ro["masks"][:-1].reshape(T * N, 1), ro["actions"].reshape(T * N, -1))
    vl, al = losses.a2c_losses(values.view(T, N, 1), logp.view(T, N, 1), ro["returns"][:-1])
    optimizer.zero_grad()
    (vl * hp["value_loss_coef"] + al - ent * hp["entropy_coef"]).backward()
    gn = torch.nn.utils.clip_grad_norm_(plist, hp["max_grad_norm"])
    if trace is not None:
        trace.append(dict(grad_norm=float(gn), grads={n: t.grad.detach().clone() for n, t in p.items()}))
    optimizer.step()
    return vl.item(), al.item(), ent.item(), optimizer
