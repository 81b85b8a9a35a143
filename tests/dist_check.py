"""Real-NCCL check of the environment-sharded PPO update (run under torchrun on N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank holds N/world environments of the golden C3 rollout; the result must equal the
single-device reference golden (losses 2e-4, parameters at the multi-step noise floor) and be
bit-identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"), HERE):
    sys.path.insert(0, p)

import helpers  # noqa: E402
from a2c_ppo_acktr import algo  # noqa: E402
from a2c_ppo_acktr.model import Policy  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402


class Discrete(object):
    def __init__(self, n):
        self.n = n


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    name = os.environ.get("A2CB_DIST_CASE", "C3")
    case = helpers.load("update_" + name)
    cfg = helpers.case_config(case)
    ro = helpers.rollout_from_case(case, cfg, obs_uint8=True)
    per = cfg["N"] // world
    lo, hi = rank * per, (rank + 1) * per
    torch.manual_seed(1)
    pol = Policy(cfg["obs_shape"], Discrete(6), base_kwargs={"recurrent": cfg["recurrent"]}).to(dev)
    st = RolloutStorage(cfg["T"], per, cfg["obs_shape"], Discrete(6), helpers.hidden_size_of(cfg), obs_dtype=torch.uint8)
    for nme in ("obs", "recurrent_hidden_states", "rewards", "value_preds", "action_log_probs", "actions", "masks",
                "bad_masks"):
        getattr(st, nme).copy_(ro[nme][:, lo:hi])
    st.to(dev)
    if os.environ.get("A2CB_DIST_STAGED", "0") == "1":          # frames ingested from pinned host memory (bench.py's e2e arm)
        host = st.obs.detach().cpu().clone().pin_memory()
        st.obs.fill_(3)
        st.stage_host({"obs": host})
    st.compute_returns(ro["next_value"][lo:hi].to(dev), cfg["use_gae"], cfg["gamma"], cfg["gae_lambda"],
                       cfg["use_proper_time_limits"], schedule=1)
    if cfg["algo"] == "acktr":
        agent = algo.A2C_ACKTR(pol, cfg["value_loss_coef"], cfg["entropy_coef"], acktr=True)
    elif cfg["algo"] == "a2c":
        agent = algo.A2C_ACKTR(pol, cfg["value_loss_coef"], cfg["entropy_coef"], lr=cfg["lr"], eps=cfg["eps"],
                               alpha=cfg["alpha"], max_grad_norm=cfg["max_grad_norm"])
    else:
        agent = algo.PPO(pol, cfg["clip_param"], cfg["ppo_epoch"], cfg["num_mini_batch"], cfg["value_loss_coef"],
                         cfg["entropy_coef"], lr=cfg["lr"], eps=cfg["eps"], max_grad_norm=cfg["max_grad_norm"])
    agent.set_distributed(world, rank, (lo, hi), cfg["N"])
    torch.manual_seed(7)
    losses = agent.update(st)
    # recurrent cases (C5s / C5m): the default sharding draws the reference's global randperm(N), so the result is
    # the reference golden's, minibatch for minibatch
    np.testing.assert_allclose(np.array(losses), case["losses"], rtol=2e-4, atol=2e-6)
    flat = pol._flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "parameters diverged across ranks"
    for k, p in pol.named_parameters():
        helpers.assert_summary_close(helpers.summary(p), case["final_" + k], rtol=1e-2, atol=1e-6, what=k, elem_frac=0.1)
    assert helpers.rel_l2(case, "finaldense_", dict(pol.named_parameters())) <= (1e-4 if cfg["recurrent"] else 2e-3)
    # second update (the balanced recurrent scheme re-uses its shadow rollout and re-lays resident frames out early)
    losses2 = None
    if "losses2" in case and cfg["algo"] == "ppo" and cfg["recurrent"]:
        st.after_update()
        torch.manual_seed(8)
        losses2 = agent.update(st)
        np.testing.assert_allclose(np.array(losses2), case["losses2"], rtol=5e-4, atol=5e-6)
        flat = pol._flat.clone()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "parameters diverged across ranks (second update)"
    dist.barrier()
    if rank == 0:
        comm = getattr(agent, "_comm", None)
        how = "none" if comm is None else ("p2p" if getattr(comm, "p2p", None) is not None else "nccl")
        print("dist_check ok: case=%s world=%d exchange=%s losses=%s second=%s" % (name, world, how, losses, losses2))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
