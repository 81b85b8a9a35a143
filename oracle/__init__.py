"""CPU oracle for the rollout-batch policy-update path.

TEST INFRASTRUCTURE ONLY.  This package restates, on the CPU with plain
torch/numpy fp32 ops, the algorithm of the reference
(ikostrikov/pytorch-a2c-ppo-acktr-gail, files a2c_ppo_acktr/{storage,model,
distributions,utils}.py and a2c_ppo_acktr/algo/{ppo,a2c_acktr,kfac}.py).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it -- always as the checker or the timed CPU baseline,
never as a code path of the product (`pytorch-a2c-ppo-acktr-gail_b200/`).

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: `tests/golden/make_golden.py` imports the unmodified reference from
/root/reference (through `oracle/ref_shim.py`, which only stubs the absent gym /
stable_baselines3 imports and the removed `torch.symeig`), runs it on seeded
synthetic rollouts and commits the results under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every oracle function against those files.
All arithmetic is third-party torch (pinned: torch 2.11.0, the version the golden
files were produced with).
"""
