"""Times Policy.act (the actor-side forward, reference model.py:54-66 / main.py:115-134) on the device.
    python tools/act_bench.py [N ...]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200"))
from a2c_ppo_acktr import _native  # noqa: E402
from a2c_ppo_acktr.model import Policy  # noqa: E402
from a2c_ppo_acktr.storage import RolloutStorage  # noqa: E402


class Discrete(object):
    def __init__(self, n):
        self.n = n


def main():
    dev = torch.device("cuda:0")
    for recurrent in (False, True):
        for N in [int(a) for a in sys.argv[1:]] or [8, 256]:
            torch.manual_seed(1)
            pol = Policy((4, 84, 84), Discrete(6), base_kwargs={"recurrent": recurrent}).to(dev)
            H = pol.recurrent_hidden_state_size
            T = 16
            st = RolloutStorage(T, N, (4, 84, 84), Discrete(6), H, obs_dtype=torch.uint8)
            st.to(dev)
            st.obs.random_(0, 256)

            def loop(fused):
                for t in range(T):
                    if fused:
                        pol.act_into(st, t)
                    else:
                        v, a, lp, h = pol.act(st.obs[t], st.recurrent_hidden_states[t], st.masks[t])
                        st.value_preds[t].copy_(v)
                        st.actions[t].copy_(a)
                        st.action_log_probs[t].copy_(lp)
                        st.recurrent_hidden_states[t + 1].copy_(h)
            for fused in ([False, True] if hasattr(pol, "act_into") else [False]):
                for _ in range(3):                            # eager, graph capture, first replay
                    loop(fused)
                torch.cuda.synchronize()
                l0 = _native.launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0 = time.perf_counter()
                e0.record()
                for _ in range(4):
                    loop(fused)
                e1.record()
                torch.cuda.synchronize()
                w1 = time.perf_counter()
                n = 4 * T
                print("recurrent=%d N=%d fused=%d: %.1f us/act device, %.1f us/act wall, %.1f native launches/act"
                      % (recurrent, N, fused, 1e3 * e0.elapsed_time(e1) / n, 1e6 * (w1 - w0) / n,
                         (_native.launch_count() - l0) / n))


if __name__ == "__main__":
    main()
