"""Generate tests/golden/pecoinf_*.npz by running the UNMODIFIED reference on CPU: SpinSystemFactory /
SpinSystemUnbiased of rlsolver/methods/ECO_S2V/src/envs/inference_network_env.py with inference_PECO.py's
configuration (train_and_inference/inference_PECO.py:50-64) on one graph shared by all envs (SetGraphGenerator).
Build container only:  python tools/make_goldens_inference.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.methods.ECO_S2V.src.envs import inference_network_env as ine  # noqa: E402
from rlsolver.methods.ECO_S2V.src.envs import util_envs as ue  # noqa: E402
from rlsolver.methods.ECO_S2V.src.envs import util_envs_PECO as up  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CPU = th.device("cpu")


def case(name, matrix, num_envs, steps, seed, spin_basis, half=False):
    th.manual_seed(seed)
    gg = up.SetGraphGenerator(matrix, device="cpu")
    n = matrix.shape[0]
    env = ine.SpinSystemFactory.get(
        gg, 2 * n, observables=ue.ECO_PECO_OBSERVABLES, reward_signal=ue.RewardSignal.BLS, extra_action=ue.ExtraAction.NONE,
        optimisation_target=ue.OptimisationTarget.CUT, spin_basis=spin_basis, norm_rewards=True, memory_length=None,
        horizon_length=None, stag_punishment=None, basin_reward=1. / 20, reversible_spins=True, if_greedy=False,
        use_tensor_core=half, device=CPU, num_envs=num_envs)
    out = dict(matrix=matrix.numpy().copy(), spins0=env.state[:, 0, :].numpy().copy(), state0=env.state.numpy().copy(),
               score0=env.score.numpy().copy(), best0=np.asarray(env.get_best_cut().numpy()),
               best_spins0=env.best_spins.numpy().copy(), obs0=env.get_observation().numpy().copy(),
               max_steps=np.asarray(2 * n), binary=np.asarray(spin_basis == ue.SpinBasis.BINARY))
    acts, states, dones, scores, bests, best_spins = [], [], [], [], [], []
    for t in range(steps):
        action = acts[-1].clone() if t % 4 == 3 else th.randint(0, n, (num_envs,))
        obs, done = env.step(action)
        acts.append(action), states.append(env.state.numpy().copy()), dones.append(done.numpy().copy())
        scores.append(env.score.numpy().copy()), bests.append(env.get_best_cut().numpy().copy())
        best_spins.append(env.best_spins.numpy().copy())
    out.update(actions=np.stack([a.numpy() for a in acts]), states=np.stack(states), dones=np.stack(dones),
               scores=np.stack(scores), best_scores=np.stack(bests), best_spins=np.stack(best_spins),
               obs_last=obs.numpy().copy())
    p = os.path.join(OUT, f"pecoinf{'half' if half else ''}_{name}.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, {k: v.shape for k, v in out.items() if k in ("matrix", "states", "best0", "best_scores")})


def sym(n, p, seed, weights):
    g = th.Generator().manual_seed(seed)
    a = (th.rand((n, n), generator=g) < p).float().triu(1)
    if weights == "pm1":
        a = a * (2. * th.randint(0, 2, (n, n), generator=g) - 1.)
    elif weights == "float":
        a = a * th.rand((n, n), generator=g)
    return a + a.T


def main():
    case("n40_uniform", sym(40, 0.2, 1, "01"), 13, 16, 701, ue.SpinBasis.BINARY)
    case("n150_pm1", sym(150, 0.06, 2, "pm1"), 9, 12, 702, ue.SpinBasis.SIGNED)
    case("n24_float", sym(24, 0.4, 3, "float"), 6, 10, 703, ue.SpinBasis.BINARY)
    # use_tensor_core=True: state, scores and the observation in float16 (inference_network_env.py:143-145, 212-236)
    case("n40_uniform", sym(40, 0.2, 1, "01"), 13, 16, 704, ue.SpinBasis.BINARY, half=True)
    case("n100_pm1", sym(100, 0.08, 4, "pm1"), 9, 30, 705, ue.SpinBasis.SIGNED, half=True)


if __name__ == "__main__":
    main()
