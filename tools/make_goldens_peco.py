"""Generate tests/golden/peco_*.npz by running the UNMODIFIED reference on CPU: SpinSystemFactory /
SpinSystemUnbiased of rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py with the PECO training
configuration (train_PECO.py:34-44) on ER and BA graph batches from util_envs_PECO.py.
Build container only:  python tools/make_goldens_peco.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.methods.ECO_S2V.src.envs import spinsystem_PECO as ss  # noqa: E402
from rlsolver.methods.ECO_S2V.src.envs import util_envs as ue  # noqa: E402
from rlsolver.methods.ECO_S2V.src.envs import util_envs_PECO as up  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CPU = th.device("cpu")


def case(name, gg, max_steps, steps, seed, **env_args):
    th.manual_seed(seed)
    env = ss.SpinSystemFactory.get(gg, max_steps, device=CPU, num_envs=gg.num_envs, **env_args)
    n, e = env.n_spins, env.num_envs
    out = dict(matrix=env.matrix.numpy().copy(), spins0=env.state[:, 0, :].numpy().copy(), state0=env.state.numpy().copy(),
               score0=env.score.numpy().copy(), max_local=env.max_local_reward_available_.numpy().copy(),
               max_steps=np.asarray(max_steps), obs0=env.get_observation().numpy().copy(),
               reward_signal=np.asarray(env_args["reward_signal"].value), norm_rewards=np.asarray(env_args["norm_rewards"]),
               binary=np.asarray(env_args["spin_basis"] == ue.SpinBasis.BINARY),
               basin=np.asarray(-1.0 if env_args.get("basin_reward") is None else env_args["basin_reward"]),
               stag=np.asarray(-1.0 if env_args.get("stag_punishment") is None else env_args["stag_punishment"]))
    acts, states, rews, dones, scores, bests, best_spins = [], [], [], [], [], [], []
    for t in range(steps):
        if t % 3 == 2:      # revisit pressure: undo the previous flip now and then (exercises the history buffer)
            action = acts[-1].clone()
        else:
            action = th.randint(0, n, (e,))
        obs, rew, done = env.step(action)
        acts.append(action), states.append(env.state.numpy().copy()), rews.append(rew.numpy().copy())
        dones.append(done.numpy().copy()), scores.append(env.score.numpy().copy())
        bests.append(env.best_score.numpy().copy()), best_spins.append(env.best_spins.numpy().copy())
    out.update(actions=np.stack([a.numpy() for a in acts]), states=np.stack(states), rewards=np.stack(rews),
               dones=np.stack(dones), scores=np.stack(scores), best_scores=np.stack(bests),
               best_spins=np.stack(best_spins), obs_last=obs.numpy().copy())
    p = os.path.join(OUT, f"peco_{name}.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, {k: v.shape for k, v in out.items() if k in ("matrix", "states", "rewards")})


def main():
    eco = dict(observables=ue.ECO_PECO_OBSERVABLES, reward_signal=ue.RewardSignal.BLS, extra_action=ue.ExtraAction.NONE,
               optimisation_target=ue.OptimisationTarget.CUT, spin_basis=ue.SpinBasis.BINARY, norm_rewards=True,
               memory_length=None, horizon_length=None, stag_punishment=None, basin_reward=1. / 20,
               reversible_spins=True)
    case("er20_discrete_bls", up.RandomERGraphGenerator(20, 0.15, ue.EdgeType.DISCRETE, 24, CPU), 14, 14, 501, **eco)
    case("ba20_uniform_bls", up.RandomBAGraphGenerator(20, 4, ue.EdgeType.UNIFORM, 17, CPU), 10, 10, 502, **eco)
    dense = dict(eco, reward_signal=ue.RewardSignal.DENSE, spin_basis=ue.SpinBasis.SIGNED, norm_rewards=False,
                 basin_reward=None, stag_punishment=0.25, horizon_length=7)
    case("er33_uniform_dense_stag", up.RandomERGraphGenerator(33, 0.2, ue.EdgeType.UNIFORM, 9, CPU), 12, 9, 503, **dense)
    cbls = dict(eco, reward_signal=ue.RewardSignal.CUSTOM_BLS, basin_reward=None,
                observables=[ue.Observable.SPIN_STATE, ue.Observable.EPISODE_TIME, ue.Observable.IMMEDIATE_REWARD_AVAILABLE])
    case("er40_random_cbls", up.RandomERGraphGenerator(40, 0.3, ue.EdgeType.RANDOM, 6, CPU), 8, 8, 504, **cbls)


if __name__ == "__main__":
    main()
