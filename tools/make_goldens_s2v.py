"""Generate tests/golden/s2v_*.npz by running the UNMODIFIED reference on CPU: the single-env NumPy SpinSystem of
rlsolver/methods/ECO_S2V/src/envs/spinsystem.py in the S2V-DQN configuration (train_S2V.py:37-47: SPIN_STATE only,
DENSE reward, norm_rewards, irreversible spins), one instance per graph, stacked into batch arrays.  Actions are a
random order of each env's unflipped spins (what an S2V agent's action mask allows); the second case runs until no
spin is left, so `done` fires from the "no more spins to flip" rule before max_steps.
Build container only:  python tools/make_goldens_s2v.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_import  # noqa: E402

ref_import.setup()
from rlsolver.methods.ECO_S2V.src.envs import spinsystem as ss  # noqa: E402
from rlsolver.methods.ECO_S2V.src.envs import util_envs as ue  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def case(name, make_gg, num_envs, max_steps, steps, seed, observables):
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    env_args = dict(observables=observables, reward_signal=ue.RewardSignal.DENSE, extra_action=ue.ExtraAction.NONE,
                    optimisation_target=ue.OptimisationTarget.CUT, spin_basis=ue.SpinBasis.BINARY, norm_rewards=True,
                    memory_length=None, horizon_length=None, stag_punishment=None, basin_reward=None,
                    reversible_spins=False)
    mats, obs0, acts, states, rews, dones, scores, bests = [], [], [], [], [], [], [], []
    for _ in range(num_envs):
        env = ss.SpinSystemFactory.get(make_gg(), max_steps, **env_args)
        n = env.n_spins
        mats.append(np.asarray(env.matrix, np.float32).copy())
        obs0.append(env.get_observation().copy())
        order = rng.permutation(n)[:steps]
        a_e, s_e, r_e, d_e, sc_e, b_e = [], [], [], [], [], []
        for t in range(steps):
            _, rew, done, _ = env.step(int(order[t]))
            a_e.append(order[t]), s_e.append(env.state.copy()), r_e.append(rew), d_e.append(bool(done))
            sc_e.append(env.score), b_e.append(env.best_score)
        acts.append(a_e), states.append(s_e), rews.append(r_e), dones.append(d_e), scores.append(sc_e), bests.append(b_e)
    out = dict(matrix=np.stack(mats), obs0=np.stack(obs0), max_steps=np.asarray(max_steps),
               observables=np.asarray([o.value for o in observables]),
               actions=np.asarray(acts).T.copy(), states=np.asarray(states).transpose(1, 0, 2, 3).copy(),
               rewards=np.asarray(rews, np.float64).T.copy(), dones=np.asarray(dones).T.copy(),
               scores=np.asarray(scores, np.float64).T.copy(), best_scores=np.asarray(bests, np.float64).T.copy())
    p = os.path.join(OUT, f"s2v_{name}.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, {k: v.shape for k, v in out.items()})


def main():
    case("er30_discrete", lambda: ue.RandomERGraphGenerator(30, 0.15, ue.EdgeType.DISCRETE), 12, 60, 20, 601,
         ue.S2V_OBSERVABLES)
    case("ba24_uniform_to_the_end", lambda: ue.RandomBAGraphGenerator(24, 4, ue.EdgeType.UNIFORM), 7, 48, 24, 602,
         [ue.Observable.SPIN_STATE, ue.Observable.IMMEDIATE_REWARD_AVAILABLE, ue.Observable.EPISODE_TIME])


if __name__ == "__main__":
    main()
