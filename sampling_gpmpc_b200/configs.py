"""Parameter dictionaries for the named workloads, restricted to the keys the hot path reads.

The reference's config system is one yaml per experiment loaded into a plain dict (main.py:34-38); a user
of this package keeps passing that dict.  These builders exist so that bench.py, smoke() and the GPU tests
can construct the same shapes where /root/reference (and its yamls) is absent.  Values are the yaml's:

  car_residual_fs   params/params_car_residual_fs.yaml:10-14 (lf, lr), :21-47 (agent), :75-88 (optimizer)
  pendulum1D        params/params_pendulum1D_samples.yaml:5-13, :15-46, :64-80
  pendulum2D        params/params_pendulum.yaml:1-47, :66-90
"""
from __future__ import annotations

import copy

_CAR_K = [[-2.96060719e-03, -8.05127934e-02, -1.34526882e+00, -8.17447260e-04],
          [-9.35534702e-01, 2.97777988e-02, -3.23316428e-03, -2.19242872e+00]]


def car_residual_fs(num_dyn_samples: int = 4000, steps: int = 50, with_derivatives: bool = False) -> dict:
    """Car forward rollout (benchmarking/simulate_forward_sampling_car.py).  ``with_derivatives=False`` is the
    script as shipped (value-only model on the real data, T=1); True is the iterative-conditioning variant
    (T=3, every sampled point becomes training data)."""
    return {
        "env": {"start": [0.0, 1.95, 0.0, 14.0], "goal_state": [70.0, 1.95, 0.0, 0.0], "dynamics": "bicycle_Bdx",
                "params": {"lf": 1.105, "lr": 1.738}, "prior_dyn_meas": True, "train_data_has_derivatives": False,
                "use_model_without_derivatives": not with_derivatives, "n_data_x": 5, "n_data_u": 9},
        "agent": {"g_dim": {"ny": 3, "nx": 1, "nu": 1}, "dim": {"ny": 3, "nx": 4, "nu": 2},
                  "Dyn_gp_lengthscale": {"both": [[[2.0, 1.1508]], [[1.7, 1.15]], [[5.22931, 1.9544]]]},
                  "Dyn_gp_noise": 1.0e-7, "Dyn_gp_outputscale": {"both": [0.055, 0.075, 0.01]},
                  "Dyn_gp_task_noises": {"val": [1.0, 1.48, 0.515], "multiplier": 1.0e-7},
                  "Dyn_gp_beta": 30.0, "mean_shift_val": 2, "num_dyn_samples": num_dyn_samples,
                  "mean_as_dyn_sample": False, "true_dyn_as_sample": False, "Dyn_gp_jitter": 1.0e-20,
                  "Dyn_gp_variance_is_zero": -1, "Dyn_gp_min_data_dist": -1,
                  "feedback": {"use": True}},
        "common": {"use_cuda": True, "num_MPC_itrs": steps, "dynamics_rejection": False},
        "optimizer": {"H": 1, "u_min": [-0.6, -2], "u_max": [0.6, 2], "x_min": [-2.14, -2.0, -1.0, 9],
                      "x_max": [70, 16.0, 1.0, 16], "SEMPC": {"max_sqp_iter": 2}, "dt": 0.06,
                      "terminal_tightening": {"K": copy.deepcopy(_CAR_K)}},
        "experiment": {"rnd_seed": {"use": True, "value": 123456}},
    }


def pendulum1D_sqp(num_dyn_samples: int = 70, n_mpc: int = 55) -> dict:
    """1-D pendulum closed-loop shape: ns=70, H=17, T=3 -> q=51 joint scalars per SQP linearisation."""
    return {
        "env": {"start": [2.15, 2.3], "goal_state": [3.1416, 0.0], "dynamics": "Pendulum1D", "prior_dyn_meas": True,
                "train_data_has_derivatives": False, "use_model_without_derivatives": False, "n_data_x": 4,
                "n_data_u": 9, "params": {"m": 1.0, "l": 10.0, "g": 9.81}},
        "agent": {"g_dim": {"ny": 1, "nx": 1, "nu": 1}, "dim": {"ny": 2, "nx": 2, "nu": 1},
                  "Dyn_gp_lengthscale": {"both": [[1.84, 1.92]]}, "Dyn_gp_noise": 1.0e-6,
                  "Dyn_gp_outputscale": {"both": [0.03]},
                  "Dyn_gp_task_noises": {"val": [3.8, 1.27, 3.8], "multiplier": 1.0e-6}, "Dyn_gp_beta": 2.5,
                  "mean_shift_val": 2, "num_dyn_samples": num_dyn_samples, "mean_as_dyn_sample": False,
                  "true_dyn_as_sample": False, "Dyn_gp_jitter": 1.0e-6, "Dyn_gp_variance_is_zero": -1,
                  "Dyn_gp_min_data_dist": -1, "feedback": {"use": True}},
        "common": {"use_cuda": True, "num_MPC_itrs": n_mpc, "dynamics_rejection": False},
        "optimizer": {"H": 17, "u_min": [-5.0], "u_max": [5.0], "x_min": [2.1, -2.5], "x_max": [3.6, 2.5],
                      "SEMPC": {"max_sqp_iter": 1}, "dt": 0.015,
                      "terminal_tightening": {"K": [[-18.82703934, -7.32095004]]}},
        "experiment": {"rnd_seed": {"use": True, "value": 123456}},
    }


def pendulum2D_rollout(num_dyn_samples: int = 20, steps: int = 30, min_data_dist: float = 1.0e-4) -> dict:
    """2-D pendulum true-reachable-set rollout (benchmarking/simulate_true_reachable_set.py on params_pendulum.yaml): real
    data WITH derivatives (m = 45*4 = 180 observed scalars), T=4, one point per step, the zero-variance switch
    (params_pendulum.yaml:45) and the min-distance filter of update_hallucinated_Dyn_dataset (:46, 1e-4) on."""
    return {
        "env": {"start": [0.0, 0.0], "goal_state": [2.5, 0.0], "dynamics": "pendulum", "prior_dyn_meas": True,
                "train_data_has_derivatives": True, "use_model_without_derivatives": False, "n_data_x": 3,
                "n_data_u": 5, "params": {"m": 1.0, "l": 1.0, "g": 9.81}},
        "agent": {"g_dim": {"ny": 2, "nx": 2, "nu": 1}, "dim": {"ny": 2, "nx": 2, "nu": 1},
                  "Dyn_gp_lengthscale": {"both": [[[5.2649, 4.5967, 7.0177]], [[3.9696, 2.1265, 6.6749]]]},
                  "Dyn_gp_noise": 1.0e-6, "Dyn_gp_outputscale": {"both": [0.65, 0.55]},
                  "Dyn_gp_task_noises": {"val": [3.8, 1.27, 3.8, 1.27], "multiplier": 1.0e-5}, "Dyn_gp_beta": 2.5,
                  "mean_shift_val": 2, "num_dyn_samples": num_dyn_samples, "mean_as_dyn_sample": False,
                  "true_dyn_as_sample": False, "Dyn_gp_jitter": 1.0e-6, "Dyn_gp_variance_is_zero": 1.1e-6,
                  "Dyn_gp_min_data_dist": min_data_dist, "feedback": {"use": False}},
        "common": {"use_cuda": True, "num_MPC_itrs": steps, "dynamics_rejection": False},
        "optimizer": {"H": 1, "u_min": [-8], "u_max": [8], "x_min": [-2.14, -2.5], "x_max": [2.14, 2.5],
                      "SEMPC": {"max_sqp_iter": 1}, "dt": 0.015, "terminal_tightening": {"K": [[0.0, 0.0]]}},
        "experiment": {"rnd_seed": {"use": True, "value": 123456}},
    }
