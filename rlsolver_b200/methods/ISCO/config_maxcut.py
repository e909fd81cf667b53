"""Run constants of the ISCO / PISCO max-cut samplers (same names and defaults as
rlsolver/methods/ISCO/config_maxcut.py:1-12).  The sampler classes read them when constructed,
so a driver may assign to them first (e.g. `config_maxcut.BATCH_SIZE = 512`)."""
import torch as th

INIT_TEMPERATURE = 1.0
FINAL_TEMPERATURE = 0
CHAIN_LENGTH = 200
BATCH_SIZE = 1
DATAPATH = "../../../rlsolver/data/syn_BA/BA_100_ID0.txt"
GPU_ID = 0

DEVICE = th.device(f"cuda:{GPU_ID}")     # CUDA only: there is no CPU path in rlsolver_b200
