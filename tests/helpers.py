"""Shared by the GPU parity tests, smoke() and bench.py: oracle state -> kernel argument dicts."""
import torch

import lrp_oracle as O
from lrpx import decoder as D

_GRID_RENAME = {"feat": "F", "st": "s"}
_AOA_RENAME = {"feat": "F"}


def gridtd_kernel_state(oracle_states, device):
    """oracle.gridtd_explainer_forward dicts (one per image) -> stacked kernel state."""
    keys = D.GRIDTD_IMAGE_KEYS + D.GRIDTD_STEP_KEYS + D.GRIDTD_STEP1_KEYS
    conv = [{k: st[_GRID_RENAME.get(k, k)] for k in keys} for st in oracle_states]
    return D.stack_states(conv, D.GRIDTD_IMAGE_KEYS, D.GRIDTD_STEP_KEYS, D.GRIDTD_STEP1_KEYS, device)


def aoa_kernel_state(oracle_states, device):
    keys = D.AOA_IMAGE_KEYS + D.AOA_STEP_KEYS + D.AOA_STEP1_KEYS
    conv = [{k: st[_AOA_RENAME.get(k, k)] for k in keys} for st in oracle_states]
    return D.stack_states(conv, D.AOA_IMAGE_KEYS, D.AOA_STEP_KEYS, D.AOA_STEP1_KEYS, device)


def to_dev(d, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in d.items()}
