"""Result files of the max-cut path, in the reference's text format
(rlsolver/methods/util_write_read_result.py:28-82, 139-158, 219-240; `calc_result_file_name`
rlsolver/methods/util.py:200-211).  Host-side only: the solution leaves the device once, as a bool row.

    // obj: 273
    // running_duration: 12
    // num_nodes: 100          (write_graph_result with num_nodes)
    // alg_name: greedy
    // <key>: <value>          (info_dict)
    1 2                        node id (1-based), partition (+1 when plus1)
"""
from __future__ import annotations

import os
import string
from typing import List, Optional, Union

import numpy as np
import torch as th

from .util_evaluator import EncoderBase64


def calc_result_file_name(file: str, add_tail: Optional[str] = ''):
    new_file = str(file)
    if 'data' in new_file:
        new_file = new_file.replace('data', 'result')
    splits = new_file.split('/')
    result_dir = new_file.split(splits[-1])[0]
    if result_dir and not os.path.exists(result_dir):
        os.mkdir(result_dir)
    if add_tail is not None:
        new_file = new_file.replace('.txt', '') + add_tail + '.txt'
    return new_file


def _as_int_list(solution) -> List[int]:
    if isinstance(solution, th.Tensor):
        solution = solution.detach().to("cpu")
        return [int(v) for v in solution.reshape(-1).tolist()]
    if isinstance(solution, np.ndarray):
        return [int(v) for v in solution.reshape(-1).tolist()]
    return [int(bool(v)) if isinstance(v, (bool, np.bool_)) else int(v) for v in solution]


def write_graph_result(obj: Union[float, int], running_duration: Optional[int], num_nodes: Optional[int], alg_name: str,
                       solution, filename: str, plus1=True, info_dict: dict = {}):
    solution = _as_int_list(solution)
    add_tail = '_' if running_duration is None else '_' + str(int(running_duration)) if 'data' in filename else None
    new_filename = calc_result_file_name(filename, add_tail)
    while os.path.exists(new_filename):            # never overwrite: append a random letter (reference behaviour)
        assert '.txt' in new_filename
        parts = new_filename.split('.txt')
        assert len(parts) == 2
        letters = string.ascii_lowercase
        new_filename = parts[0] + letters[np.random.randint(0, len(letters))] + '.txt'
    print("result filename: ", new_filename)
    with open(new_filename, 'w', encoding="UTF-8") as new_file:
        prefix = '// '
        new_file.write(f"{prefix}obj: {obj}\n")
        new_file.write(f"{prefix}running_duration: {running_duration}\n")
        if num_nodes is not None:
            new_file.write(f"// num_nodes: {num_nodes}\n")
        new_file.write(f"{prefix}alg_name: {alg_name}\n")
        for key, value in info_dict.items():
            new_file.write(f"{prefix}{key}: {value}\n")
        for i in range(len(solution)):
            new_file.write(f"{i + 1} {solution[i] + 1}\n" if plus1 else f"{i + 1} {solution[i]}\n")
    return new_filename


def write_result(obj: Union[float, int], running_duration: Optional[int], alg_name: str, solution, filename: str,
                 plus1=True, info_dict: dict = {}):
    return write_graph_result(obj, running_duration, None, alg_name, solution, filename, plus1, info_dict)


def obtain_first_number(s: str):
    res = ''
    pass_first_digit = False
    for ch in s:
        if ch.isdigit() or ch == '.':
            res += ch
            pass_first_digit = True
        elif pass_first_digit:
            break
    return int(float(res))


def read_graph_result_comments(filename: str):
    """(num_nodes, ID, running_duration, obj, obj_bound) from the `//` header.  The reference leaves `obj_bound`
    unbound when the file has no such line (UnboundLocalError); here it is None."""
    num_nodes, running_duration, obj, obj_bound = None, None, None, None
    ID = int(filename.split('ID')[1].split('_')[0])
    with open(filename, 'r') as file:
        for line in file:
            if '//' in line:
                if 'num_nodes:' in line:
                    num_nodes = int(line.split('num_nodes:')[1])
                    break
                if 'running_duration:' in line:
                    running_duration = obtain_first_number(line)
                if 'obj:' in line:
                    obj = float(line.split('obj:')[1])
                if 'obj_bound:' in line:
                    obj_bound = float(line.split('obj_bound:')[1])
    return num_nodes, ID, running_duration, obj, obj_bound


def read_solution(filename: str, plus1=True) -> th.Tensor:
    """The partition column of a result file as a bool row (the inverse of write_graph_result)."""
    vals = []
    with open(filename, 'r') as file:
        for line in file:
            if line.startswith('//') or not line.strip():
                continue
            _, part = line.split()[:2]
            vals.append(int(part) - (1 if plus1 else 0))
    return th.tensor(vals, dtype=th.bool)


def calc_obj_maxcut_xstr(x_str: str, filename: str, device=None):
    """Cut value of a base-64 solution string on the graph file (reference: networkx `obj_maxcut`; here the
    packed-spin kernel)."""
    from ..envs.env_L2A import EnvMaxcut
    from .util_read_data import read_mygraph
    device = th.device("cuda") if device is None else device
    sim = EnvMaxcut(mygraph=read_mygraph(filename), device=device)
    x = EncoderBase64(encode_len=sim.num_nodes).str_to_bool(x_str).to(sim.device)
    return int(sim.calculate_obj_values(x[None, :])[0])


__all__ = ["calc_result_file_name", "write_graph_result", "write_result", "read_graph_result_comments",
           "read_solution", "obtain_first_number", "calc_obj_maxcut_xstr"]
