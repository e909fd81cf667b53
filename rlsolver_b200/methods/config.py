"""Import surface of rlsolver/methods/config.py that the hot path touches (lines 1-41):
type aliases, GraphType, calc_device and the graph-generation defaults."""
from enum import Enum, unique
from typing import List, Tuple

import torch as th

MyGraph = List[Tuple[int, int, int]]      # (node0, node1, weight), 0-based
MyNeighbor = List[List[int]]


@unique
class GraphType(Enum):
    BA = "BA"   # barabasi_albert
    ER = "ER"   # erdos_renyi
    PL = "PL"   # powerlaw_cluster


def calc_device(gpu_id: int):
    return th.device(f"cuda:{gpu_id}" if th.cuda.is_available() and gpu_id >= 0 else "cpu")


GPU_ID: int = 0
GRAPH_TYPE = GraphType.PL
GRAPH_TYPES: List[GraphType] = [GraphType.ER, GraphType.PL, GraphType.BA]
