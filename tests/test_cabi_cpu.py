"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, and its native graph builder agrees with the oracle / the reference goldens.
No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_files
from oracle import maxcut as om
from synth import gset_like

import rlsolver_b200
from rlsolver_b200 import _lib
from rlsolver_b200.graph_store import GraphStore


@pytest.fixture(scope="module")
def lib():
    rlsolver_b200.build()
    return rlsolver_b200.lib()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "rlsolver_b200.h")).read()
    declared = set(re.findall(r"\b(rlsb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rlsolver_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table and header drifted"
    assert lib.rlsb_version() >= 100


@pytest.mark.parametrize("path", golden_files("maxcut_"), ids=os.path.basename)
def test_graph_builder_matches_reference(lib, path):
    z = np.load(path)
    edges = [tuple(int(t) for t in row) for row in z["edges"]]
    bidir = bool(z["bidirectional"])
    st = GraphStore(edges, bidir, device=None)
    assert st.num_nodes == int(z["num_nodes"]) and st.num_edges == int(z["num_edges"])
    arrs = st.export()
    n0 = np.repeat(np.arange(st.num_nodes), np.diff(arrs["listed_ptr"]))
    assert np.array_equal(n0, z["n0_ids"]) and np.array_equal(arrs["listed_col"], z["n1_ids"])
    assert np.array_equal(np.diff(arrs["listed_ptr"]), z["n0_num_n1"])
    g = om.build_graph_store(edges, bidir)
    fptr, fcol = om.full_neighbourhood(g)
    assert np.array_equal(arrs["full_ptr"], fptr) and np.array_equal(arrs["full_col"], fcol)


@pytest.mark.parametrize("name", ["G14", "G22", "G70"])
def test_levels_are_a_valid_schedule(lib, name):
    edges = gset_like(name)
    st = GraphStore(edges, True, device=None)
    a = st.export()
    level_of = np.empty(st.num_nodes, np.int64)
    for l in range(st.num_levels):
        nodes = a["level_nodes"][a["level_ptr"][l]:a["level_ptr"][l + 1]]
        assert np.all(np.diff(nodes) > 0)
        level_of[nodes] = l
    assert sorted(a["level_nodes"].tolist()) == list(range(st.num_nodes))
    src = np.repeat(np.arange(st.num_nodes), np.diff(a["full_ptr"]))
    dst = a["full_col"]
    lo = src > dst          # every edge: the later node sits on a strictly later level
    assert np.all(level_of[src[lo]] > level_of[dst[lo]])
    assert np.all(level_of[src] != level_of[dst])


def test_builder_errors(lib):
    with pytest.raises(IndexError):      # isolated node shrinks N below the largest id (reference IndexError)
        GraphStore([(0, 1, 1), (1, 5, 1)], False, device=None)
    st = GraphStore([], False, device=None)
    assert st.num_nodes == 0 and st.num_edges == 0


def test_device_ops_refuse_host_only_graph(lib):
    st = GraphStore([(0, 1, 1), (1, 2, 1)], False, device=None)
    out = (C.c_int64 * 4)()
    rc = lib.rlsb_cut_eval(st.handle, C.c_void_p(1), 4, out, None)
    assert rc == 4 and b"host-only" in lib.rlsb_last_error()


def test_no_cpu_fallback():
    import torch as th
    from rlsolver_b200.envs.env_L2A import EnvMaxcut
    with pytest.raises(RuntimeError, match="CUDA"):
        EnvMaxcut(mygraph=[(0, 1, 1)], device=th.device("cpu"))


def _sell_cross(sell, words_of):
    """NumPy model of vcount.cuh sell_cross: per slot, per env, neighbours on the other side.
    Column ids come in blocks of 4 rounds, lane-major inside a block."""
    out = {}
    for s in range(len(sell["off"]) - 1):
        rb, re_ = int(sell["off"][s]), int(sell["off"][s + 1])
        cols = sell["col"][rb * 128:re_ * 128].reshape(re_ - rb, 32, 4)
        for lane in range(32):
            node = int(sell["node"][s * 32 + lane])
            if node == 0xFFFF:
                continue
            nb = cols[:, lane, :].reshape(-1).astype(np.int64)
            out[node] = (words_of[:, nb] ^ words_of[:, [node]]).sum(axis=1), int(sell["half"][s * 32 + lane]), s
    return out


@pytest.mark.parametrize("path", golden_files("maxcut_"), ids=os.path.basename)
def test_sell_structures_model_the_reference(lib, path):
    """The SELL-32 slices the tile kernels walk reproduce (a) the per-node cross counts and
    (b), level by level, the exhaustive single-flip pass of the reference golden."""
    z = np.load(path)
    edges = [tuple(int(t) for t in row) for row in z["edges"]]
    bidir = bool(z["bidirectional"])
    st = GraphStore(edges, bidir, device=None)
    g = om.build_graph_store(edges, bidir)
    n, npad = st.num_nodes, st.padded_nodes
    xs = np.zeros((z["xs"].shape[0], npad), bool)
    xs[:, :n] = z["xs"]
    listed = st.export_sell(0)
    assert len(listed["off"]) - 1 == npad // 32
    assert np.array_equal(listed["node"], np.arange(npad, dtype=np.uint16))
    cc = _sell_cross(listed, xs)
    got = np.stack([cc[i][0] for i in range(n)], axis=1)
    assert np.array_equal(got, om.node_cross_counts_raw(g, z["xs"]))
    # sweep: slices of one level in any order, levels in order
    sweep = st.export_sell(1)
    ls = sweep["level_slice"]
    assert len(ls) == st.num_levels + 1 and ls[-1] == len(sweep["off"]) - 1
    seen = []
    cur = xs.copy()
    for l in range(st.num_levels):
        # evaluate every slot of the level on the state before the level, then apply
        res = _sell_cross({"off": sweep["off"][ls[l]:ls[l + 1] + 1], "node": sweep["node"][ls[l] * 32:ls[l + 1] * 32],
                           "half": sweep["half"][ls[l] * 32:ls[l + 1] * 32], "col": sweep["col"]}, cur)
        nxt = cur.copy()
        for node, (cross, half, _) in res.items():
            flip = cross <= half
            nxt[flip, node] = ~nxt[flip, node]
            seen.append(node)
        cur = nxt
    assert sorted(seen) == list(range(n))
    want_xs, want_vs = z["xs"].copy(), om.cut_values(g, z["xs"]).astype(np.int64)
    om.sweep_literal(g, want_xs, want_vs)
    assert np.array_equal(cur[:, :n], want_xs)


def test_metro_workspace_bytes_is_host_arithmetic():
    """rlsb_metro_workspace_bytes needs no device: 8 bytes of draws per (iteration, chain), the accept words, the
    per-iteration counters, the packed tiles and the rate table, each section rounded up to 256 bytes."""
    import rlsolver_b200
    lib = rlsolver_b200.lib()
    n, c, t = 2000, 4096, 1000
    np_, tiles = 2016, 128
    up = lambda b: (b + 255) // 256 * 256          # noqa: E731
    want = up(t * c * 8) + up(t * tiles * 4) + up(t * 4) + up(tiles * np_ * 4) + up(2 * n * 4) + 256
    assert lib.rlsb_metro_workspace_bytes(n, c, t) == want
    assert lib.rlsb_metro_workspace_bytes(0, c, t) == -1 and lib.rlsb_metro_workspace_bytes(n, -1, t) == -1
    assert lib.rlsb_metro_workspace_bytes(37, 33, 0) > 0
