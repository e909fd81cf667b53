"""Host-side pieces of the ISCO mirror that need no GPU."""
import torch as th


def test_load_data_matches_params(tmp_path):
    """util_maxcut.load_data: same dict as the reference's loader for a small file with a duplicate
    edge (later weight wins, edge reported once) and a comment line."""
    from rlsolver_b200.methods.ISCO.util_maxcut import load_data
    p = tmp_path / "g.txt"
    p.write_text("5 5\n1 2 1\n// comment\n2 3 1\n3 1 1\n4 5 1\n2 1 1\n")
    d = load_data(str(p), device="cpu")
    assert d["num_nodes"] == 5 and d["num_edges"] == 5
    assert d["edge_from"].tolist()[:4] == [0, 0, 1, 3] and d["edge_to"].tolist()[:4] == [1, 2, 2, 4]
    assert d["adj_matrix"].shape == (8, 8) and d["adj_matrix"].dtype == th.float16
    assert float(d["adj_matrix"].sum()) == 8.0


def test_path_helpers_shapes_and_mask_counts():
    """multinomial picks exactly path_length[b] sites per chain and scores only those."""
    from oracle.isco import mh_step, multinomial
    th.manual_seed(3)
    lp = th.log_softmax(th.randn(6, 23), dim=-1)
    pl = th.tensor([1, 2, 5, 23 - 1, 7, 3])
    sel, ll = multinomial(lp, pl)
    assert sel["selected_mask"].sum(dim=1).tolist() == pl.tolist()
    assert bool(((ll != 0) <= sel["selected_mask"].bool()).all()) and bool((ll <= 0).all())
    y, acc = mh_step(th.tensor([0.0, -1e30]), th.zeros(2, 4), th.ones(2, 4))
    assert acc.tolist() == [True, False] and y.sum(dim=1).tolist() == [4.0, 0.0]
