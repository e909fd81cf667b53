"""Import helpers for running the UNMODIFIED reference on CPU in the build container.

Only used by tools/make_goldens*.py (fixture generation) -- never at test or
bench time: /root/reference does not exist on the GPU box.
The reference imports `plotly` unconditionally (rlsolver/methods/util_read_data.py:3-4)
and MCPG needs `torch_geometric.data.Data` / `torch_scatter`; stub them.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("RLSOLVER_REF", "/root/reference")


def setup():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("plotly", "plotly.io", "plotly.graph_objects", "torch_scatter"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["torch_scatter"].scatter = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgd = types.ModuleType("torch_geometric.data")

        class Data:  # minimal stand-in: attribute bag with num_edges
            def __init__(self, **kw):
                self.__dict__.update(kw)

            @property
            def num_edges(self):
                return self.edge_index.shape[1]

            def to(self, device):   # torch_geometric moves every tensor attribute
                for k, v in list(self.__dict__.items()):
                    if hasattr(v, "to") and hasattr(v, "dtype"):
                        self.__dict__[k] = v.to(device)
                return self

        tgd.Data = Data
        tg.data = tgd
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.data"] = tgd


def load_by_path(modname: str, relpath: str, extra_sys_path=()):
    for p in extra_sys_path:
        full = os.path.join(REF, p)
        if full not in sys.path:
            sys.path.insert(0, full)
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
