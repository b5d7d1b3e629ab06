"""Test double for amira_b200.device_graph.DeviceGraph backed by the C oracle, so that the host-side
mirror (encoding, materialisation, removal mirroring, upstream binding) can be exercised on a box
without a GPU.  TEST INFRASTRUCTURE ONLY -- the product never uses it."""
import numpy as np

from oracle import c_oracle


class OracleBackedDevice:
    def __init__(self, device=0):
        self.owner = None
        self._g = None
        self._masks = None

    def build(self, ids, off, k, pos_start=None, pos_end=None, on_device=False):
        self._g = c_oracle.COracleGraph(ids, off, k, pos_start, pos_end)
        self.k, self.R, self.has_pos = k, len(off) - 1, pos_start is not None
        return self

    def arrays(self):
        return self._g.arrays()

    def arrays_reads_only(self):
        return {"win_node": self._g.arrays()["win_node"]}

    def sizes(self):
        a = self._g.arrays()
        return {"nodes": len(a["node_cov"]), "edges": len(a["edge_cov"]), "windows": len(a["win_node"])}

    def _with_masks(self, op):
        before = self._g.arrays()
        op()
        after = self._g.arrays()

        def keep(old, new):
            m = np.zeros(len(old), bool)
            j = 0
            for i in range(len(old)):
                if j < len(new) and old[i] == new[j]:
                    m[i] = True
                    j += 1
            assert j == len(new)
            return m

        nk = keep([tuple(r) for r in before["node_key"].tolist()], [tuple(r) for r in after["node_key"].tolist()])
        old_idx = np.flatnonzero(nk)
        ekey = lambda a, remap: list(zip(remap[a["edge_src"]].tolist(), remap[a["edge_tgt"]].tolist(),
                                         (a["edge_sd"] * a["edge_td"]).tolist()))
        ident = np.arange(len(before["node_cov"]))
        ek = keep(ekey(before, ident), ekey(after, old_idx) if len(after["edge_src"]) else [])
        self._masks = (nk, ek)

    def filter_graph(self, a, b):
        self._with_masks(lambda: self._g.filter_graph(a, b))
        return self

    def remove_low_coverage_components(self, c):
        self._with_masks(lambda: self._g.remove_low_coverage_components(c))

    def filter_masks(self):
        return self._masks
