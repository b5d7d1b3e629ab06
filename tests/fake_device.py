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
        self._read_len = np.diff(np.asarray(off, np.int64))
        return self

    def build_resident(self, encoded, k, device=None):
        return self.build(encoded.ids, encoded.off, k, encoded.pos_start, encoded.pos_end)

    # ---- the post-build scans of the C ABI, restated with numpy on the oracle's arrays ----
    def sizes_early(self):
        a = self._g.arrays()
        return {"nodes": len(a["node_cov"]), "edges": len(a["edge_cov"]), "windows": len(a["win_node"]),
                "short_reads": int(a["is_short"].sum())}

    def read_length_coverages(self, min_lens):
        a = self._g.arrays()
        return np.asarray([int(np.count_nonzero(self._read_len[a["node_reads"]] >= k)) for k in min_lens], np.int64)

    def node_coverage_stats(self):
        c = self._g.arrays()["node_cov"]
        return int(c.sum(dtype=np.int64)), int(c.max()) if len(c) else 0

    def junk_read_mask(self, error_rate):
        a = self._g.arrays()
        out = np.zeros(self.R, np.uint8)
        for r in range(self.R):
            if a["is_short"][r]:
                out[r] = 2
                continue
            w = a["win_node"][a["win_off"][r]:a["win_off"][r + 1]]
            out[r] = int(np.count_nonzero(w < 0)) <= round(len(w) * (1 - error_rate))
        return out

    def nodes_containing(self, ranks):
        key = np.abs(self._g.arrays()["node_key"])
        return np.isin(key, np.asarray(ranks)).any(axis=1) if len(key) else np.zeros(0, bool)

    def linear_steps(self):
        a = self._g.arrays()
        n = len(a["node_cov"])
        deg = (np.diff(a["fw_off"]) + np.diff(a["bw_off"])).astype(np.uint32)
        out = {"degree": deg}
        for side, off, edges in (("fw", a["fw_off"], a["fw_edges"]), ("bw", a["bw_off"], a["bw_edges"])):
            nxt, dr, ext = np.full(n, -1, np.int32), np.zeros(n, np.int8), np.zeros(n, np.uint8)
            for i in range(n):
                cnt = off[i + 1] - off[i]
                if (cnt == 1) if side == "fw" else (cnt > 0):
                    e = edges[off[i]]
                    t = a["edge_tgt"][e]
                    nxt[i], dr[i] = t, a["edge_td"][e]
                    ext[i] = deg[t] in (1, 2) and t != i
            out[side + "_next"], out[side + "_dir"], out[side + "_ext"] = nxt, dr, ext
        return out

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
