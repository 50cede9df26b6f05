"""Synthetic navigation worlds for the fine-tune / inference path (nav.py): there is no Matterport data or simulator
offline, so tests and `bench.py --workload nav_*` walk random connectivity graphs whose observations follow the
schema the reference agent consumes (map_nav_src/r2r/env.py `_get_obs`: viewpoint, heading, elevation, position,
feature [36, image_feat + angle_feat], candidate = [{viewpointId, pointId, position, feature}], instr_encoding)."""
import math

import numpy as np


def view_angles():
    """Heading / elevation of the 36 discretised views (12 headings x 3 elevations) = `get_view_rel_angles(12)` of
    map_nav_src/utils/data.py:184-200, the table the agent reads (agent.py:1409,1416)."""
    ang = np.zeros((36, 2), dtype=np.float32)
    for i in range(36):
        ang[i] = ((i % 12) * math.radians(30), (i // 12 - 1) * math.radians(30))
    return ang


class NavWorld:
    """`n` viewpoints at random positions, each linked to its 2-5 nearest neighbours (symmetric, connected)."""

    def __init__(self, n=24, seed=0, feat=768, angle_feat=4):
        rng = np.random.RandomState(seed)
        self.n, self.feat, self.angle_feat = n, feat, angle_feat
        self.pos = np.concatenate([rng.uniform(-12, 12, (n, 2)), rng.uniform(-1.5, 1.5, (n, 1))], 1)
        self.ids = [f"vp{seed:02d}_{i:03d}" for i in range(n)]
        d = np.linalg.norm(self.pos[:, None] - self.pos[None], axis=-1)
        adj = np.zeros((n, n), dtype=bool)
        for i in range(n):
            for j in np.argsort(d[i])[1:1 + rng.randint(2, 5)]:
                adj[i, j] = adj[j, i] = True
        order = np.argsort(self.pos[:, 0])  # a spanning chain keeps the graph connected
        for a, b in zip(order[:-1], order[1:]):
            adj[a, b] = adj[b, a] = True
        self.adj = adj
        self.views = rng.randn(n, 36, feat).astype(np.float32)
        self.rng = rng

    def observe(self, i, heading=0.0, elevation=0.0, instr=None):
        ang = view_angles()
        rel = np.stack([np.sin(ang[:, 0] - heading), np.cos(ang[:, 0] - heading), np.sin(ang[:, 1] - elevation),
                        np.cos(ang[:, 1] - elevation)], 1).astype(np.float32)
        feature = np.concatenate([self.views[i], rel], 1)
        nbrs = np.nonzero(self.adj[i])[0][:8]
        points = np.random.RandomState(1000 + i).permutation(36)[:len(nbrs)]  # one distinct view per neighbour
        cands = [{"viewpointId": self.ids[j], "pointId": int(p), "position": self.pos[j].tolist(),
                  "feature": feature[p].copy()} for j, p in zip(nbrs, points)]
        ob = {"viewpoint": self.ids[i], "heading": float(heading), "elevation": float(elevation),
              "position": self.pos[i].tolist(), "feature": feature, "candidate": cands}
        if instr is not None:
            ob["instr_encoding"] = instr
        return ob

    def index(self, vp):
        return self.ids.index(vp)


def make_instr(rng, L):
    n = int(rng.randint(max(4, L // 2), L + 1))
    return [0] + rng.randint(3, 50000, n - 2).tolist() + [2]


def world_tables(world, max_hops_inf=10 ** 6):
    """(pos, dist, hops, cands, view_ang) of a NavWorld in the layout featurizer.GraphWorld takes: all-pairs shortest
    distances / hop counts by Floyd-Warshall, candidate tables in `observe()`'s neighbour order."""
    n = world.n
    d = np.full((n, n), np.inf)
    hops = np.full((n, n), max_hops_inf, dtype=np.int64)
    np.fill_diagonal(d, 0.0)
    np.fill_diagonal(hops, 0)
    e = np.linalg.norm(world.pos[:, None] - world.pos[None], axis=-1)
    d[world.adj] = e[world.adj]
    hops[world.adj] = 1
    for k in range(n):
        via = d[:, k, None] + d[None, k, :]
        better = via < d
        d[better] = via[better]
        hv = hops[:, k, None] + hops[None, k, :]
        hops[better] = np.broadcast_to(hv, (n, n))[better]
    cands = []
    for i in range(n):
        nbrs = np.nonzero(world.adj[i])[0][:8]
        points = np.random.RandomState(1000 + i).permutation(36)[:len(nbrs)]
        cands.append([(int(j), int(p), 0.0, 0.0) for j, p in zip(nbrs, points)])
    return world.pos.astype(np.float64), d.astype(np.float32), hops.astype(np.int32), cands, view_angles()
