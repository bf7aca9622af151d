"""CPU prototype of the INCREMENTAL mutual-component update (planning tool for the env kernels; not product code).

After removing node a only the component that contained a can change, and inside it only what is reachable from the
end points of edges that just died.  Invariant kept per layer l: every layer-l connected piece of a component either
contains a pending source of layer l, or the component is layer-l connected.  A multi-source search from the pending
sources of one (component, layer) stops as soon as at most one group still has a frontier: the exhausted groups are
split off (fresh labels), the rest keeps its label without being visited.  Splits create new pending sources in the
OTHER layers (end points of the edges that now cross labels).  Checked here against the from-scratch restatement
(oracle.mutual_labels) on random removals: identical partitions, LMCC sizes and alive degrees at every step.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import md_oracle as orc  # noqa: E402
from mdcommunity_b200 import gmm  # noqa: E402


class Incremental:
    def __init__(self, n, layers):
        self.n, self.L = n, len(layers)
        self.adj = []
        for e in layers:
            a = [[] for _ in range(n)]
            for u, v in np.asarray(e).tolist():
                a[u].append(v)
                a[v].append(u)
            self.adj.append(a)
        self.lab = np.zeros(n, np.int64)
        self.next_label = n
        self.visited_total = 0
        self.tasks = 0

    def init_from(self, labels):
        self.lab = labels.copy()

    def remove(self, a):
        lab, L = self.lab, self.L
        c = lab[a]
        lab[a] = -1
        pend = [set() for _ in range(L)]
        for l in range(L):
            for v in self.adj[l][a]:
                if lab[v] == c:
                    pend[l].add(v)
        l = 0
        idle = 0
        while idle < L:
            if not pend[l]:
                idle += 1
                l = (l + 1) % L
                continue
            idle = 0
            srcs = sorted(pend[l])
            pend[l] = set()
            by_lab = {}
            for v in srcs:
                if lab[v] >= 0:
                    by_lab.setdefault(lab[v], []).append(v)
            for x, S in sorted(by_lab.items()):
                if len(S) < 2:
                    continue
                self.tasks += 1
                pieces = self.search(l, x, S)
                for piece in pieces:  # exhausted groups that must leave x
                    newl = self.next_label
                    self.next_label += 1
                    for u in piece:
                        lab[u] = newl
                # pending sources in the other layers: end points of edges that now cross labels inside old x
                moved = set(u for p in pieces for u in p)
                for u in moved:
                    for l2 in range(L):
                        if l2 == l:
                            continue
                        for v in self.adj[l2][u]:
                            if lab[v] >= 0 and lab[v] != lab[u] and (v in moved or lab[v] == x):
                                # the edge (u, v) was alive (both had label x) and is cut now
                                pend[l2].add(u)
                                pend[l2].add(v)
            l = (l + 1) % L

    def search(self, l, x, S):
        """Multi-source search in layer l inside label x.  Returns the list of groups (node lists) to split off."""
        lab, adj = self.lab, self.adj[l]
        own = {}
        parent = list(range(len(S)))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i

        frontier = []
        for i, v in enumerate(S):
            own[v] = i
            frontier.append(v)
        while True:
            # roots that still have a frontier
            active = set(find(own[v]) for v in frontier)
            roots = set(find(i) for i in range(len(S)))
            if len(roots) == 1 or len(active) <= 1:
                break
            nxt = []
            for u in frontier:
                gu = own[u]
                for v in adj[u]:
                    if lab[v] != x:
                        continue
                    if v not in own:
                        own[v] = gu
                        nxt.append(v)
                    else:
                        ra, rb = find(gu), find(own[v])
                        if ra != rb:
                            parent[max(ra, rb)] = min(ra, rb)
            frontier = nxt
        self.visited_total += len(own)
        roots = sorted(set(find(i) for i in range(len(S))))
        if len(roots) == 1:
            return []
        active = set(find(own[v]) for v in frontier)
        groups = {}
        for v, i in own.items():
            groups.setdefault(find(i), []).append(v)
        if active:
            keep = next(iter(active))
        else:
            keep = min(roots, key=lambda r: min(groups[r]))  # all exhausted: the group with the smallest node keeps x
        return [sorted(groups[r]) for r in sorted(roots, key=lambda r: min(groups[r])) if r != keep]


def canon(lab):
    """partition signature: for every alive node the smallest node of its block"""
    out = np.full(len(lab), -1, np.int64)
    first = {}
    for v, x in enumerate(lab.tolist()):
        if x >= 0 and x not in first:
            first[x] = v
    for v, x in enumerate(lab.tolist()):
        if x >= 0:
            out[v] = first[x]
    return out


def main():
    rng = np.random.default_rng(0)
    total_steps = 0
    for trial in range(12):
        n = int(rng.integers(60, 400))
        if trial % 3 == 2:
            L = 3
            layers = gmm.random_sparse_layers(n, L, int(1.6 * n), seed=trial)
        else:
            g = gmm.gmm_graph(n, 100 + trial, cost="unit")
            layers = g.edge_list
        og = orc.MultiplexGraph(n, layers)
        env = orc.OracleEnv("unit")
        env.s0(og)
        inc = Incremental(n, layers)
        inc.init_from(env.labels)
        steps = 0
        while not env.isTerminal():
            active, degs, _ = env.status()
            ids = np.nonzero(active)[0]
            a = int(ids[np.argmax(degs[0][ids] + degs[1][ids])]) if steps % 3 else int(rng.choice(ids))
            env.stepWithoutReward(a)
            inc.remove(a)
            ref = env.labels.copy()
            ref[env.removed] = -1
            assert np.array_equal(canon(inc.lab), canon(ref)), (trial, steps)
            steps += 1
        total_steps += steps
        print("trial %d: n=%d L=%d steps=%d ok; search tasks %d, nodes visited per step %.1f" % (
            trial, n, len(layers), steps, inc.tasks, inc.visited_total / max(1, steps)))
    print("all partitions identical over %d steps" % total_steps)


if __name__ == "__main__":
    main()
