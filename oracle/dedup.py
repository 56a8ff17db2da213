"""Oracle restatement of the reference's overlap de-duplication.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/src/classpose/entrypoints/predict_wsi.py:896-965
(`deduplicate`): scipy `KDTree.query_pairs(max_dist)` on the centroids, greedy grouping of the pairs, keep the largest
cell of every group.  The reference iterates a Python *set* of pairs, so for chains of three or more linked cells its
result depends on set order; `reference_greedy` makes the order explicit (sorted pairs).  For isolated pairs and
cliques -- what overlapping tiles produce -- every order gives the same result, which `components_keep_largest`
(the rule the CUDA path implements) also gives.  scipy is the library the reference calls, so the pair set is pinned.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import KDTree


def query_pairs(centers, max_dist=7.5):
    return sorted(KDTree(np.asarray(centers, np.float64)).query_pairs(max_dist))


def reference_greedy(centers, sizes, max_dist=7.5, pairs=None):
    """Keep mask of the reference algorithm, pairs visited in sorted order."""
    pairs = query_pairs(centers, max_dist) if pairs is None else pairs
    groups, member_to_group = {}, {}
    for a, b in pairs:
        if a not in member_to_group and b not in member_to_group:
            g = len(groups)
            groups[g] = []
            member_to_group[a] = g
            member_to_group[b] = g
        else:
            g = member_to_group[a] if a in member_to_group else member_to_group[b]
        if a not in groups[g]:
            groups[g].append(a)
        if b not in groups[g]:
            groups[g].append(b)
    remove = set()
    for g in groups.values():
        if len(g) > 1:
            largest = g[int(np.argmax([sizes[i] for i in g]))]
            remove.update(i for i in g if i != largest)
    keep = np.ones(len(sizes), bool)
    keep[list(remove)] = False
    return keep


def components_keep_largest(centers, sizes, max_dist=7.5):
    """Connected components of the 'within max_dist' relation; the largest cell of each survives (ties: lowest index)."""
    n = len(sizes)
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in query_pairs(centers, max_dist):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    best = {}
    for i in range(n):
        r = find(i)
        if r not in best or sizes[i] > sizes[best[r]]:
            best[r] = i
    keep = np.zeros(n, bool)
    keep[list(best.values())] = True
    return keep
