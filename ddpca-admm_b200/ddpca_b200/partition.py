"""Subdomain -> GPU mapping: greedy bin packing of bodies by the size of their
finest operator.  NVSwitch makes all GPU pairs equidistant, so only the load is balanced; bodies
joined by an interface prefer the same rank when that does not hurt the balance."""
from __future__ import annotations


def partition_bodies(weights, contBody, nranks: int, slack: float = 0.05):
    """Return body_rank[v].  weights[v] ~ nnz of consStif[maxiLeve] of body v."""
    nb = len(weights)
    if nranks <= 1:
        return [0] * nb
    order = sorted(range(nb), key=lambda v: -weights[v])
    load = [0.0] * nranks
    rank = [-1] * nb
    neigh = [[] for _ in range(nb)]
    for a, b in contBody:
        neigh[a].append(b)
        neigh[b].append(a)
    total = float(sum(weights))
    for v in order:
        best = min(range(nranks), key=lambda r: load[r])
        # prefer a rank that already holds a neighbour if it stays within `slack` of the lightest one
        cands = {rank[u] for u in neigh[v] if rank[u] >= 0}
        pick = best
        for r in sorted(cands, key=lambda r: load[r]):
            if load[r] + weights[v] <= load[best] + weights[v] + slack * total / nranks:
                pick = r
                break
        rank[v] = pick
        load[pick] += weights[v]
    return rank


def cross_interfaces(contBody, body_rank):
    """Interfaces whose two sides live on different ranks (their traces are exchanged)."""
    return [ts for ts, (a, b) in enumerate(contBody) if body_rank[a] != body_rank[b]]
