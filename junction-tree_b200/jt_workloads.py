"""Synthetic networks for BASELINE.json's five configs (definitions: SURVEY.md section 8d).

Each generator returns a dict with ``factors`` (list of variable-label lists), ``sizes``
(``{label: size}``), ``values`` (list of float64 arrays, strictly positive unless stated) and
optionally ``order`` (elimination order), ``evidence_vars``.
"""

import math

import numpy as np


def sprinkler():
    """Config 1: the README network of the reference (``README.md:92-132``)."""
    sizes = {"cloudy": 2, "sprinkler": 2, "rain": 2, "wet_grass": 2}
    factors = [["cloudy"], ["cloudy", "sprinkler"], ["cloudy", "rain"],
               ["rain", "sprinkler", "wet_grass"]]
    values = [
        np.array([0.5, 0.5]),
        np.array([[0.5, 0.5], [0.9, 0.1]]),
        np.array([[0.8, 0.2], [0.2, 0.8]]),
        np.array([[[1, 0], [0.1, 0.9]], [[0.1, 0.9], [0.01, 0.99]]], dtype=np.float64),
    ]
    return {"name": "sprinkler", "factors": factors, "sizes": sizes, "values": values}


def random_dag(n, max_par, smin, smax, window, seed, evidence_frac=0.2):
    """Configs 2 and 5: windowed random DAG with random CPTs (verbatim recipe)."""
    rng = np.random.default_rng(seed)
    labels = ["v%03d" % i for i in range(n)]
    sz = [int(rng.integers(smin, smax + 1)) for _ in range(n)]
    factors = [[labels[0]]]
    for i in range(1, n):
        lo = max(0, i - window)
        k = min(int(rng.integers(1, min(i, max_par) + 1)), i - lo)
        parents = sorted(lo + rng.choice(i - lo, size=k, replace=False))
        factors.append([labels[p] for p in parents] + [labels[i]])
    sizes = {labels[i]: sz[i] for i in range(n)}
    vrng = np.random.default_rng(seed + 1000)
    values = []
    for f in factors:
        t = vrng.random(tuple(sizes[v] for v in f)) + 0.05
        values.append(t / t.sum(axis=-1, keepdims=True))
    n_ev = int(math.ceil(evidence_frac * n))
    evidence_vars = labels[n - n_ev:]
    return {"name": "dag%d" % n, "factors": factors, "sizes": sizes, "values": values,
            "evidence_vars": evidence_vars, "seed": seed}


def dag37():
    """Config 2: n=37, <=4 parents, 2-4 states, window 8, seed 0."""
    return random_dag(37, 4, 2, 4, 8, 0)


def dag500():
    """Config 5: n=500, <=3 parents, 2-8 states, window 8, seed 1."""
    return random_dag(500, 3, 2, 8, 8, 1)


def draw_evidence(net, B, seed=None):
    """int32 [B, |E|] observed states, ``default_rng(seed+2000)``."""
    seed = net.get("seed", 0) if seed is None else seed
    rng = np.random.default_rng(seed + 2000)
    card = np.array([net["sizes"][v] for v in net["evidence_vars"]])
    return rng.integers(0, card, size=(B, len(card))).astype(np.int32)


def ising(n, seed=0):
    """Config 3: binary n x n Ising grid, row-major elimination sweep."""
    rng = np.random.default_rng(seed)
    lab = lambda i, j: "x%02d_%02d" % (i, j)
    factors, values = [], []
    for i in range(n):
        for j in range(n):
            h = rng.normal(0.0, 0.1)
            factors.append([lab(i, j)])
            values.append(np.exp(h * np.array([-1.0, 1.0])))
            if j + 1 < n:
                J = rng.normal(0.0, 0.5)
                factors.append([lab(i, j), lab(i, j + 1)])
                values.append(np.exp(J * np.array([[1.0, -1.0], [-1.0, 1.0]])))
            if i + 1 < n:
                J = rng.normal(0.0, 0.5)
                factors.append([lab(i, j), lab(i + 1, j)])
                values.append(np.exp(J * np.array([[1.0, -1.0], [-1.0, 1.0]])))
    sizes = {lab(i, j): 2 for i in range(n) for j in range(n)}
    order = [lab(i, j) for i in range(n) for j in range(n)]
    return {"name": "ising%dx%d" % (n, n), "factors": factors, "sizes": sizes, "values": values,
            "order": order, "evidence_vars": [lab(n - 1, j) for j in range(n)], "seed": seed}


def large_state_tree(card=(64, 96, 128, 64, 96, 128), seed=0):
    """Config 4: six variables a..f, factors = cliques {a,b,c} {b,c,d} {c,d,e} {d,e,f}."""
    labels = list("abcdef")[:len(card)]
    sizes = dict(zip(labels, card))
    factors = [labels[i:i + 3] for i in range(len(labels) - 2)]
    rng = np.random.default_rng(seed)
    values = [rng.random(tuple(sizes[v] for v in f)) + 0.05 for f in factors]
    return {"name": "large_state_tree", "factors": factors, "sizes": sizes, "values": values,
            "evidence_vars": [labels[-1]], "seed": seed}


def huang_darwiche():
    """8-variable network of Huang & Darwiche (values as in the reference's end-to-end test,
    ``tests/test_junctiontree.py:163-242``)."""
    sizes = {v: 2 for v in "ABCDEFGH"}
    factors = [["A"], ["A", "B"], ["A", "C"], ["B", "D"], ["C", "E"], ["C", "G"],
               ["D", "E", "F"], ["E", "G", "H"]]
    values = [
        np.array([0.5, 0.5]),
        np.array([[0.6, 0.4], [0.5, 0.5]]),
        np.array([[0.8, 0.2], [0.3, 0.7]]),
        np.array([[0.5, 0.5], [0.1, 0.9]]),
        np.array([[0.4, 0.6], [0.7, 0.3]]),
        np.array([[0.9, 0.1], [0.8, 0.2]]),
        np.array([[[0.01, 0.99], [0.99, 0.01]], [[0.99, 0.01], [0.99, 0.01]]]),
        np.array([[[0.05, 0.95], [0.05, 0.95]], [[0.05, 0.95], [0.95, 0.05]]]),
    ]
    return {"name": "huang_darwiche", "factors": factors, "sizes": sizes, "values": values}


def wisconsin():
    """6-variable network (reference ``tests/test_junctiontree.py:422-481``)."""
    sizes = {v: 2 for v in "ABCDEF"}
    factors = [["A"], ["B", "A"], ["C", "A"], ["B", "D"], ["C", "E"], ["D", "E", "F"]]
    values = [
        np.array([0.9, 0.1]),
        np.array([[0.1, 0.9], [0.9, 0.1]]),
        np.array([[0.8, 0.3], [0.2, 0.7]]),
        np.array([[0.3, 0.7], [0.6, 0.4]]),
        np.array([[0.6, 0.4], [0.5, 0.5]]),
        np.array([[[0.2, 0.8], [0.6, 0.4]], [[0.5, 0.5], [0.9, 0.1]]]),
    ]
    return {"name": "wisconsin", "factors": factors, "sizes": sizes, "values": values}
