"""Pairwise contraction-order heuristic standing in for opt_einsum.paths.greedy.
Only the ORDER of pairwise contractions depends on it, never the arithmetic."""
import itertools


def greedy(inputs, output, size_dict, memory_limit=None):
    terms = [frozenset(t) for t in inputs]
    keep = frozenset(output)
    path = []

    def volume(symbols):
        v = 1
        for s in symbols:
            v *= size_dict[s]
        return v

    while len(terms) > 1:
        choice = None
        for i, j in itertools.combinations(range(len(terms)), 2):
            rest = keep.union(*(t for n, t in enumerate(terms) if n != i and n != j))
            merged = frozenset(s for s in terms[i] | terms[j] if s in rest)
            score = (not (terms[i] & terms[j]), volume(merged) - volume(terms[i]) - volume(terms[j]))
            if choice is None or score < choice[0]:
                choice = (score, i, j, merged)
        _, i, j, merged = choice
        path.append((i, j))
        terms = [t for n, t in enumerate(terms) if n != i and n != j] + [merged]
    return path if path else [(0,)]
